# dev: quick GPU check after a change
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in "NASR_STREAM_GRAPH=1 NASR_SMALL_GATHER=1" "NASR_STREAM_GRAPH=0 NASR_SMALL_GATHER=1" "NASR_STREAM_GRAPH=1 NASR_SMALL_GATHER=0" "NASR_STREAM_GRAPH=0 NASR_SMALL_GATHER=0"; do
env $v python bench.py --steps 10 --warmup 3 --configs cfg5 --cfg5-seconds 600 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']['cfg5']
print('$v')
for k in ('chunk_65536','chunk_1024'):
    r=c.get(k) or c; print('  ',k, {q:r.get(q) for q in ('samples_per_s','us_per_chunk','launches_per_chunk','wall_samples_per_s')} if isinstance(r,dict) else r)
if 'error' in c: print(c['error'])"
done
