# dev: quick GPU check after a change
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/e2e_breakdown.py 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --configs cfg4 --no-cpu-baseline > gpurun_out/r2_bench_b.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_b.log").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"])
c=d["configs"]["cfg4"]; print("cfg4",c.get("samples_per_s"),c.get("e2e"),c.get("error"))
PY
NASR_HOST_PIPE=0 python bench.py --steps 20 --warmup 5 --configs cfg4 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']['cfg4']; print('no-pipe cfg4', c.get('samples_per_s'), c.get('e2e'))"
