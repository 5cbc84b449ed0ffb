mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "streaming or tensor_core_path or cfg5 or tap_pass or 64_channels" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --configs cfg5,shipped --cfg5-seconds 600 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['configs']['cfg5']
print('value', d['value'], 'e2e', d['e2e']['value'])
for k in ('chunk_65536','chunk_1024'):
    r=c.get(k) or c; print('  ',k, {q:r.get(q) for q in ('samples_per_s','us_per_chunk','launches_per_chunk','wall_samples_per_s')} if isinstance(r,dict) else r)
if 'error' in c: print(c['error'])
s=d['configs']['shipped']
if 'error' in s: print(s['error'])
else:
    for k,v in s['checkpoints'].items(): print('  ',k, v['block_paths'], v['msamples_per_s'], v['msamples_per_s_fp32'], v['parity_vs_reference_golden'])
"
