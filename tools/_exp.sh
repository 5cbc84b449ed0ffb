timeout 250 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1_final.log; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_final.log')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['steps'], d['clocks']); print(d['roofline']['frac'], d['roofline']['hbm_frac'], [round(x*1e3,1) for x in d['roofline']['block_ms']])"
timeout 100 python bench.py --clips-per-gpu 8 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('B8', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 120 python tools/stream_bench.py 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
