set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_s1_pytest.log 2>&1; tail -5 gpurun_out/r2_s1_pytest.log
{
for B in 1 8; do
  python tools/ring_exp.py $B
  NASR_RB_DBG=256 python tools/ring_exp.py $B
done
python tools/ring_exp.py 1 480000 cfg3
NASR_RB_DBG=256 python tools/ring_exp.py 1 480000 cfg3
python tools/ring_timeline.py 1
python tools/toep_timeline.py
NASR_RB_DBG=1 python tools/ring_exp.py 1
} > gpurun_out/r2_s1_exp.log 2>&1
cat gpurun_out/r2_s1_exp.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2_s1_bench.log 2>&1; tail -c 3000 gpurun_out/r2_s1_bench.log
