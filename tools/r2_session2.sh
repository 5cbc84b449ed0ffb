set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_s2_pytest.log 2>&1; tail -3 gpurun_out/r2_s2_pytest.log
{
python tools/ring_exp.py 1
python tools/ring_exp.py 8
python tools/ring_exp.py 1 480000 cfg3
python tools/ring_timeline.py 1
} > gpurun_out/r2_s2_exp.log 2>&1
cat gpurun_out/r2_s2_exp.log
ncu --set full --import-source on --clock-control none -k regex:ring_block -s 12 -c 2 -o gpurun_out/r2_ring_b1 -f python tools/ring_exp.py 1 > gpurun_out/r2_s2_ncu1.log 2>&1; tail -2 gpurun_out/r2_s2_ncu1.log
ncu --set full --import-source on --clock-control none -k regex:toep_first -s 2 -c 1 -o gpurun_out/r2_toep_b1 -f python tools/ring_exp.py 1 > gpurun_out/r2_s2_ncu2.log 2>&1; tail -2 gpurun_out/r2_s2_ncu2.log
