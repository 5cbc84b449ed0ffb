# round-2 ncu evidence (one B200): launch list of the bench command + --set full captures of the dominant kernels
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs "" > gpurun_out/r2_launches_bench.log 2>&1
NCU="ncu --set full --import-source on --clock-control none -f"
$NCU -k regex:ring_block -s 12 -c 2 -o gpurun_out/r2_ring_b1_final python tools/ring_exp.py 1 > gpurun_out/r2_ncu_a.log 2>&1
$NCU -k regex:ring_block -s 9 -c 1 -o gpurun_out/r2_ring_b64 python tools/ring_exp.py 64 > gpurun_out/r2_ncu_b.log 2>&1
$NCU -k regex:ring_block -s 12 -c 1 -o gpurun_out/r2_ring_gcn_b1 python tools/ring_exp.py 1 480000 cfg3 > gpurun_out/r2_ncu_c.log 2>&1
$NCU -k regex:toep_first -s 2 -c 1 -o gpurun_out/r2_toep_b1_final python tools/ring_exp.py 1 > gpurun_out/r2_ncu_d.log 2>&1
tail -3 gpurun_out/r2_ncu_?.log
ls -la gpurun_out/*.ncu-rep
