python bench.py 2>&1 | tail -1 > gpurun_out/bench_ring2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_ring_path.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ring_block|first_block" -s 40 -c 4 -o gpurun_out/r1_ring_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
cut -c1-400 gpurun_out/bench_ring2.log
