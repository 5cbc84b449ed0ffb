python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"ring_block|toep_first" -s 40 -c 3 -o gpurun_out/r1_final_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
cut -c1-150 gpurun_out/bench_r1_final.log
