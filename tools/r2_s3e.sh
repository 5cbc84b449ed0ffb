# dev: memcheck of the new kernel variants + racecheck attempt
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -x -k "tap_pass or 64_channels or sixteen or ckpt_WaveNet_99_egfx or ckpt_GCN_3_egfxset_20240324_160003_48kHz_cond" > gpurun_out/r2_memcheck_new_paths.log 2>&1
tail -12 gpurun_out/r2_memcheck_new_paths.log
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r2_memcheck_new_paths.log
