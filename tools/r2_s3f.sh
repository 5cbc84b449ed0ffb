mkdir -p gpurun_out
python tools/ckpt_bench.py --json gpurun_out/s3_ckpt.json 2>&1 | grep -v "^Dilations" > gpurun_out/s3_ckpt.log
NCU="ncu --set full --import-source on --clock-control none -f"
$NCU -k regex:ring_block -s 12 -c 2 -o gpurun_out/r2_ring_gcn16 python tools/ckpt_bench.py --only WaveNet_egfx --iters 2 > gpurun_out/r2_ncu_g.log 2>&1
cat gpurun_out/s3_ckpt.log
cuobjdump -sass neural_audio_spring_reverb_b200/build/default/ring_block.cu.o | grep -oE "UTCHMMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|LDTM[.A-Z0-9_x]*|STTM[.A-Z0-9_x]*|UTCBAR[.A-Z0-9_]*" | sort | uniq -c > gpurun_out/r2_sass_ring.txt
