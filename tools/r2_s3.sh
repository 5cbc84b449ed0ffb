# dev: session-3 GPU check: new tap-pass / 16-channel paths, full GPU suite, per-checkpoint table
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "tap_pass or sixteen or golden" 2>&1 | tail -15 > gpurun_out/s3_new.log
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/s3_all.log
python tools/ckpt_bench.py --json gpurun_out/s3_ckpt.json > gpurun_out/s3_ckpt.log 2>&1
cat gpurun_out/s3_new.log gpurun_out/s3_all.log gpurun_out/s3_ckpt.log | tail -60
