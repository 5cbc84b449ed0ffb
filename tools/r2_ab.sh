# dev: streaming A/B inside one session
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "streaming or cfg5 or tensor_core_path" 2>&1 | tail -6
for v in "NASR_CHAIN=1" "NASR_CHAIN=0" "NASR_CHAIN=1 NASR_GATHER_TILES=8" "NASR_CHAIN=0 NASR_GATHER_TILES=8"; do
  echo "== $v"
  env $v timeout 300 python tools/stream_bench.py 2>&1 | grep streaming
done
