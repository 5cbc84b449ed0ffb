# dev: streaming A/B inside one session
mkdir -p gpurun_out
for v in "NASR_GATHER_TILES=2" "NASR_GATHER_TILES=4" "NASR_GATHER_TILES=8" "NASR_GATHER_TILES=2 NASR_PDL=0"; do
  echo "== $v"
  env $v python tools/stream_bench.py 2>&1 | grep streaming
done
python tools/ring_exp.py 1 2>&1 | tail -2
python -m pytest tests -m gpu -q -x -k "streaming or cfg5" 2>&1 | tail -2
