# dev: quick GPU check of the ring kernel after a change (tests + per-block times, A/B under debug switches)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for B in 1 8; do
for d in 0 256 4096 16 2; do NASR_RB_DBG=$d python tools/ring_exp.py $B; done
done
