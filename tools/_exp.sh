for d in 0 1 2 3; do NASR_TOEP_DBG=$d timeout 60 python tools/ring_exp.py 1 2>&1 | tail -1 | cut -c1-120; done
for d in 0 1 2 3; do NASR_TOEP_DBG=$d timeout 60 python tools/ring_exp.py 8 2>&1 | tail -1 | cut -c1-120; done
