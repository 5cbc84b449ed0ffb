timeout 100 python tools/ring_debug.py 2>&1 | grep -v Dilations | grep "auto" | awk '{print $2,$3,$4,$5}' | tr '\n' ';'; echo
for v in 1 0; do NASR_TMA_STORE=$v timeout 60 python tools/ring_exp.py 8 2>&1 | tail -1 | cut -c1-140; done
for v in 1 0; do NASR_TMA_STORE=$v timeout 60 python tools/ring_exp.py 1 2>&1 | tail -1 | cut -c1-140; done
timeout 250 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
