timeout 100 python tools/e2e_probe.py 2>&1 | tail -6
NASR_ZEROCOPY=0 timeout 100 python tools/e2e_probe.py 2>&1 | tail -2
timeout 250 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
