timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for cfg in "T2V_OVERLAP=1" "T2V_OVERLAP=0" "T2V_OVERLAP=1"; do echo "== $cfg"; env $cfg PROBE_ITERS=6 timeout 200 python profiles/tools/probe_step.py 64 120 800 tf32 2>&1 | tail -2 | head -1; done
