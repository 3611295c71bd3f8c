# A/B timing of a full-size train step under different switches (run from the repo root on the GPU box)
for cfg in "T2V_TWO_CHAINS=1" "T2V_TWO_CHAINS=0"; do echo "== $cfg"; env $cfg PROBE_ITERS=6 timeout 200 python profiles/tools/probe_step.py 64 120 800 tf32 2>&1 | tail -2; done
