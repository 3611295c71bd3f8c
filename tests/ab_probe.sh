for cfg in "T2V_ATTN_MODE=row" "T2V_ATTN_MODE=split"; do echo "== $cfg"; env $cfg PROBE_ITERS=6 timeout 200 python tests/probe_step.py 64 120 800 tf32 2>&1 | tail -2; done
