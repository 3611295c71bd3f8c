timeout 600 python -m pytest tests/test_gpu_units.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
PROBE_ITERS=6 timeout 200 python profiles/tools/probe_step.py 64 120 800 tf32 2>&1 | tail -2 | head -1
TL_COUNT=20 timeout 300 python profiles/tools/timeline_step.py 64 120 800 tf32 2>&1 | grep -v Warn | tail -52
