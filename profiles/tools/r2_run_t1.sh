#!/bin/bash
# full GPU tests, default bench (no CPU leg), full-step timeline
O=gpurun_out/${1:-t1}; mkdir -p $O
timeout 1300 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log; tail -4 $O/tests.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 > $O/bench.json 2> $O/bench.err
python - $O/bench.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print("%.2f ms/step  e2e %.2f ms  fwd %.2f us  bwd %.2f us  launches %d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_step"], d["decoder_step_backward"]["value"], d["gpu_launches"]))
except Exception as e:
    print("bench FAILED", e)
PY
timeout 600 python profiles/tools/timeline_full.py 64 120 800 fp16 0 > $O/timeline.txt 2>&1
head -4 $O/timeline.txt | tail -2
