#!/bin/bash
# A/B of the round's scheduling changes: merged-tap weight gradients, branch priorities, Postnet dW beside the backward loop
O=gpurun_out/${1:-y1}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_units.py -x -q -k "rowred" > $O/tests_rowred.log 2>&1; tail -3 $O/tests_rowred.log
run() { # name, env...
  n=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 > $O/bench_$n.json 2> $O/bench_$n.err
  python - $O/bench_$n.json $n <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print("%-22s %.2f ms/step  e2e %.2f ms  fwd %.2f us  bwd %.2f us" % (sys.argv[2], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_step"], d["decoder_step_backward"]["value"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run all X=1
run no_taps1 T2V_TAPS1=0
run no_post_branch T2V_POST_DW_BRANCH=0
run no_prio T2V_PRIO=0
run old T2V_TAPS1=0 T2V_POST_DW_BRANCH=0 T2V_PRIO=0
timeout 600 python profiles/tools/timeline_full.py 64 120 800 fp16 20 > $O/timeline.txt 2>&1
head -4 $O/timeline.txt | tail -2
