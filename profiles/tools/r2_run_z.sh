#!/bin/bash
# full GPU tests + A/B of: in-loop LSTM bias gradients, d(memory) on the tensor core; full-step timeline
O=gpurun_out/${1:-z1}; mkdir -p $O
timeout 1300 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log; tail -5 $O/tests.log
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
run no_bias_in_loop T2V_BIAS_IN_LOOP=0
run no_dmem_tc T2V_DMEM_TC=0
timeout 600 python profiles/tools/timeline_full.py 64 120 800 fp16 0 > $O/timeline.txt 2>&1
head -4 $O/timeline.txt | tail -2
