#!/bin/bash
# generic A/B: r2_run_ab.sh OUT "name1:ENV=1 ENV2=0" "name2:..." ...   then a full-step timeline with the default settings
O=gpurun_out/$1; shift; mkdir -p $O
for spec in "$@"; do
  n=${spec%%:*}; e=${spec#*:}
  env $e timeout 300 python bench.py --no-cpu-baseline --steps 10 > $O/bench_$n.json 2> $O/bench_$n.err
  python - $O/bench_$n.json $n <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print("%-22s %.2f ms/step  e2e %.2f ms  fwd %.2f us  bwd %.2f us  infer %.2f us  stft %.3f ms frac %.3f" % (sys.argv[2], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_step"], d["decoder_step_backward"]["value"], d["decoder_step_inference"]["value"], d["stft_mel"]["ms"], d["stft_mel"]["frac"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
timeout 600 python profiles/tools/timeline_full.py 64 120 800 fp16 0 > $O/timeline.txt 2>&1
head -4 $O/timeline.txt | tail -2
