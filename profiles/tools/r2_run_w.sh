#!/bin/bash
# HEAD confirmation run: GPU tests, default bench line (with the CPU baseline leg), full-step timeline
O=gpurun_out/${1:-w1}; mkdir -p $O
S=$(date +%s)
timeout 1300 python -m pytest tests -m gpu -x -q --durations=15 > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
tail -4 $O/tests.log; echo "tests $(( $(date +%s) - S )) s"
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
cut -c1-300 $O/bench.json; echo "bench $(( $(date +%s) - S )) s"
timeout 600 python profiles/tools/timeline_full.py 64 120 800 fp16 20 > $O/timeline.txt 2>&1
head -3 $O/timeline.txt | tail -2; echo "timeline $(( $(date +%s) - S )) s"
