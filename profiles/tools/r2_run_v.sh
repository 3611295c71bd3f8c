#!/bin/bash
O=gpurun_out/${1:-v1}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
tail -4 $O/tests.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
cut -c1-300 $O/bench.json
timeout 600 python profiles/tools/timeline_full.py 64 120 800 fp16 20 > $O/timeline.txt 2>&1
head -3 $O/timeline.txt | tail -2
