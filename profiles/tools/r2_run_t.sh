#!/bin/bash
O=gpurun_out/t3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_units.py -x -q -k "backward or persistent" > $O/tests_units.log 2>&1; echo "exit $?" >> $O/tests_units.log
tail -5 $O/tests_units.log
timeout 900 python -m pytest tests/test_gpu_oracle_shapes.py -x -q -s > $O/tests_oracle.log 2>&1; echo "exit $?" >> $O/tests_oracle.log
tail -5 $O/tests_oracle.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
cut -c1-300 $O/bench.json; grep -o '"decoder_step_backward": {"value": [0-9.]*' $O/bench.json
T2V_PERSIST_TRACE=1 timeout 300 python profiles/tools/trace_persist_bwd.py > $O/trace_bwd.txt 2>&1; head -40 $O/trace_bwd.txt
