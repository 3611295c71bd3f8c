mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -q -x -s -k "persistent" > gpurun_out/r2o_units.log 2>&1
timeout 300 python profiles/tools/trace_persist_bwd.py > gpurun_out/r2o_trace_bwd.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_oracle_shapes.py -q -s -k "bench_batch" > gpurun_out/r2o_oracle.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2o_bench_fp16.json 2> gpurun_out/r2o_bench_fp16.err
tail -n 4 gpurun_out/r2o_units.log; head -22 gpurun_out/r2o_trace_bwd.txt; grep "^\[\|^\.\[\|^F\[\|grad \|worst\|passed\|failed" gpurun_out/r2o_oracle.log; head -c 300 gpurun_out/r2o_bench_fp16.json
