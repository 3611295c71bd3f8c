mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -q -x -k "persistent or 16bit" > gpurun_out/r2r_units.log 2>&1
timeout 300 python profiles/tools/trace_persist.py fp16 > gpurun_out/r2r_trace_fp16.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_oracle_shapes.py tests/test_gpu_parity.py -q > gpurun_out/r2r_oracle.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2r_bench_fp16.json 2> gpurun_out/r2r_bench_fp16.err
tail -n 4 gpurun_out/r2r_units.log; head -30 gpurun_out/r2r_trace_fp16.txt; tail -n 3 gpurun_out/r2r_oracle.log; head -c 300 gpurun_out/r2r_bench_fp16.json
