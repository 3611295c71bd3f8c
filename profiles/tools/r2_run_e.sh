mkdir -p gpurun_out
./profiles/tools/sw64_probe > gpurun_out/r2e_sw64_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_units.py -q -s -k "rowred or persistent or 16bit" > gpurun_out/r2e_units.log 2>&1
timeout 300 python profiles/tools/trace_persist.py fp16 > gpurun_out/r2e_trace_fp16.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_oracle_shapes.py -q -s > gpurun_out/r2e_oracle.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_oracle_shapes.py > gpurun_out/r2e_all.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench_fp16.json 2> gpurun_out/r2e_bench_fp16.err
head -20 gpurun_out/r2e_sw64_probe.txt; tail -n 5 gpurun_out/r2e_units.log; grep "^\[" gpurun_out/r2e_oracle.log; tail -n 12 gpurun_out/r2e_all.log; head -c 600 gpurun_out/r2e_bench_fp16.json
