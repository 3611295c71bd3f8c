mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -q -x -k "rowred" > gpurun_out/r2l_units.log 2>&1
timeout 900 python -m pytest tests/test_gpu_oracle_shapes.py -q -s -k "bench_batch and fp16" > gpurun_out/r2l_oracle.log 2>&1
timeout 600 python profiles/tools/profile_step.py 64 120 800 fp16 gpurun_out/r2l_kernel_time.md > gpurun_out/r2l_profile.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2l_bench_fp16.json 2> gpurun_out/r2l_bench_fp16.err
tail -n 3 gpurun_out/r2l_units.log; grep "^\[\|grad \|worst\|passed\|failed" gpurun_out/r2l_oracle.log; head -16 gpurun_out/r2l_kernel_time.md; head -c 300 gpurun_out/r2l_bench_fp16.json
