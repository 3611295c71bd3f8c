mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_units.py -q -s -k "bilstm" > gpurun_out/r2f_units.log 2>&1
timeout 600 python profiles/tools/profile_step.py 64 120 800 fp16 gpurun_out/r2f_kernel_time.md > gpurun_out/r2f_profile.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2f_bench_fp16.json 2> gpurun_out/r2f_bench_fp16.err
tail -n 5 gpurun_out/r2f_units.log; cat gpurun_out/r2f_kernel_time.md; head -c 300 gpurun_out/r2f_bench_fp16.json
