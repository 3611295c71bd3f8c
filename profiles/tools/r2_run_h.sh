mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_all.log 2>&1
timeout 600 python profiles/tools/profile_step.py 64 120 800 fp16 gpurun_out/r2h_kernel_time.md > gpurun_out/r2h_profile.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2h_bench_fp16.json 2> gpurun_out/r2h_bench_fp16.err
tail -n 6 gpurun_out/r2h_all.log; head -24 gpurun_out/r2h_kernel_time.md; head -c 300 gpurun_out/r2h_bench_fp16.json
