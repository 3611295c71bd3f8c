mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_all.log 2>&1
timeout 600 python profiles/tools/profile_step.py 64 120 800 fp16 gpurun_out/r2n_kernel_time.md > gpurun_out/r2n_profile.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2n_bench_fp16.json 2> gpurun_out/r2n_bench_fp16.err
timeout 600 python bench.py --no-cpu-baseline --precision bf16 > gpurun_out/r2n_bench_bf16.json 2> gpurun_out/r2n_bench_bf16.err
timeout 600 python bench.py --no-cpu-baseline --precision tf32 > gpurun_out/r2n_bench_tf32.json 2> gpurun_out/r2n_bench_tf32.err
tail -n 5 gpurun_out/r2n_all.log; head -30 gpurun_out/r2n_kernel_time.md; for p in fp16 bf16 tf32; do head -c 250 gpurun_out/r2n_bench_$p.json; echo; done
