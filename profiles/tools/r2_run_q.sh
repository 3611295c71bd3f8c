mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_units.py -q -x -s -k "rowred16 or fp16_operands" > gpurun_out/r2q_units.log 2>&1
timeout 900 python -m pytest tests/test_gpu_oracle_shapes.py -q -s -k "bench_batch and fp16" > gpurun_out/r2q_oracle.log 2>&1
timeout 600 python profiles/tools/profile_step.py 64 120 800 fp16 gpurun_out/r2q_kernel_time.md > gpurun_out/r2q_profile.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2q_bench_fp16.json 2> gpurun_out/r2q_bench_fp16.err
tail -n 4 gpurun_out/r2q_units.log; grep "^\[\|^\.\[\|^F\[\|grad \|worst\|passed\|failed" gpurun_out/r2q_oracle.log; head -14 gpurun_out/r2q_kernel_time.md; head -c 300 gpurun_out/r2q_bench_fp16.json
