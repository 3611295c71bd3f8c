mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_units.py -q -k "rowred or pack_step_tiles16 or 16bit" -s > gpurun_out/r2a_units.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_oracle_shapes.py -q -s > gpurun_out/r2a_oracle.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_oracle_shapes.py > gpurun_out/r2a_all.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_fp16.json 2> gpurun_out/r2a_bench_fp16.err
timeout 600 python bench.py --precision tf32 --no-cpu-baseline > gpurun_out/r2a_bench_tf32.json 2> gpurun_out/r2a_bench_tf32.err
tail -5 gpurun_out/r2a_units.log gpurun_out/r2a_oracle.log gpurun_out/r2a_all.log
cat gpurun_out/r2a_bench_fp16.json
