mkdir -p gpurun_out
./profiles/tools/mn_probe > gpurun_out/r2b_mn_probe.txt 2>&1
timeout 300 python profiles/tools/trace_persist.py fp16 > gpurun_out/r2b_trace_fp16.txt 2>&1
timeout 300 python profiles/tools/trace_persist.py tf32 > gpurun_out/r2b_trace_tf32.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_oracle_shapes.py -q -s -k "not bench_batch" > gpurun_out/r2b_oracle.log 2>&1
head -60 gpurun_out/r2b_mn_probe.txt
