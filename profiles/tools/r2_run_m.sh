mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2m_all.log 2>&1
timeout 300 python profiles/tools/trace_persist_bwd.py > gpurun_out/r2m_trace_bwd.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench_fp16.json 2> gpurun_out/r2m_bench_fp16.err
tail -n 12 gpurun_out/r2m_all.log; head -24 gpurun_out/r2m_trace_bwd.txt; head -c 300 gpurun_out/r2m_bench_fp16.json; tail -3 gpurun_out/r2m_bench_fp16.err
