mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2s_all.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2s_bench_full.json 2> gpurun_out/r2s_bench_full.err
tail -n 4 gpurun_out/r2s_all.log; tail -n 6 gpurun_out/r2s_smoke.log; head -c 400 gpurun_out/r2s_bench_full.json; tail -2 gpurun_out/r2s_bench_full.err
