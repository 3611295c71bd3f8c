# One GPU call: parity tests, bench line, ncu launch list of the bench command, ncu --set full of the persistent decoder kernel.
mkdir -p gpurun_out
(time timeout 420 python -m pytest tests -m gpu -x -q) > gpurun_out/q_tests.log 2>&1; tail -3 gpurun_out/q_tests.log
timeout 400 python bench.py > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; cat gpurun_out/q_bench.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/q_ncu_bench.log 2>&1
wc -l gpurun_out/q_launches.csv
NOTRACE=1 timeout 400 ncu --set full --import-source on --cache-control none --clock-control none -k regex:dec_persist --launch-skip 2 --launch-count 1 -o gpurun_out/q_ncu_full_persist python profiles/tools/trace_persist.py 64 120 120 > gpurun_out/q_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -12
