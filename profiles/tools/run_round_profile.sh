# One GPU call: parity tests, bench line, ncu launch list of the bench command, ncu --set full of both persistent decoder kernels,
# per-kernel time table of the train step.
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()"
(time timeout 420 python -m pytest tests -m gpu -x -q) > gpurun_out/q_tests.log 2>&1; grep "passed\|failed" gpurun_out/q_tests.log
timeout 400 python bench.py > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; cat gpurun_out/q_bench.json
timeout 300 python profiles/tools/profile_step.py 64 120 800 tf32 gpurun_out/q_kernel_time.md > gpurun_out/q_prof.log 2>&1; head -12 gpurun_out/q_kernel_time.md
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4500 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/q_ncu_bench.log 2>&1
wc -l gpurun_out/q_launches.csv
NOTRACE=1 timeout 400 ncu --set full --import-source on --cache-control none --clock-control none -k regex:dec_persist_fwd --launch-skip 2 --launch-count 1 -o gpurun_out/q_ncu_full_persist_fwd python profiles/tools/trace_persist.py 64 120 120 > gpurun_out/q_ncu_full_fwd.log 2>&1
NOTRACE=1 timeout 400 ncu --set full --import-source on --cache-control none --clock-control none -k regex:dec_persist_bwd --launch-skip 1 --launch-count 1 -o gpurun_out/q_ncu_full_persist_bwd python profiles/tools/trace_persist_bwd.py 64 120 120 > gpurun_out/q_ncu_full_bwd.log 2>&1
ls -la gpurun_out/ | grep q_
