#!/bin/bash
# Final evidence of round 2 on the shipped build: GPU tests, smoke, bench lines (fp16 default with the CPU leg, bf16, tf32), 200-step
# soak, kernel-time table, ncu launch list of the bench command, ncu --set full per kernel family, compute-sanitizer.
# Timings are never taken under ncu / the sanitizer.  The cluster kernels are captured with T2V_COOP=0 (ncu rejects cooperative + cluster).
O=gpurun_out/${1:-fin2}; mkdir -p $O
S=$(date +%s); el() { echo "[$(( $(date +%s) - S )) s] $1"; }
timeout 900 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log; tail -3 $O/tests.log; el tests
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "exit $?" >> $O/smoke.log; tail -2 $O/smoke.log; el smoke
timeout 600 python bench.py > $O/bench_fp16.json 2> $O/bench_fp16.err; echo "exit $?" >> $O/bench_fp16.err; cut -c1-260 $O/bench_fp16.json; el bench
for prec in bf16 tf32; do
  timeout 300 python bench.py --precision $prec --no-cpu-baseline > $O/bench_$prec.json 2> $O/bench_$prec.err; cut -c80-200 $O/bench_$prec.json
done; el "bench bf16 tf32"
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > $O/soak_200.json 2> $O/soak_200.err; echo "exit $?" >> $O/soak_200.err; cut -c80-200 $O/soak_200.json; el soak
timeout 300 python profiles/tools/profile_step.py 64 120 800 fp16 $O/kernel_time_C3_step_fp16.md > $O/profile_step.out 2>&1; el "kernel time table"
timeout 300 python profiles/tools/timeline_full.py 64 120 800 fp16 0 > $O/timeline_full.txt 2>&1; el timeline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_bench.out 2>&1; el "launch list"
# ---- ncu --set full: (a) the two persistent loops, (b) everything else of one train step, (c) free-running decode + STFT
T2V_COOP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dec_persist" -c 2 -o $O/ncu_loops -f python profiles/tools/one_step.py fp16 > $O/ncu_loops.out 2>&1
ncu -i $O/ncu_loops.ncu-rep --page raw --csv > $O/ncu_full_dec_persist_loops.csv 2>/dev/null; el "ncu loops"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|bilstm_seq|gru_seq|bn_act|colreduce4|adam_clip|gemm_simt|sum_rows|im2col|col2im|relu_drop" -c 400 -o $O/ncu_step -f python profiles/tools/one_step.py fp16 > $O/ncu_step.out 2>&1
ncu -i $O/ncu_step.ncu-rep --page raw --csv > $O/ncu_full_train_step_kernels.csv 2>/dev/null; el "ncu step kernels"
ONE_STEP_INFER=1 T2V_COOP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stft_mel_fused|dec_persist_fwd_kernel<1" -c 2 -o $O/ncu_infer_stft -f python profiles/tools/one_step.py fp16 16 60 64 > $O/ncu_infer_stft.out 2>&1
ncu -i $O/ncu_infer_stft.ncu-rep --page raw --csv > $O/ncu_full_infer_stft.csv 2>/dev/null; el "ncu infer stft"
rm -f $O/*.ncu-rep
# ---- compute-sanitizer on a small step through every hand-synchronised kernel (_san build = longer bounded waits only)
for tool in memcheck racecheck synccheck; do
  T2V_LIB_SUFFIX=_san timeout 700 compute-sanitizer --tool $tool --error-exitcode 7 python profiles/tools/sanitize_small.py fp16 > $O/sanitizer_${tool}_fp16.log 2>&1
  echo "exit code $?" >> $O/sanitizer_${tool}_fp16.log; tail -3 $O/sanitizer_${tool}_fp16.log | cut -c1-160; el "sanitizer $tool"
done
T2V_LIB_SUFFIX=_san timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/tools/sanitize_small.py tf32 > $O/sanitizer_memcheck_tf32.log 2>&1
echo "exit code $?" >> $O/sanitizer_memcheck_tf32.log; tail -2 $O/sanitizer_memcheck_tf32.log | cut -c1-160; el "sanitizer memcheck tf32"
gzip -f $O/*.csv
ls -la $O | head -40
