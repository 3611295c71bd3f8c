mkdir -p gpurun_out/ev
O=gpurun_out/ev
# ---- 1. compute-sanitizer on every hand-synchronised kernel (small shapes, _san build = longer bounded waits only)
for tool in memcheck racecheck synccheck; do
  T2V_LIB_SUFFIX=_san timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python profiles/tools/sanitize_small.py fp16 > $O/sanitizer_${tool}_fp16.log 2>&1
  echo "exit code $?" >> $O/sanitizer_${tool}_fp16.log
done
T2V_LIB_SUFFIX=_san timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/tools/sanitize_small.py tf32 > $O/sanitizer_memcheck_tf32.log 2>&1
echo "exit code $?" >> $O/sanitizer_memcheck_tf32.log
# ---- 2. soak: 200 timed steps back to back, then the same with the Postnet-dW side branch beside the persistent backward kernel
timeout 900 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > $O/soak_200.json 2> $O/soak_200.err; echo "exit code $?" >> $O/soak_200.err
T2V_POST_DW_BRANCH=1 timeout 900 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > $O/soak_200_post_dw_branch.json 2> $O/soak_200_post_dw_branch.err; echo "exit code $?" >> $O/soak_200_post_dw_branch.err
# ---- 3. ncu launch list of the bench command
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_bench.out 2>&1
# ---- 4. ncu --set full per kernel family
for k in dec_persist_fwd dec_persist_bwd gemm_tc_kernel gemm_tc_mn bilstm_seq_fwd bilstm_seq_bwd bn_act_fwd colreduce adam_clip gemm_simt; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o $O/ncu_$k -f python profiles/tools/one_step.py fp16 > $O/ncu_$k.out 2>&1
  ncu -i $O/ncu_$k.ncu-rep --page raw --csv > $O/ncu_$k.csv 2>/dev/null
done
ls -la $O | head -60
