mkdir -p gpurun_out/ev2
O=gpurun_out/ev2
for tool in memcheck racecheck synccheck; do
  T2V_LIB_SUFFIX=_san timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python profiles/tools/sanitize_small.py fp16 > $O/sanitizer_${tool}_fp16.log 2>&1
  echo "exit code $?" >> $O/sanitizer_${tool}_fp16.log
done
for prec in tf32 bf16; do
T2V_LIB_SUFFIX=_san timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/tools/sanitize_small.py $prec > $O/sanitizer_memcheck_$prec.log 2>&1
echo "exit code $?" >> $O/sanitizer_memcheck_$prec.log
done
timeout 900 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > $O/soak_200.json 2> $O/soak_200.err; echo "exit code $?" >> $O/soak_200.err
T2V_POST_DW_BRANCH=1 timeout 900 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > $O/soak_200_post_dw_branch.json 2> $O/soak_200_post_dw_branch.err; echo "exit code $?" >> $O/soak_200_post_dw_branch.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_bench.out 2>&1
for k in dec_persist_fwd dec_persist_bwd gemm_tc_mn16 bilstm_seq_fwd bilstm_seq_bwd colreduce4; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o $O/ncu_$k -f python profiles/tools/one_step.py fp16 > $O/ncu_$k.out 2>&1
  ncu -i $O/ncu_$k.ncu-rep --page raw --csv > $O/ncu_$k.csv 2>/dev/null
done
# the Postnet conv GEMMs (split3, N = 512, K = 2560 x 3 terms) are the large gemm_tc_kernel<256,...> launches after the decoder loop
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 16 -c 10 -o $O/ncu_gemm_tc_postnet -f python profiles/tools/one_step.py fp16 > $O/ncu_gemm_tc_postnet.out 2>&1
ncu -i $O/ncu_gemm_tc_postnet.ncu-rep --page raw --csv > $O/ncu_gemm_tc_postnet.csv 2>/dev/null
ONE_STEP_INFER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stft_mel_fused|dec_persist_fwd_kernel<1" -c 3 -o $O/ncu_infer_stft -f python profiles/tools/one_step.py fp16 16 60 64 > $O/ncu_infer_stft.out 2>&1
ncu -i $O/ncu_infer_stft.ncu-rep --page raw --csv > $O/ncu_infer_stft.csv 2>/dev/null
rm -f $O/*.ncu-rep
ls -la $O | head -40
