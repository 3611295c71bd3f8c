#!/bin/bash
# Final-build check: GPU tests, smoke, default bench, and the ncu captures of the cluster kernels (plain launch: ncu rejects the
# cooperative + cluster launch with LaunchFailed, so these captures run with T2V_COOP=0; timings are never taken under ncu).
O=gpurun_out/fin; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "exit $?" >> $O/smoke.log
python bench.py > $O/bench.json 2> $O/bench.err; echo "exit $?" >> $O/bench.err
for k in dec_persist_fwd dec_persist_bwd gru_seq; do
  T2V_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o $O/ncu_$k -f python profiles/tools/one_step.py fp16 > $O/ncu_$k.out 2>&1
  ncu -i $O/ncu_$k.ncu-rep --page raw --csv > $O/ncu_$k.csv 2>/dev/null
done
ONE_STEP_INFER=1 T2V_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_persist_fwd -s 1 -c 1 -o $O/ncu_infer -f python profiles/tools/one_step.py fp16 16 60 64 > $O/ncu_infer.out 2>&1
ncu -i $O/ncu_infer.ncu-rep --page raw --csv > $O/ncu_infer.csv 2>/dev/null
ONE_STEP_INFER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_mel_fused -c 1 -o $O/ncu_stft -f python profiles/tools/one_step.py fp16 16 60 64 > $O/ncu_stft.out 2>&1
ncu -i $O/ncu_stft.ncu-rep --page raw --csv > $O/ncu_stft.csv 2>/dev/null
rm -f $O/*.ncu-rep
tail -3 $O/tests.log; tail -2 $O/smoke.log; cat $O/bench.json | cut -c1-400
