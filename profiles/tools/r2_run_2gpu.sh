#!/bin/bash
# 2-GPU evidence: NCCL DP parity test + weak-scaling bench line (launched like the driver does), NCCL init log kept
O=gpurun_out/${1:-g2}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > $O/tests_dp.log 2>&1; echo "exit $?" >> $O/tests_dp.log; tail -3 $O/tests_dp.log
T2V_NCCL_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "exit $?" >> $O/bench_2gpu.err
cut -c1-400 $O/bench_2gpu.json
grep -h "NCCL INFO.*\(Connected\|Channel\|NVLS\|comm 0x\|nranks\)" gpurun_out/nccl_init_* 2>/dev/null | head -12 > $O/nccl_init_excerpt.txt; wc -l $O/nccl_init_excerpt.txt
