mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2i_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dp.py -q -s > gpurun_out/r2i_dp_test.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
tail -n 6 gpurun_out/r2i_dp_test.log; head -c 400 gpurun_out/r2i_bench_2gpu.json; grep -i "NVLS\|nranks\|Connected\|comm 0x" gpurun_out/r2i_bench_2gpu.err | head -12
