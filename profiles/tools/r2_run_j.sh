mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2j_topo.txt 2>&1
for N in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2j_bench_${N}gpu.json 2> gpurun_out/r2j_bench_${N}gpu.err
head -c 330 gpurun_out/r2j_bench_${N}gpu.json; echo
done
grep -i "nvls\|nranks" gpurun_out/r2j_bench_8gpu.err | head -6
