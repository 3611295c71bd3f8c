mkdir -p gpurun_out
for cfg in "1 1 1" "1 0 1" "1 2 1" "1 1 0" "0 0 1"; do set -- $cfg; echo "== WA=$1 WD=$2 MEM=$3"; NOTRACE=1 T2V_PERSIST_WA_HINT=$1 T2V_PERSIST_WD_HINT=$2 T2V_PERSIST_MEM_HINT=$3 timeout 100 python profiles/tools/trace_persist.py 2>&1 | tail -1; done > gpurun_out/p6_hints.txt 2>&1
cat gpurun_out/p6_hints.txt
NOTRACE=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:dec_persist --launch-skip 2 --launch-count 1 --csv --log-file gpurun_out/p6_ncu_dram.csv python profiles/tools/trace_persist.py 64 120 200 > gpurun_out/p6_ncu.log 2>&1
tail -8 gpurun_out/p6_ncu_dram.csv
