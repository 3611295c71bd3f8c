# Soak test of the bench line: N back-to-back runs, counts failures (use before enabling anything that runs beside the persistent
# decoder kernels, e.g. T2V_POST_DW_BRANCH=1).  usage: [ENV=...] bash profiles/tools/soak_bench.sh [N]
N=${1:-10}
python -c "import torch; torch.zeros(1).cuda()"
ok=0; bad=0
for i in $(seq 1 $N); do
  if timeout 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > /tmp/soak_$i.json 2> /tmp/soak_$i.err; then
    ok=$((ok+1)); python -c "import json; d=json.load(open('/tmp/soak_$i.json')); print($i, d['ms_per_step'], d['decoder_step_backward']['value'])"
  else
    bad=$((bad+1)); echo "$i FAILED"; grep -m1 "Error" /tmp/soak_$i.err
  fi
done
echo "ok=$ok failed=$bad"
