# A/B of the Postnet-dW side branch on the bench line (ms per train step, or the failure)
python -c "import torch; torch.zeros(1).cuda()"
for v in 1 0 1 0; do
  echo "== T2V_POST_DW_BRANCH=$v"
  T2V_POST_DW_BRANCH=$v timeout 200 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > /tmp/ab.json 2> /tmp/ab.err
  echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("/tmp/ab.json")); print("ms_per_step", d["ms_per_step"], "bwd", d["decoder_step_backward"]["value"])
except Exception as e:
    print("FAILED:", open("/tmp/ab.err").read()[-1500:].split("Traceback")[-1][:600])
PY
done
