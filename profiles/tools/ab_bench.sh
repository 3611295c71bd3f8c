# A/B of engine-level switches on the bench line (ms per train step)
python -c "import torch; torch.zeros(1).cuda()"
for v in 1 0; do echo "== T2V_POST_DW_BRANCH=$v"; T2V_POST_DW_BRANCH=$v timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -1
