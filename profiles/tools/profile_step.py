"""Scratch profiler: per-kernel device time of full-size train steps via torch.profiler (CUPTI activity records; real
warm-cache durations, unlike ncu's serialised cold-cache replays).  usage: profile_step.py B Ti To precision [out.md]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tacotron2-vae_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import model as M
from hparams import create_hparams
from loss_function import Tacotron2Loss_VAE
from oracle import port
from t2v import optim

B, Ti, To, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
hp = create_hparams("anneal_function=constant")
m = M.Tacotron2(hp).cuda().train(); m.precision = prec
crit = Tacotron2Loss_VAE(hp)
opt = optim.FusedAdamClip(m)
batch = port.synthetic_batch(B, Ti, To, seed=0)
x, y = m.parse_batch(batch)

def step():
    opt.zero_grad(); out = m(x); loss, _, _, _ = crit(out, y, 0); loss.backward(); opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
N = 2
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        agg[name][0] += 1; agg[name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
if os.environ.get("T2V_PROFILE_LIST"):      # per-launch durations (us) of the kernels whose name contains the given substring
    pat = os.environ["T2V_PROFILE_LIST"]
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and pat in e.name), key=lambda e: e.time_range.start)
    print("per-launch us of *%s*: %s" % (pat, " ".join("%.0f" % (e.device_time if hasattr(e, "device_time") else e.cuda_time) for e in evs[:len(evs) // N])))
tot = sum(v[1] for v in agg.values())
lines = ["| kernel | launches/step | total us/step | avg us | share |", "|---|---:|---:|---:|---:|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    lines.append("| `%s` | %d | %.0f | %.2f | %.1f%% |" % (k[:64], n // N, t / N, t / n, 100 * t / tot))
lines.append("| **sum of kernel time** | %d | %.0f | | |" % (sum(v[0] for v in agg.values()) // N, tot / N))
print("\n".join(lines))
if len(sys.argv) > 5:
    open(sys.argv[5], "w").write("\n".join(lines) + "\n")
