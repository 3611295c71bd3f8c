"""Device timeline of ONE whole train step (graph replay): every kernel with stream, start offset and duration, plus the
critical-path summary (time before / between / after the two persistent decoder loops).  CUPTI activity records via torch.profiler.
usage: timeline_full.py B Ti To precision [min_us]"""
import os, sys, json, tempfile
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tacotron2-vae_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import model as M
from hparams import create_hparams
from loss_function import Tacotron2Loss_VAE
from oracle import port
from t2v import optim

B, Ti, To, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
min_us = float(sys.argv[5]) if len(sys.argv) > 5 else 20.0
hp = create_hparams("anneal_function=constant")
m = M.Tacotron2(hp).cuda().train(); m.precision = prec
crit = Tacotron2Loss_VAE(hp)
opt = optim.FusedAdamClip(m)
x, y = m.parse_batch(port.synthetic_batch(B, Ti, To, seed=0))
def step():
    opt.zero_grad(); out = m(x); loss, _, _, _ = crit(out, y, 0); loss.backward(); opt.step()
for _ in range(4):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
def short(n):
    return n.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:44]
t0 = ev[0]["ts"]
end = max(e["ts"] + e["dur"] for e in ev)
print("step: %d device activities, %.2f ms from first start to last end" % (len(ev), (end - t0) / 1e3))
fw = [e for e in ev if "dec_persist_fwd" in e["name"]][0]
bw = [e for e in ev if "dec_persist_bwd" in e["name"]][0]
print("before forward loop %.2f ms | forward loop %.2f | between loops %.2f | backward loop %.2f | after backward loop %.2f" % (
    (fw["ts"] - t0) / 1e3, fw["dur"] / 1e3, (bw["ts"] - fw["ts"] - fw["dur"]) / 1e3, bw["dur"] / 1e3, (end - bw["ts"] - bw["dur"]) / 1e3))
print("activities >= %.0f us (offset ms, stream, duration us, name); '.' lines aggregate the shorter ones in between" % min_us)
small_n, small_t = 0, 0.0
for e in ev:
    if e["dur"] >= min_us:
        if small_n:
            print("      .  %d shorter activities, %.0f us in total" % (small_n, small_t)); small_n, small_t = 0, 0.0
        print("  %8.3f  s%-3s %9.1f  %s" % ((e["ts"] - t0) / 1e3, e["args"].get("stream"), e["dur"], short(e["name"])))
    else:
        small_n += 1; small_t += e["dur"]
if small_n:
    print("      .  %d shorter activities, %.0f us in total" % (small_n, small_t))
