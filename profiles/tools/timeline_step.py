"""Scratch: device timeline (stream, start offset, duration) of a few consecutive decoder steps inside the replayed
train-step graph, from the CUPTI chrome trace.  usage: timeline_step.py B Ti To precision"""
import os, sys, json, tempfile
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tacotron2-vae_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import model as M
from hparams import create_hparams
from loss_function import Tacotron2Loss_VAE
from oracle import port

B, Ti, To, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
hp = create_hparams("anneal_function=constant")
m = M.Tacotron2(hp).cuda().train(); m.precision = prec
crit = Tacotron2Loss_VAE(hp)
x, y = m.parse_batch(port.synthetic_batch(B, Ti, To, seed=0))
def step():
    m.zero_grad(); out = m(x); loss, _, _, _ = crit(out, y, 0); loss.backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
def short(n):
    return n.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:34]
def dump(marker, title, count=26):
    idx = [i for i, e in enumerate(ev) if marker in e["name"]]
    if len(idx) < To // 2 + 2:
        print("no marker", marker); return
    i0 = idx[To // 2]
    t0 = ev[i0]["ts"]
    print("== %s: kernels around the middle step (us relative to the marker kernel start)" % title)
    for e in ev[i0 - 4:i0 + count]:
        print("  stream %-4s %9.2f  +%6.2f  %s" % (e["args"].get("stream"), e["ts"] - t0, e["dur"], short(e["name"])))
    steps = [ev[idx[k + 1]]["ts"] - ev[idx[k]]["ts"] for k in range(To // 4, 3 * To // 4)]
    print("  marker-to-marker period: mean %.2f us" % (sum(steps) / len(steps)))
dump("attn3_rowq_kernel", "forward loop", int(os.environ.get("TL_COUNT", 26)))
dump("attn2_bwd_dq_kernel", "backward loop", int(os.environ.get("TL_COUNT", 30)))
