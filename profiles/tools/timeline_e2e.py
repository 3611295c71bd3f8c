"""Device timeline of consecutive END-TO-END train steps (pinned host batch -> device, step, loss read back), as bench.py's e2e leg
runs them: memcpy activities, idle gaps on the device and the step period.  usage: timeline_e2e.py B Ti To precision"""
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
hp = create_hparams("anneal_function=constant")
dev = torch.device("cuda", 0)
m = M.Tacotron2(hp).cuda().train(); m.precision = prec
crit = Tacotron2Loss_VAE(hp)
opt = optim.FusedAdamClip(m)
host = [t.pin_memory() for t in port.synthetic_batch(B, Ti, To, seed=0)]
loss_host = torch.zeros(1).pin_memory()
def step(it):
    b = tuple(t.to(dev, non_blocking=True) for t in host)
    x = (b[0], b[1], b[2], Ti, b[4], b[5].float(), b[6].float()); y = (b[2], b[3])
    opt.zero_grad(); out = m(x); loss, _, _, _ = crit(out, y, it); loss.backward(); opt.step()
    loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
for it in range(4):
    step(it)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for it in range(4):
        step(it)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
fw = [e for e in ev if "dec_persist_fwd" in e["name"]]
print("forward-loop starts (ms):", ["%.2f" % ((e["ts"] - t0) / 1e3) for e in fw], " period(s):", ["%.2f" % ((b["ts"] - a["ts"]) / 1e3) for a, b in zip(fw, fw[1:])])
print("memcpy activities (offset ms, duration us, bytes, name):")
for e in ev:
    if e["cat"] == "gpu_memcpy" and (e["args"].get("bytes", 0) >= 1 << 16 or "HtoD" in e["name"] or "DtoH" in e["name"]):
        print("  %9.3f %9.1f %10d  %s" % ((e["ts"] - t0) / 1e3, e["dur"], e["args"].get("bytes", 0), e["name"][:40]))
print("idle gaps >= 50 us (offset ms, gap us, next activity):")
busy_end = ev[0]["ts"] + ev[0]["dur"]
for e in ev[1:]:
    if e["ts"] - busy_end >= 50:
        print("  %9.3f %9.1f  %s" % ((busy_end - t0) / 1e3, e["ts"] - busy_end, e["name"][:50]))
    busy_end = max(busy_end, e["ts"] + e["dur"])
# activities between the end of adam_clip and the next forward-loop: what the step boundary costs
ad = [e for e in ev if "adam_clip" in e["name"]]
if len(ad) >= 2 and len(fw) >= 3:
    a0 = ad[1]["ts"] + ad[1]["dur"]
    print("step boundary after the 2nd optimizer step: activities until the 3rd forward loop starts")
    for e in ev:
        if e["ts"] >= a0 and e["ts"] < fw[2]["ts"] and e["dur"] >= 15:
            print("  %9.3f s%-4s %9.1f  %s" % ((e["ts"] - t0) / 1e3, e["args"].get("stream"), e["dur"], e["name"].replace("(anonymous namespace)::", "").replace("void ", "")[:50]))
