"""Scratch: in-kernel time stamps of the persistent decoder BACKWARD loop (T2V_PERSIST_TRACE, decoder_persist_bwd.cu) and
event-timed us/step of the reverse loop alone.  usage: trace_persist_bwd.py [B Ti To]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tacotron2-vae_b200"))
import torch
from oracle import port
from t2v import engine

B, Ti, To = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 120, 400)
dev = torch.device("cuda")
P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
ops = engine.Ops(sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "fp16")
mem = torch.randn(B, Ti, 512, device=dev) * 0.5
mel = torch.randn(B, 80, To, device=dev) * 2 - 5
in_len = torch.full((B,), Ti, device=dev, dtype=torch.long)
dO = torch.randn(To * B, 84, device=dev) * 0.1
O, _, ctx = engine.decoder_forward(ops, P, mem, mel, in_len, True, None, None, 1, -float("inf"), dev)
torch.cuda.synchronize()
# time decoder_backward's loop by stage tracing: run the whole backward, the loop dominates; report via CUDA events around it
import t2v.engine as E
orig = E.L
times = []
def timed_L(name, *a):
    if name == "t2v_decoder_bwd_steps":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(name, *a); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3 / To)
        return r
    return orig(name, *a)
E.L = timed_L
for it in range(3):
    grads = {}
    dmem, br = engine.decoder_backward(ops, P, dO, ctx, dev, grads); br.join(); torch.cuda.synchronize()
    print("bwd loop %d: %.2f us/step" % (it, times[-1]), flush=True)
if not os.environ.get("NOTRACE"):
    os.environ["T2V_PERSIST_TRACE"] = "1"
    grads = {}
    dmem, br = engine.decoder_backward(ops, P, dO, ctx, dev, grads); br.join(); torch.cuda.synchronize()
