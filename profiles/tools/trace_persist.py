"""Scratch: in-kernel time stamps (SM clocks) of a few mid-sequence steps of the persistent decoder loop
(T2V_PERSIST_TRACE, decoder_persist.cu) + event-timed us/step of the loop alone.  usage: trace_persist.py [fp16|bf16|tf32] [B Ti To]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tacotron2-vae_b200"))
import torch
from oracle import port
from t2v import engine
from t2v._lib import call as L

PREC = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "fp16"
_a = [x for x in sys.argv[1:] if x.isdigit()]
B, Ti, To = (int(x) for x in _a[:3]) if len(_a) >= 3 else (64, 120, 400)
dev = torch.device("cuda")
P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
ops = engine.Ops(PREC)
mem = torch.randn(B, Ti, 512, device=dev) * 0.5
mel = torch.randn(B, 80, To, device=dev) * 2 - 5
in_len = torch.full((B,), Ti, device=dev, dtype=torch.long)
_, _, ctx = engine.decoder_forward(ops, P, mem, mel, in_len, True, None, None, 1, -float("inf"), dev)
torch.cuda.synchronize()
S = ctx["S"]
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L("t2v_decoder_fwd_steps", S, 0, To); e1.record(); torch.cuda.synchronize()
    print("loop %d: %.2f us/step" % (it, e0.elapsed_time(e1) * 1e3 / To), flush=True)
if not os.environ.get("NOTRACE"):
    os.environ["T2V_PERSIST_TRACE"] = "1"
    L("t2v_decoder_fwd_steps", S, 0, To)
    torch.cuda.synchronize()
