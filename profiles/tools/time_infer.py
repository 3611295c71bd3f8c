"""Scratch: us per free-running decoder step (config 5: B=16, Ti=120, 1000 steps) through bench.inference_decoder_step.
usage: time_infer.py [B Ti n]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tacotron2-vae_b200"))
import torch
import bench
import model as M
from hparams import create_hparams
B, Ti, n = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 120, 1000)
m = M.Tacotron2(create_hparams("anneal_function=constant")).cuda().eval()
for mode in ("1", "0"):
    os.environ["T2V_PERSIST"] = mode
    r = bench.inference_decoder_step(m, torch.device("cuda"), "tf32", B=B, Ti=Ti, n=n)
    print("T2V_PERSIST=%s: %.2f us/step (B=%d, Ti=%d, %d steps, finite=%s)" % (mode, r["value"], B, Ti, n, r["finite"]), flush=True)
if os.environ.get("TRACE"):
    from oracle import port
    from t2v import engine, infer
    os.environ["T2V_PERSIST"] = "1"
    os.environ["T2V_PERSIST_TRACE"] = "1"
    P = m._state()
    sess = infer.DecoderSession(engine.Ops("tf32"), P, torch.randn(B, Ti, 512, device="cuda"), None, 300, training=False, seed=7)
    sess.run_free(300, 2.0, seed=7)
    torch.cuda.synchronize()
