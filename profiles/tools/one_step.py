"""One C3-shaped train step (eager launches, no CUDA graph) + one free-running decode, for ncu captures:
    ncu --set full -k regex:<kernel> -c 2 python profiles/tools/one_step.py [precision] [B Ti To]"""
import os, sys
os.environ.setdefault("T2V_GRAPHS", "0")
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tacotron2-vae_b200"))
import torch
import model as M
from hparams import create_hparams
from loss_function import Tacotron2Loss_VAE
from oracle import port
from t2v import engine, infer, optim

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B, Ti, To = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (64, 120, 800)
hp = create_hparams("anneal_function=constant")
m = M.Tacotron2(hp).cuda().train(); m.precision = prec
crit = Tacotron2Loss_VAE(hp)
opt = optim.FusedAdamClip(m)
x, y = m.parse_batch(port.synthetic_batch(B, Ti, To, seed=0))
opt.zero_grad(); out = m(x); loss = crit(out, y, 0)[0]; loss.backward(); opt.step()
torch.cuda.synchronize()
if os.environ.get("ONE_STEP_INFER"):
    with torch.no_grad():
        sess = infer.DecoderSession(engine.Ops(prec), m._state(), torch.randn(16, 120, 512, device="cuda"), None, 200, training=False, seed=3)
        sess.run_free(200, 0.5, seed=3)
        torch.cuda.synchronize()
    from layers import TacotronSTFT
    st = TacotronSTFT(1024, 256, 1024, 80, 16000, 0.0, 8000.0).cuda()
    st.mel_spectrogram((torch.rand(16, 160000, device="cuda") * 2 - 1))
    torch.cuda.synchronize()
print("ok", float(loss))
