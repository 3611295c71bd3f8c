"""Small end-to-end problem through every hand-synchronised kernel (persistent decoder forward / backward / free-running loops, persistent
BiLSTM forward / backward, tcgen05 GEMMs) for compute-sanitizer runs:
    T2V_LIB_SUFFIX=_san compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/tools/sanitize_small.py [precision]
(the _san build only raises the bounded-wait limit: the tools slow the kernels down by orders of magnitude)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tacotron2-vae_b200"))
import torch
import model as M
from hparams import create_hparams
from loss_function import Tacotron2Loss_VAE
from oracle import port
from t2v import engine, infer

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B, Ti, To = 3, 10, 4
hp = create_hparams("anneal_function=constant")
m = M.Tacotron2(hp).cuda().train(); m.precision = prec
m._graph_cache = None
batch = port.synthetic_batch(B, Ti, To, seed=0)
x, y = m.parse_batch(batch)
out = m(x)
loss = Tacotron2Loss_VAE(hp)(out, y, 0)[0]
loss.backward()
torch.cuda.synchronize()
print("train step ok, loss %.5f" % loss.item(), flush=True)
with torch.no_grad():
    ops = engine.Ops(prec)
    sess = infer.DecoderSession(ops, m._state(), torch.randn(2, Ti, 512, device="cuda"), None, 4, training=False, seed=3)
    sess.run_free(4, 0.5, seed=3)
    torch.cuda.synchronize()
    mel, gate, align = sess.outputs(4)
    print("free-running decode ok, finite: %s" % bool(torch.isfinite(mel).all()), flush=True)
