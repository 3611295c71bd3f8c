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
torch.cuda.synchronize()
print("free-running decode ok", flush=True)
# kernels the small train step does not reach: the fused STFT / mel kernel, 256-row GEMM tiles (two accumulators), the row-reduction
# GEMMs with taps / column split / batches, the fp16 split product
from layers import TacotronSTFT
from t2v._lib import call as L
st = TacotronSTFT(1024, 256, 1024, 80, 16000, 0.0, 8000.0).cuda()
mel = st.mel_spectrogram(torch.rand(2, 4000, device="cuda") * 2 - 1)
dev = torch.device("cuda")
M_, N_, Ci = 256 * 150, 512, 64
X = torch.randn(M_ + 4, Ci, device=dev); W = torch.randn(N_, 5 * Ci, device=dev) * 0.05
Xh, Xl = (torch.empty(M_ + 4, Ci, device=dev, dtype=torch.int16) for _ in range(2))
Wh, Wl = (torch.empty(N_, 5 * Ci, device=dev, dtype=torch.int16) for _ in range(2))
L("t2v_split16", X, Xh, Xl, X.numel(), 1.0); L("t2v_split16", W, Wh, Wl, W.numel(), 16.0)
D = torch.zeros(M_, N_, device=dev)
L("t2v_gemm_tc_split3_16", Xh, Xl, Ci, M_ + 4, Ci, Wh, Wl, 5 * Ci, N_, 5 * Ci, D, N_, None, M_, N_, Ci, 5, 1, Ci, 0, 0, 1.0 / 16, 256)
A = torch.randn(700, 512, device=dev).half(); Bm = torch.randn(700, 512, device=dev).half()
D2 = torch.zeros(512, 5 * 512, device=dev)
L("t2v_gemm_tc_rowred16", A.view(torch.int16), 512, 512, 2, Bm.view(torch.int16), 512, 512, 0, D2, 5 * 512, 600, 3, 1, 1.0, None, 1, 5, None, 0, 0)
Da, Db = torch.zeros(512, 256, device=dev), torch.zeros(512, 256, device=dev)
L("t2v_gemm_tc_rowred16", A.view(torch.int16), 512, 512, 0, Bm.view(torch.int16), 512, 512, 0, Da, 256, 700, 2, 1, 1.0, None, 1, 1, Db, 256, 256)
Af = torch.rand(3, 100, 120, device=dev); Bf = torch.randn(100, 3, 512, device=dev); Dd = torch.empty(3, 120, 512, device=dev)
L("t2v_gemm_tc_rowred_batched", Af, 120, 120, 100, Bf, 3 * 512, 512, 512, Dd, 512, 120 * 512, 100, 3, 1.0)
torch.cuda.synchronize()
print("extra kernels ok", float(mel.mean()), float(D.abs().mean()), float(D2.abs().mean()), float(Dd.abs().mean()), flush=True)
