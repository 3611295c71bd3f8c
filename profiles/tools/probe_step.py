import sys, time, os
R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tacotron2-vae_b200'))
import torch
import model as M
from hparams import create_hparams
from loss_function import Tacotron2Loss_VAE
from oracle import port
B,Ti,To = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prec = sys.argv[4]
hp = create_hparams("anneal_function=constant")
m = M.Tacotron2(hp).cuda().train(); m.precision = prec
crit = Tacotron2Loss_VAE(hp)
batch = port.synthetic_batch(B,Ti,To,seed=0)
x,y = m.parse_batch(batch)
for it in range(int(os.environ.get('PROBE_ITERS','2'))):
    t0=time.perf_counter()
    out = m(x); loss,_,_,_ = crit(out,y,0); loss.backward(); torch.cuda.synchronize()
    print("iter",it,"total ms",(time.perf_counter()-t0)*1e3, "loss", loss.item(), flush=True)
print("max mem GB", torch.cuda.max_memory_allocated()/1e9)
