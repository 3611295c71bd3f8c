"""One training process per GPU (reference multiproc.py:1-23): `python multiproc.py train.py ...` spawns
`train.py ... --n_gpus=N --group_name=... --rank=i` for every visible device and waits for them."""
import subprocess
import sys
import time

import torch

if __name__ == "__main__":
    argv = list(sys.argv)
    n = torch.cuda.device_count()
    argv += ["--n_gpus={}".format(n), "--group_name=group_{}".format(time.strftime("%Y_%m_%d-%H%M%S")), ""]
    workers = []
    for i in range(n):
        argv[-1] = "--rank={}".format(i)
        out = None if i == 0 else open("logs/GPU_{}.log".format(i), "w")
        workers.append(subprocess.Popen([sys.executable] + argv[1:], stdout=out))
    for p in workers:
        p.wait()
