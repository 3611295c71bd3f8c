"""Run the reference's own train.py (unchanged) on top of this engine:
    python run_reference_train.py /path/to/reference/train.py -o out -l logs --hparams batch_size=64,...
This directory is put first on sys.path, so `from model import Tacotron2`, `from distributed import ...`,
`from loss_function import ...`, `from data_utils import ...`, `from logger import ...`, `from hparams import ...`
(train.py:8-19) resolve to the B200 engine while the training loop itself is the reference's file."""
import os
import runpy
import sys

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    script = sys.argv[1]
    sys.path.insert(0, here)
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")
