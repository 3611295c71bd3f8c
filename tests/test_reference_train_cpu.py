"""The reference's OWN train.py, unmodified, on top of this engine (north star: "the reference's train loop runs unchanged").

Build container only (marker `reference`: needs /root/reference; there is no GPU here).  train.py is executed with runpy exactly as
tacotron2-vae_b200/run_reference_train.py does -- this package first on sys.path, so `from model import Tacotron2`, `from distributed
import apply_gradient_allreduce`, `from data_utils import ...`, `from loss_function import ...`, `from logger import ...`,
`from hparams import create_hparams` (train.py:8-19) resolve to the B200 engine.  It runs argument parsing, hparams parsing, model
construction (`Tacotron2(hparams).cuda()` -- .cuda() is an identity here), Adam construction over model.parameters(), the logger, the
DataLoader (TextMelLoader with load_mel_from_disk / TextMelCollate, one worker), `model.train()`, `model.zero_grad()`,
`model.parse_batch(batch)` and reaches `model(x)`, where the engine refuses CPU tensors: that RuntimeError is the expected end."""
import os
import runpy
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tacotron2-vae_b200")

pytestmark = pytest.mark.reference


def _corpus(tmp):
    rng = np.random.RandomState(0)
    lines = []
    for i, txt in enumerate(["감정있는 한국어 목소리 생성", "안녕하세요 반갑습니다", "오늘 날씨가 좋네요", "테스트 문장입니다"]):
        path = os.path.join(tmp, "m%d.npy" % i)
        np.save(path, rng.randn(80, 20 + 3 * i).astype(np.float32))
        lines.append("%s|%s|0|%d" % (path, txt, i % 4))
    fl = os.path.join(tmp, "files.txt")
    open(fl, "w", encoding="utf-8").write("\n".join(lines))
    return fl


def test_reference_train_py_runs_unchanged_up_to_the_first_forward(tmp_path, monkeypatch):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check (on a GPU box tests/test_gpu_trainloop.py runs the same call sequence to completion)")
    fl = _corpus(str(tmp_path))
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "manual_seed", lambda *a, **k: None)
    saved_path, saved_argv = list(sys.path), list(sys.argv)
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                  if k in ("model", "hparams", "data_utils", "loss_function", "logger", "distributed", "fp16_optimizer", "layers",
                           "modules", "utils", "text", "plotting_utils", "CoordConv", "stft", "audio_processing")}
    seen = {}
    try:
        sys.path.insert(0, PKG)
        sys.argv = ["/root/reference/train.py", "-o", str(tmp_path / "out"), "-l", "logs",
                    "--hparams=batch_size=2,load_mel_from_disk=True,training_files=%s,validation_files=%s,anneal_function=constant" % (fl, fl)]
        import model as engine_model          # the module train.py's `from model import Tacotron2` will get
        orig_forward = engine_model.Tacotron2.forward

        def spy(self, inputs):
            seen["model_class_file"] = sys.modules[type(self).__module__].__file__
            seen["training"] = self.training
            seen["n_params"] = sum(p.numel() for p in self.parameters())
            seen["x_shapes"] = [tuple(t.shape) if torch.is_tensor(t) else t for t in inputs]
            return orig_forward(self, inputs)
        monkeypatch.setattr(engine_model.Tacotron2, "forward", spy)
        with pytest.raises(RuntimeError, match="needs CUDA tensors"):
            runpy.run_path("/root/reference/train.py", run_name="__main__")
    finally:
        sys.path[:] = saved_path
        sys.argv[:] = saved_argv
        for k in list(sys.modules):
            if k in saved_mods:
                sys.modules.pop(k)
        sys.modules.update(saved_mods)
    assert seen["model_class_file"].startswith(PKG), seen
    assert seen["training"] is True and seen["n_params"] == 28875057
    assert seen["x_shapes"][0][0] == 2 and seen["x_shapes"][2][:2] == (2, 80)      # (text [B,Ti], in_len, mel [B,80,To], ...)
    assert os.path.isdir(str(tmp_path / "out"))                                     # prepare_directories_and_logger ran
