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
    for which, st in enumerate((_lib.T2VDecoderSeq, _lib.T2VDecoderBwd, _lib.T2VDecoderInfer)):
        assert ctypes.sizeof(st) == loaded.t2v_sizeof_decoder_structs(which), st.__name__


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


def test_persistent_loop_chunk_schedule_partitions_k():
    """Host-side restatement of the K-chunk schedule of the persistent decoder kernels (decoder_persist.cu: att_kofs / dec_kofs /
    att_chunk_infer / dec_chunk_infer; decoder_persist_bwd.cu: K slice 1024*rank + 32*j): over the 4 cluster ranks the chunks must
    cover every column of XA (1792) / XD (2560) / the 4096 gate rows exactly once, in every order the kernels use."""
    def att_kofs(j, r):
        return 64 * r + 32 * j if j < 2 else (768 + 256 * r + 32 * (j - 2) if j < 10 else 256 + 128 * r + 32 * (j - 10))

    def dec_kofs(j, r):
        return 256 * r + 32 * j if j < 8 else (1024 + 128 * r + 32 * (j - 8) if j < 12 else 1536 + 256 * r + 32 * (j - 12))

    def att_chunk_infer(j):
        return j + 2 if j < 12 else j - 12

    def dec_chunk_infer(j):
        return j + 12 if j < 8 else j - 8

    for kofs, nch, width in ((att_kofs, 14, 1792), (dec_kofs, 20, 2560)):
        cols = sorted(c for r in range(4) for j in range(nch) for c in range(kofs(j, r), kofs(j, r) + 32))
        assert cols == list(range(width))
    assert sorted(att_chunk_infer(j) for j in range(14)) == list(range(14))
    assert sorted(dec_chunk_infer(j) for j in range(20)) == list(range(20))
    # inference order: h_att (cols 768..), ctx (256..767), prenet (0..255) for the attention_rnn; h_dec, h_att, ctx for the decoder_rnn
    assert [att_kofs(att_chunk_infer(j), 0) for j in (0, 8, 12)] == [768, 256, 0]
    assert [dec_kofs(dec_chunk_infer(j), 0) for j in (0, 8, 16)] == [1536, 0, 1024]
    rows = sorted(k for r in range(4) for j in range(32) for k in range(1024 * r + 32 * j, 1024 * r + 32 * j + 32))
    assert rows == list(range(4096))
    # output-column ownership of the backward GEMMs: 32 clusters x 56 / 80 columns
    assert 32 * 56 == 1792 and 32 * 80 == 2560 and 56 % 4 == 0 and 80 % 4 == 0


def test_flat_copy_runs_groups_the_gradients_that_are_not_in_place():
    """t2v.functions.flat_copy_runs: slices of the flat gradient buffer that their producers already filled are skipped, the tensors in
    between are grouped into maximal runs (one concatenation each)"""
    from t2v.functions import flat_copy_runs
    names, sizes, base = list("abcdef"), [10, 4, 6, 100, 1, 3], 4096
    offs = [0, 10, 14, 20, 120, 121]
    ptr = lambda i: base + 4 * offs[i]
    # b and d were written in place, the rest elsewhere
    ptrs = [7, ptr(1), 9, ptr(3), 11, 13]
    assert flat_copy_runs(names, sizes, ptrs, base) == [(0, 10, ["a"]), (14, 20, ["c"]), (120, 124, ["e", "f"])]
    # nothing in place: one run over everything; everything in place: no copies
    assert flat_copy_runs(names, sizes, [0] * 6, base) == [(0, 124, names)]
    assert flat_copy_runs(names, sizes, [ptr(i) for i in range(6)], base) == []
    # a pointer that lies inside the buffer but at the wrong offset is not "in place"
    assert flat_copy_runs(["a", "b"], [10, 4], [ptr(1), ptr(0)], base) == [(0, 14, ["a", "b"])]


def test_grads_container_hands_out_flat_slices_only_when_they_fit():
    """t2v.engine.Grads.out: the destination of a large gradient is the flat-buffer slice when name, shape and contiguity match,
    a fresh tensor otherwise (CPU tensors: no kernel involved)"""
    import torch
    from t2v import engine
    flat = torch.ones(24)
    g = engine.Grads({"w": flat[:12].view(3, 4), "v": flat[12:].view(4, 3)})
    a = g.out("w", 3, 4, dev=torch.device("cpu"), zero=True)
    assert a.data_ptr() == flat.data_ptr() and float(flat[:12].abs().sum()) == 0.0 and float(flat[12:].sum()) == 12.0
    b = g.out("v", 3, 4, dev=torch.device("cpu"))                 # wrong shape -> not the slice
    assert b.data_ptr() != flat[12:].data_ptr() and tuple(b.shape) == (3, 4)
    c = g.out("missing", 2, 2, dev=torch.device("cpu"), zero=True)
    assert float(c.abs().sum()) == 0.0
    assert engine._gout({}, "w", 2, 2, dev=torch.device("cpu"), zero=True).shape == (2, 2)   # plain dict: always a fresh tensor


def test_pick_splits_stays_within_the_reduction_and_fills_waves():
    """t2v.engine.Ops.pick_splits: the split count of a row-reduction GEMM never exceeds the number of reduction iterations, leaves
    at least 16 iterations per split when there are enough output tiles, and is 1 when one wave of tiles already fills the GPU"""
    from t2v.engine import Ops
    for tiles in (1, 2, 4, 20, 37, 112, 148, 160, 300, 1000):
        for iters in (1, 3, 8, 25, 100, 800, 1600):
            sp = Ops.pick_splits(tiles, iters)
            assert 1 <= sp <= max(1, iters), (tiles, iters, sp)
            if tiles * 8 >= 148 and sp > 1:
                assert iters // sp >= 16, (tiles, iters, sp)
    assert Ops.pick_splits(148, 800) == 1            # one full wave: nothing to gain from splitting
    assert Ops.pick_splits(20, 800) > 1              # a Conv1d weight gradient with all taps in one launch: split to fill two waves
    assert Ops.pick_splits(4, 5) == 1                # too few iterations to split


def test_fused_adam_state_dict_follows_the_torch_layout():
    """t2v.optim.FusedAdamClip.state_dict / load_state_dict (train.py:92-119 checkpoints): torch.optim.Adam's layout, parameters that
    were never stepped have no state, a round trip through a second optimizer restores the moments (pure torch: runs on the CPU)"""
    import torch
    from t2v import optim
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3, bias=False))
    opt = optim.FusedAdamClip(net, lr=2e-3, weight_decay=1e-6, max_norm=1.0)
    sd0 = opt.state_dict()
    assert sd0["state"] == {} and sd0["param_groups"][0]["lr"] == 2e-3 and sd0["param_groups"][0]["params"] == [0, 1, 2]
    ref = torch.optim.Adam(net.parameters(), lr=2e-3, weight_decay=1e-6)
    assert set(ref.state_dict()["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "amsgrad", "params"}
    assert set(sd0["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "amsgrad", "params"}
    opt.m.copy_(torch.randn_like(opt.m)); opt.v.copy_(torch.rand_like(opt.v)); opt.step_count = 7
    sd = opt.state_dict()
    assert sorted(sd["state"]) == [0, 1, 2]
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 7.0
    assert tuple(sd["state"][0]["exp_avg"].shape) == (4, 5) and tuple(sd["state"][2]["exp_avg_sq"].shape) == (3, 4)
    net2 = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3, bias=False))
    opt2 = optim.FusedAdamClip(net2)
    opt2.load_state_dict(sd)
    assert opt2.step_count == 7 and torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v)
    assert opt2.param_groups[0]["lr"] == 2e-3
    # the parameters now live in one flat buffer and .grad views of the flat gradient buffer are adopted on step()
    assert net[0].weight.data_ptr() == opt.params.data_ptr()
