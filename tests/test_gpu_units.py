"""Kernel-level checks on the B200 (all through the C ABI in include/t2v_b200.h via t2v._lib)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from t2v import _lib
    _lib.lib()
    return _lib.call


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_gemm_f32_strided(L):
    torch.manual_seed(0)
    dev = "cuda"
    for (M, N, K) in [(64, 81, 1536), (300, 130, 77), (5, 4096, 1792)]:
        A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); bias = torch.randn(N, device=dev)
        C = torch.empty(M, N, device=dev)
        L("t2v_gemm_f32", A, K, 1, B, K, 1, C, N, M, N, K, 1.0, 0.0, bias, 1, 0, 0, 0)
        ref = (A.double() @ B.double().t() + bias.double()).float()
        assert _rel(C, ref) < 1e-5
        # transposed access of both operands + accumulate
        At = A.t().contiguous(); Bt = B.t().contiguous()
        C2 = ref.clone()
        L("t2v_gemm_f32", At, 1, M, Bt, 1, N, C2, N, M, N, K, 0.5, 1.0, None, 1, 0, 0, 0)
        ref2 = ref + 0.5 * (A.double() @ B.double().t()).float()
        assert _rel(C2, ref2) < 1e-5


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 128, 128), (256, 64, 96, 64), (1000, 300, 1792, 128), (64, 4096, 2560, 128),
                                      (515, 81, 1024, 128), (700, 512, 80, 256)])
def test_gemm_tc_tf32_plain(L, M, N, K, bn):
    torch.manual_seed(1)
    dev = "cuda"
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); bias = torch.randn(N, device=dev)
    D = torch.full((M, N), float("nan"), device=dev)
    L("t2v_gemm_tc", A, K, M, K, B, K, N, K, D, N, bias, M, N, K, 1, 0, 0, 0, 0, 4, 1, 0, 0, 1.0, bn)
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().t() + bias.double()).float()
    err = _rel(D, ref)
    assert err < 2e-3, err          # tf32 operands (10-bit mantissa), fp32 accumulate


def test_gemm_tc_split_k_and_atomic(L):
    torch.manual_seed(2)
    dev = "cuda"
    M, N, K = 64, 4096, 1792
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    parts = torch.empty(4, M, N, device=dev)
    L("t2v_gemm_tc", A, K, M, K, B, K, N, K, parts, N, None, M, N, K, 1, 0, 0, 0, 0, 4, 4, M * N, 0, 1.0, 128)
    ref = (A.double() @ B.double().t()).float()
    assert _rel(parts.sum(0), ref) < 2e-3
    D = torch.zeros(M, N, device=dev)
    L("t2v_gemm_tc", A, K, M, K, B, K, N, K, D, N, None, M, N, K, 1, 0, 0, 0, 0, 4, 8, 0, 1, 1.0, 128)
    assert _rel(D, ref) < 2e-3


def test_gemm_tc_taps_is_conv1d(L):
    """k=5/p=2 Conv1d over padded channels-last rows == F.conv1d (reference model.py:105-177 layers)."""
    torch.manual_seed(3)
    dev = "cuda"
    for (B, T, Ci, Co) in [(3, 37, 80, 512), (2, 50, 512, 80), (4, 30, 512, 512)]:
        x = torch.randn(B, Ci, T, device=dev); w = torch.randn(Co, Ci, 5, device=dev) * 0.05; b = torch.randn(Co, device=dev)
        Tp = T + 4
        X = torch.zeros(B * Tp, Ci, device=dev)
        L("t2v_bct_to_padded", x, X, B, Ci, T, 0.0)
        Wk = torch.empty(Co, 5 * Ci, device=dev)
        L("t2v_conv1d_pack", w, Wk, Co, Ci, 5, 0, 0)
        ref = torch.nn.functional.conv1d(x.cpu(), w.cpu(), b.cpu(), padding=2)
        for mode in ("f32", "tc"):
            Y = torch.zeros(B * Tp, Co, device=dev)
            M = B * Tp - 4
            if mode == "tc":
                L("t2v_gemm_tc", X, Ci, B * Tp, Ci, Wk, 5 * Ci, Co, 5 * Ci, Y.data_ptr() + 8 * Co, Co, b, M, Co, Ci, 5, 1, Ci, 0, 0,
                  4, 1, 0, 0, 1.0, 128)
            else:
                L("t2v_gemm_f32", X, Ci, 1, Wk, 5 * Ci, 1, Y.data_ptr() + 8 * Co, Co, M, Co, 5 * Ci, 1.0, 0.0, b, 1, 0, 0, 0)
            out = torch.empty(B, Co, T, device=dev)
            L("t2v_padded_to_bct", Y, None, out, B, Co, T, None, 0.0)
            err = _rel(out.cpu(), ref)
            assert err < (2e-3 if mode == "tc" else 1e-5), (mode, err)


def test_bn_act_dropout_matches_torch(L):
    torch.manual_seed(4)
    dev = "cuda"
    B, C, T = 3, 512, 21
    x = torch.randn(B, C, T, device=dev) * 2 + 0.5
    mask = (torch.rand(B, C, T, device=dev) >= 0.5).float()
    gamma = torch.rand(C, device=dev) + 0.5; beta = torch.randn(C, device=dev) * 0.1
    Tp = T + 4
    Y = torch.zeros(B * Tp, C, device=dev)
    L("t2v_bct_to_padded", x, Y, B, C, T, 0.0)
    sums = torch.zeros(2, C, device=dev, dtype=torch.float64)
    L("t2v_col_stats", Y, B * Tp, C, Tp, 2, 2 + T, 0, sums[0], sums[1])
    mean = torch.empty(C, device=dev); invstd = torch.empty(C, device=dev)
    rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev); nbt = torch.zeros((), device=dev, dtype=torch.long)
    L("t2v_bn_finalize", sums[0], sums[1], float(B * T), C, 1e-5, 0.1, mean, invstd, rm, rv, nbt)
    out = torch.empty_like(Y)
    L("t2v_bn_act_fwd", Y, out, None, B * Tp, C, Tp, 2, 2 + T, mean, invstd, gamma, beta, 2, mask, 0, 0, 0.5, T, 0, None, None)
    o = torch.empty(B, C, T, device=dev)
    L("t2v_padded_to_bct", out, None, o, B, C, T, None, 0.0)
    xr = x.cpu().double().requires_grad_(True)
    bn = torch.nn.functional.batch_norm(xr, None, None, gamma.cpu().double(), beta.cpu().double(), True, 0.1, 1e-5)
    ref = torch.tanh(bn) * mask.cpu().double() * 2.0
    assert _rel(o.cpu().double(), ref.detach()) < 1e-5
    assert int(nbt) == 1
    assert torch.allclose(rm.cpu(), 0.1 * x.cpu().mean((0, 2)), atol=1e-5)
    assert torch.allclose(rv.cpu(), 0.9 + 0.1 * x.cpu().var((0, 2), unbiased=True), rtol=1e-4, atol=1e-5)
    # backward
    g = torch.randn(B, C, T, device=dev)
    ref.backward(g.cpu().double())
    G = torch.zeros(B * Tp, C, device=dev)
    L("t2v_bct_to_padded", g, G, B, C, T, 0.0)
    s2 = torch.zeros(2, C, device=dev, dtype=torch.float64)
    L("t2v_bn_act_bwd_reduce", G, Y, B * Tp, C, Tp, 2, 2 + T, mean, invstd, gamma, beta, 2, mask, 0, 0, 0.5, T, s2[0], s2[1])
    dY = torch.empty_like(Y)
    L("t2v_bn_act_bwd_apply", G, Y, dY, B * Tp, C, Tp, 2, 2 + T, mean, invstd, gamma, beta, 2, mask, 0, 0, 0.5, T, s2[0], s2[1],
      float(B * T), 1, 0)
    dx = torch.empty(B, C, T, device=dev)
    L("t2v_padded_to_bct", dY, None, dx, B, C, T, None, 0.0)
    assert _rel(dx.cpu().double(), xr.grad) < 1e-4
    assert float(dY.view(B, Tp, C)[:, :2].abs().sum()) == 0.0


def test_embedding_bit_exact(L):
    dev = "cuda"
    torch.manual_seed(5)
    ids = torch.randint(0, 80, (4, 33), device=dev)
    table = torch.randn(80, 512, device=dev)
    out = torch.zeros(4 * 37, 512, device=dev)
    L("t2v_embedding_fwd", ids, table, out, 4, 33, 512, 80, 0)
    ref = table.cpu()[ids.cpu()]
    assert torch.equal(out.view(4, 37, 512)[:, 2:35].cpu(), ref)


def test_attention_step_matches_port(L):
    from oracle import port
    torch.manual_seed(6)
    dev = "cuda"
    B, Ti = 5, 47
    P = port.init_params(11)
    pre = "decoder.attention_layer."
    mem = torch.randn(B, Ti, 512); q = torch.randn(B, 128)
    wprev = torch.softmax(torch.randn(B, Ti), 1); cum = torch.rand(B, Ti)
    lens = torch.tensor([47, 40, 33, 20, 9])
    pmem = mem @ P[pre + "memory_layer.linear_layer.weight"].t()
    wcat = torch.stack((wprev, cum), 1)
    loc = torch.nn.functional.conv1d(wcat, P[pre + "location_layer.location_conv.conv.weight"], padding=15)
    loc = loc.transpose(1, 2) @ P[pre + "location_layer.location_dense.linear_layer.weight"].t()
    a_ref = torch.tanh(q.unsqueeze(1) + loc + pmem)
    e = (a_ref @ P[pre + "v.linear_layer.weight"].t()).squeeze(-1)
    e = e.masked_fill(~(torch.arange(Ti)[None] < lens[:, None]), -float("inf"))
    w_ref = torch.softmax(e, 1)
    ctx_ref = torch.bmm(w_ref.unsqueeze(1), mem).squeeze(1)
    g = lambda t: t.to(dev).contiguous()
    w_out = torch.empty(B, Ti, device=dev); cum_out = torch.empty(B, Ti, device=dev)
    c1 = torch.empty(B, 512, device=dev); c2 = torch.empty(B, 600, device=dev); a_save = torch.empty(B, Ti, 128, device=dev)
    L("t2v_attn_step_fwd", g(q), 1, 0, g(wprev), Ti, g(cum), cum_out, g(pmem), g(mem),
      g(P[pre + "location_layer.location_conv.conv.weight"]), g(P[pre + "location_layer.location_dense.linear_layer.weight"]),
      g(P[pre + "v.linear_layer.weight"]), g(lens), -float("inf"), w_out, Ti, c1, 512, c2, 600, a_save, B, Ti, 0)
    assert torch.allclose(w_out.cpu(), w_ref, atol=2e-6)
    assert torch.allclose(c1.cpu(), ctx_ref, atol=1e-5)
    assert torch.equal(c2[:, :512], c1)
    assert torch.allclose(cum_out.cpu(), cum + w_ref, atol=2e-6)
    assert torch.allclose(a_save.cpu(), a_ref, atol=1e-5)


def test_loss_and_adam_match_torch(L):
    torch.manual_seed(7)
    dev = "cuda"
    B, To = 3, 19
    mel = torch.randn(B, 80, To); post = torch.randn(B, 80, To); tgt = torch.randn(B, 80, To)
    gate = torch.randn(B, To) * 3; gt = (torch.rand(B, To) > 0.7).float(); mu = torch.randn(B, 32); lv = torch.randn(B, 32) * 0.3
    acc = torch.zeros(4, device=dev, dtype=torch.float64); out = torch.empty(3, device=dev)
    g = lambda t: t.to(dev)
    L("t2v_loss_fwd", g(mel), g(post), g(tgt), mel.numel(), g(gate), g(gt), gate.numel(), g(mu), g(lv), mu.numel(), 0.001, acc, out)
    recon = ((mel - tgt) ** 2).mean() + ((post - tgt) ** 2).mean() + torch.nn.functional.binary_cross_entropy_with_logits(gate, gt)
    kl = -0.5 * torch.sum(1 + lv - mu ** 2 - lv.exp())
    assert torch.allclose(out.cpu(), torch.stack([recon + 0.001 * kl, recon, kl]), rtol=1e-5)
    # fused clip + Adam vs torch
    n = 10007
    p = torch.randn(n); gr = torch.randn(n) * 0.1
    pt = p.clone().requires_grad_(True); pt.grad = gr.clone()
    opt = torch.optim.Adam([pt], lr=1e-3, weight_decay=1e-6)
    pd, gd = g(p.clone()), g(gr.clone()); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    ss = torch.zeros(1, device=dev, dtype=torch.float64); norm = torch.zeros(1, device=dev)
    for step in (1, 2, 3):
        tn = torch.nn.utils.clip_grad_norm_([pt], 1.0)
        opt.step()
        L("t2v_grad_sumsq", gd, n, 1.0, ss)
        L("t2v_adam_clip_step", pd, gd, m, v, n, ss, 1.0, 1.0, 1e-3, 0.9, 0.999, 1e-8, 1e-6, step, norm)
        if step == 1:
            assert abs(float(norm) - float(tn)) < 1e-4 * float(tn)
        pt.grad = gr.clone(); gd.copy_(g(gr))
    assert torch.allclose(pd.cpu(), pt.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("Ti", [77, 120, 200, 300])
def test_attention_v2_matches_v1(L, Ti):
    """The SM-parallel attention kernels (attention2.cu) against the one-CTA-per-utterance reference kernels
    (attention.cu, themselves checked against the oracle above): forward and the full backward incl. the scattered
    adjoint conv and the per-slot weight-gradient partials."""
    from t2v import _lib
    torch.manual_seed(8)
    dev = "cuda"
    B = 5
    nck = int(_lib.lib().t2v_attn2_chunks(Ti))
    r = lambda *s: torch.randn(*s, device=dev)
    q = r(B, 128); wprev = torch.softmax(r(B, Ti), 1); cum = torch.rand(B, Ti, device=dev)
    pmem = r(B, Ti, 128); mem = r(B, Ti, 512)
    wconv = r(32, 2, 31) * 0.2; wloc = r(128, 32) * 0.2; v = r(128) * 0.3
    lens = torch.tensor([Ti, Ti - 7, Ti // 2, 33, 9], device=dev)
    wconvT = wconv.reshape(32, 62).t().contiguous()
    out = {}
    for ver in (1, 2, 3):
        w = torch.zeros(B, Ti, device=dev); cumo = torch.zeros(B, Ti, device=dev)
        c1 = torch.zeros(B, 512, device=dev); c2 = torch.zeros(B, 512, device=dev); a = torch.zeros(B, Ti, 128, device=dev)
        if ver == 1:
            L("t2v_attn_step_fwd", q, 1, 0, wprev, Ti, cum, cumo, pmem, mem, wconv, wloc, v, lens, -float("inf"), w, Ti, c1, 512,
              c2, 512, a, B, Ti, 0)
        elif ver == 2:
            e = torch.empty(B, Ti, device=dev)
            L("t2v_attn2_fwd", q, 1, 0, wprev, Ti, cum, cumo, pmem, mem, wconvT, wloc, v, lens, -float("inf"), e, w, Ti, c1, 512,
              c2, 512, a, B, Ti, 0)
        else:       # location term precomputed off the recurrence + the query-dependent row kernel
            pre = torch.empty(B, Ti, 128, device=dev)
            L("t2v_attn3_loc_fwd", wprev, Ti, cum, pmem, wconvT, wloc, pre, B, Ti)
            L("t2v_attn3_row_fwd", q, 1, 0, pre, cum, cumo, mem, v, lens, -float("inf"), w, Ti, c1, 512, c2, 512, a, B, Ti, 0)
        out[ver] = (w, cumo, c1, c2, a)
    for ver in (2, 3):
        for x, y in zip(out[1], out[ver]):
            assert torch.allclose(x, y, atol=2e-6, rtol=1e-5)
    w, _, _, _, a_save = out[1]
    # backward
    d1, d2, d3 = r(B, 512), r(B, 512), r(B, 512)
    dw_in = r(B, Ti) * 0.1; gc = r(B, Ti) * 0.1
    # v1
    dmem1 = torch.zeros(B, Ti, 512, device=dev); dpm1 = torch.zeros(B, Ti, 128, device=dev); dq1 = torch.zeros(B, 128, device=dev)
    dv1 = torch.zeros(B, 128, device=dev); dwl1 = torch.zeros(B, 128 * 32, device=dev); dwc1 = torch.zeros(B, 32 * 62, device=dev)
    dwo1 = torch.zeros(B, Ti, device=dev); gc1 = gc.clone()
    L("t2v_attn_step_bwd", d1, 512, d2, 512, d3, 512, dw_in, dwo1, gc1, w, Ti, wprev, Ti, cum, a_save, mem, wconv, wloc, v, lens,
      dmem1, dpm1, dq1, dv1, dwl1, dwc1, B, Ti, 0)
    # v2
    dctx = torch.empty(B, 512, device=dev); dwp = torch.empty(4, B, Ti, device=dev)
    dpm2 = torch.zeros(B, Ti, 128, device=dev); dq2 = torch.zeros(B, 128, device=dev)
    dv2 = torch.zeros(B * nck, 128, device=dev); dwl2 = torch.zeros(B * nck, 128 * 32, device=dev)
    dwc2 = torch.zeros(B * nck, 32 * 62, device=dev)
    dwo2 = torch.full((B, Ti), 7.0, device=dev); gcn = torch.full((B, Ti), -3.0, device=dev)
    de = torch.empty(B, Ti, device=dev)
    L("t2v_attn2_bwd", d1, 512, d2, 512, d3, 512, dctx, dw_in, dwo2, gc, gcn, dwp, de, w, Ti, wprev, Ti, cum, a_save, mem, wconvT,
      wloc, v, lens, dpm2, dq2, dv2, dwl2, dwc2, B, Ti)
    tol = dict(atol=2e-5, rtol=1e-4)
    assert torch.allclose(dctx, d1 + d2 + d3, **tol)
    assert torch.allclose(dmem1, w.unsqueeze(2) * dctx.unsqueeze(1), **tol)          # what the batched GEMM will produce
    assert torch.allclose(dpm1, dpm2, **tol) and torch.allclose(dq1, dq2, **tol)
    assert torch.allclose(dwo1, dwo2, **tol) and torch.allclose(gc1, gcn, **tol)
    assert torch.allclose(dv1.sum(0), dv2.sum(0), **tol)
    assert torch.allclose(dwl1.sum(0), dwl2.sum(0), atol=1e-4, rtol=1e-4)
    assert torch.allclose(dwc1.sum(0).view(32, 62), dwc2.sum(0).view(62, 32).t(), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("M", [64, 37, 16])
def test_gemm_tc_m64_tile(L, M):
    """Small-M GEMMs (the decoder-step shapes: M = batch <= 64) run on the UMMA M=64 tile."""
    torch.manual_seed(9)
    dev = "cuda"
    N, K = 4096, 1792
    A = torch.randn(M + 70, K, device=dev)[:M]          # rows beyond M exist in memory (like the next time step's rows)
    Bw = torch.randn(N, K, device=dev)
    parts = torch.full((4, M, N), float("nan"), device=dev)
    L("t2v_gemm_tc", A, K, M + 70, K, Bw, K, N, K, parts, N, None, M, N, K, 1, 0, 0, 0, 0, 4, 4, M * N, 0, 1.0, 128)
    ref = (A.double() @ Bw.double().t()).float()
    assert _rel(parts.sum(0), ref) < 2e-3


@pytest.mark.parametrize("B,Ti,To,training", [(5, 23, 12, True), (64, 120, 24, True), (16, 128, 9, False), (1, 7, 5, True)])
def test_persistent_decoder_loop_matches_per_step_launches(L, B, Ti, To, training):
    """decoder_persist.cu (one resident kernel for the whole teacher-forced loop) against the per-step launch sequence of
    decoder.cu on the same inputs / same counter RNG: mel+gate rows, alignments and every saved activation.  Both paths
    round the same operands to tf32 and accumulate in fp32; only the split-K summation order differs: tolerance 2e-4 of
    the buffer's max, and one tf32 ulp (2^-10) for XA / XD, whose h / ctx columns are stored rounded to the tf32 grid."""
    import os
    from oracle import port
    from t2v import engine
    dev = torch.device("cuda")
    P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
    ops = engine.Ops("tf32")
    g = torch.Generator().manual_seed(B * 1000 + Ti)
    memory = (torch.randn(B, Ti, 512, generator=g) * 0.5).to(dev)
    mel = (torch.randn(B, 80, To, generator=g) * 2 - 5).to(dev)
    in_len = torch.randint(max(1, Ti // 2), Ti + 1, (B,), generator=g).sort(descending=True)[0]
    in_len[0] = Ti
    in_len = in_len.to(dev)
    outs = {}
    for mode in ("0", "1"):
        os.environ["T2V_PERSIST"] = mode
        try:
            O, align, ctx = engine.decoder_forward(ops, P, memory, mel, in_len, training, None, None, 77, -float("inf"), dev)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("T2V_PERSIST", None)
        buf = ctx["buf"]
        outs[mode] = dict(O=O.clone(), align=align.clone(), **{k: buf[k].clone() for k in
                          ("XA", "XD", "CA", "CD", "CUM", "GA", "GD", "CPA", "CPD", "ASAVE")})
    for k in outs["0"]:
        a, b = outs["1"][k], outs["0"][k]
        assert torch.isfinite(a).all(), k
        err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        print("persistent vs per-step %-6s max-rel %.3e" % (k, err))
        # (O: the persistent kernel also emits the low parts of h_dec / ctx for the split projection, the per-step launches do not,
        #  so the two mel/gate rows differ by the tf32 rounding of the projection operand: 1e-3 of the max)
        assert err <= (1.5e-3 if k in ("XA", "XD") else (1e-3 if k == "O" else 2e-4)), (k, err)


@pytest.mark.parametrize("B,Ti,To", [(5, 23, 12), (64, 120, 20), (3, 128, 7), (1, 1, 4)])
def test_persistent_decoder_backward_matches_per_step_launches(L, B, Ti, To):
    """decoder_persist_bwd.cu (one resident kernel for the reverse-time loop) against the per-step launch sequence on the same
    saved forward state: d(memory), every decoder weight gradient, the dX sequences and the gate gradients.
    Different split-K order / fp32 dHq instead of a tf32 GEMM: tolerance 2e-3 of each tensor's max (measured ~1e-4)."""
    import os
    from oracle import port
    from t2v import engine
    dev = torch.device("cuda")
    P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
    ops = engine.Ops("tf32")
    g = torch.Generator().manual_seed(B * 1000 + Ti)
    memory = (torch.randn(B, Ti, 512, generator=g) * 0.5).to(dev)
    mel = (torch.randn(B, 80, To, generator=g) * 2 - 5).to(dev)
    in_len = torch.randint(max(1, Ti // 2), Ti + 1, (B,), generator=g).sort(descending=True)[0]
    in_len[0] = Ti
    in_len = in_len.to(dev)
    dO = (torch.randn(To * B, 84, generator=g) * 0.1).to(dev)
    dO[:, 81:] = 0
    outs = {}
    for mode in ("0", "1"):
        os.environ["T2V_PERSIST_BWD"] = mode
        try:
            O, align, ctx = engine.decoder_forward(ops, P, memory, mel, in_len, True, None, None, 77, -float("inf"), dev)
            grads = {}
            dmem, br = engine.decoder_backward(ops, P, dO, ctx, dev, grads)
            br.join()
            torch.cuda.synchronize()
        finally:
            os.environ.pop("T2V_PERSIST_BWD", None)
        t, _ = br.keep
        res = dict(dmem=dmem.clone(), **{k: t[k].clone() for k in ("DXA", "DXD", "DGA", "DGD", "DCTX", "DQ", "dpmem", "dCa", "dCd")})
        res.update({"grad:" + k: v.clone() for k, v in grads.items()})
        outs[mode] = res
    for k in outs["0"]:
        a, b = outs["1"][k], outs["0"][k]
        assert torch.isfinite(a).all(), k
        err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        print("persistent bwd vs per-step %-60s max-rel %.3e" % (k, err))
        assert err <= 2e-3, (k, err)


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_pack_step_tiles_is_a_pure_permutation(L, mode):
    """t2v_pack_step_tiles re-tiles a decoder-step weight matrix into the order the persistent loop kernels stream it:
    bit-exact gather, checked against the index arithmetic documented in decoder_persist.cu (forward: [cluster][rank][chunk]
    [gate*32+unit][32 k], K offsets in the order prenet / h_att / ctx resp. h_att / ctx / h_dec; backward: [cluster][rank]
    [chunk][output column][32 gate rows])."""
    dev = torch.device("cuda")
    K = 1792 if mode in (0, 2) else 2560
    g = torch.Generator().manual_seed(mode)
    if mode < 2:
        W = torch.randn(4096, K, generator=g).to(dev)
        nch = 14 if mode == 0 else 20

        def kofs(j, r):
            if mode == 0:
                return 64 * r + 32 * j if j < 2 else (768 + 256 * r + 32 * (j - 2) if j < 10 else 256 + 128 * r + 32 * (j - 10))
            return 256 * r + 32 * j if j < 8 else (1024 + 128 * r + 32 * (j - 8) if j < 12 else 1536 + 256 * r + 32 * (j - 12))
        ref = torch.empty(32, 4, nch, 128, 32, device=dev)
        rows = torch.arange(128, device=dev)
        for c in (0, 5, 31):
            src_rows = (rows // 32) * 1024 + 32 * c + (rows % 32)
            for r in range(4):
                for j in range(nch):
                    ref[c, r, j] = W[src_rows][:, kofs(j, r):kofs(j, r) + 32]
        check = (0, 5, 31)
    else:
        W = torch.randn(K, 4096, generator=g).to(dev)
        cpc = K // 32
        ref = torch.empty(32, 4, 32, cpc, 32, device=dev)
        for c in (0, 7, 31):
            for r in range(4):
                for j in range(32):
                    ref[c, r, j] = W[cpc * c:cpc * (c + 1), 1024 * r + 32 * j:1024 * r + 32 * j + 32]
        check = (0, 7, 31)
    out = torch.empty(4096 * K, device=dev)
    L("t2v_pack_step_tiles", W, mode, out)
    torch.cuda.synchronize()
    out = out.view(ref.shape)
    for c in check:
        assert torch.equal(out[c], ref[c]), (mode, c)
    assert torch.equal(out.flatten().sort()[0], W.flatten().sort()[0])      # a permutation: nothing lost, nothing duplicated


@pytest.mark.parametrize("B,Ti,n", [(3, 25, 10), (16, 120, 8), (64, 77, 6)])
def test_persistent_free_running_decode_matches_per_step_launches(L, B, Ti, n):
    """Free-running Decoder.inference (prenet -> decode -> mel/gate projection -> feedback) as one persistent kernel
    (dec_persist_fwd_kernel<INFER>) against the per-step launch sequence, same prenet masks, tf32 mode.  The feedback loop amplifies
    rounding differences (the per-step prenet GEMM truncates the unrounded mel to tf32, the persistent kernel keeps fp32), so the
    tolerance is 5e-3 of each tensor's max over a few steps; alignment rows must sum to 1 and the stop bookkeeping must agree."""
    import os
    from oracle import port
    from t2v import engine, infer
    dev = torch.device("cuda")
    P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
    ops = engine.Ops("tf32")
    g = torch.Generator().manual_seed(B * 100 + Ti)
    mem = (torch.randn(B, Ti, 512, generator=g) * 0.5).to(dev)
    pm = (torch.rand(n, 2, B, 256, generator=g) >= 0.5).float().to(dev)
    outs = {}
    for mode in ("0", "1"):
        os.environ["T2V_PERSIST"] = mode
        try:
            sess = infer.DecoderSession(ops, P, mem, None, n, training=False, seed=5)
            nfr = sess.run_free(n, 0.5, prenet_masks=pm.contiguous())
            torch.cuda.synchronize()
        finally:
            os.environ.pop("T2V_PERSIST", None)
        mel, gate, align = sess.outputs(n)
        outs[mode] = dict(mel=mel.clone(), gate=gate.clone(), align=align.clone(), nfr=nfr.clone())
    for k in ("mel", "gate", "align"):
        a, b = outs["1"][k], outs["0"][k]
        assert torch.isfinite(a).all(), k
        err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        print("persistent infer vs per-step %-6s max-rel %.3e" % (k, err))
        assert err <= 5e-3, (k, err)
    assert torch.allclose(outs["1"]["align"].sum(-1), torch.ones(B, n, device=dev), atol=1e-5)
    assert torch.equal(outs["1"]["nfr"], outs["0"]["nfr"])


@pytest.mark.parametrize("rows,n_a,n_b,a0,b0,splits", [(1000, 128, 128, 0, 0, 1), (4099, 512, 512, 2, 3, 7), (51200, 84, 1024, 0, 64, 37),
                                                       (300, 84, 80, 0, 0, 2), (777, 256, 56, 1, 0, 3)])
def test_gemm_tc_rowred_mn_major(L, rows, n_a, n_b, a0, b0, splits):
    """t2v_gemm_tc_rowred: D = A[a0:a0+rows]^T B[b0:b0+rows] straight from the row-major operands (MN-major UMMA descriptors, no
    transposed copies), split over the reduction with atomics.  Operands are pre-rounded to tf32, so only the fp32 summation
    order differs from the fp64 reference: tolerance 1e-4 of the result's max."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(rows + n_a)
    lda, ldb = (n_a + 7) // 4 * 4, n_b + 8
    A = torch.randn(rows + a0 + 5, lda, generator=g).to(dev)
    Bm = torch.randn(rows + b0 + 5, ldb, generator=g).to(dev)
    L("t2v_round_tf32", A, A.numel())
    L("t2v_round_tf32", Bm, Bm.numel())
    D = torch.zeros(n_a, n_b, device=dev)
    L("t2v_gemm_tc_rowred", A, lda, n_a, a0, Bm, ldb, n_b, b0, D, n_b, rows, splits, 0, 1, 1.0, 1)
    torch.cuda.synchronize()
    ref = (A[a0:a0 + rows, :n_a].double().t() @ Bm[b0:b0 + rows, :n_b].double()).float()
    err = float((D - ref).abs().max() / ref.abs().max())
    print("rowred rows=%d %dx%d splits=%d max-rel %.2e" % (rows, n_a, n_b, splits, err))
    assert err < 1e-4, err
    if splits == 1:      # non-atomic accumulate: D += the same product
        L("t2v_gemm_tc_rowred", A, lda, n_a, a0, Bm, ldb, n_b, b0, D, n_b, rows, 1, 0, 2, 1.0, 1)
        torch.cuda.synchronize()
        assert float((D - 2 * ref).abs().max() / ref.abs().max()) < 2e-4


@pytest.mark.parametrize("mode,fmt", [(0, 1), (1, 1), (0, 2), (1, 2)])
def test_pack_step_tiles16(L, mode, fmt):
    """16-bit re-tiling of the decoder-step weights for the fp16 / bf16 persistent loop: [cluster][rank][chunk][gate*32+unit][64 k],
    values = round-to-nearest conversions of the fp32 matrix."""
    dev = torch.device("cuda")
    K = 1792 if mode == 0 else 2560
    g = torch.Generator().manual_seed(mode)
    W = (torch.randn(4096, K, generator=g) * 0.05).to(dev)
    nch = 7 if mode == 0 else 10

    def kofs(j, r):
        if mode == 0:
            return 64 * r if j < 1 else (768 + 256 * r + 64 * (j - 1) if j < 5 else 256 + 128 * r + 64 * (j - 5))
        return 256 * r + 64 * j if j < 4 else (1024 + 128 * r + 64 * (j - 4) if j < 6 else 1536 + 256 * r + 64 * (j - 6))
    out = torch.empty(4096 * K, device=dev, dtype=torch.int16)
    L("t2v_pack_step_tiles16", W, mode, out, fmt)
    torch.cuda.synchronize()
    dt = torch.float16 if fmt == 1 else torch.bfloat16
    got = out.view(dt).view(32, 4, nch, 128, 64)
    rows = torch.arange(128, device=dev)
    for c in (0, 9, 31):
        src_rows = (rows // 32) * 1024 + 32 * c + (rows % 32)
        for r in range(4):
            for j in range(nch):
                ref = W[src_rows][:, kofs(j, r):kofs(j, r) + 64].to(dt)
                assert torch.equal(got[c, r, j], ref), (mode, fmt, c, r, j)


@pytest.mark.parametrize("B,Ti,To,training", [(5, 23, 12, True), (64, 120, 24, True), (16, 128, 9, False)])
def test_persistent_loop_16bit_operands_match_tf32_loop(L, B, Ti, To, training):
    """The fp16-operand instantiation of the persistent forward loop (kind::f16 MMAs over 16-bit copies of the weights and of
    XA / XD) against the tf32 instantiation on the same inputs and the same counter RNG.  fp16 and tf32 share the 11-bit
    significand, so the operand values are identical wherever |x| >= 6.1e-5 and the outputs agree to accumulation-order noise
    (tolerance 5e-4 of each buffer's max; 1.5e-3 = one ulp for XA / XD themselves); bf16 is held to 2e-2."""
    from oracle import port
    from t2v import engine
    dev = torch.device("cuda")
    P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
    g = torch.Generator().manual_seed(B * 1000 + Ti)
    memory = (torch.randn(B, Ti, 512, generator=g) * 0.5).to(dev)
    mel = (torch.randn(B, 80, To, generator=g) * 2 - 5).to(dev)
    in_len = torch.randint(max(1, Ti // 2), Ti + 1, (B,), generator=g).sort(descending=True)[0]
    in_len[0] = Ti
    in_len = in_len.to(dev)
    outs = {}
    for prec in ("tf32", "fp16", "bf16"):
        ops = engine.Ops(prec)
        O, align, ctx = engine.decoder_forward(ops, P, memory, mel, in_len, training, None, None, 77, -float("inf"), dev)
        torch.cuda.synchronize()
        buf = ctx["buf"]
        outs[prec] = dict(O=O.clone(), align=align.clone(), **{k: buf[k].clone() for k in ("XA", "XD", "CA", "CD", "GA", "GD", "HCLO")})
        if prec != "tf32":      # the 16-bit copies are the fp32 rows on the 16-bit grid
            dt = torch.float16 if prec == "fp16" else torch.bfloat16
            for k in ("XA", "XD"):
                assert torch.equal(buf[k + "16"].view(dt)[:To * B].float(), buf[k][:To * B]), k
    for prec, tol in (("fp16", 5e-4), ("bf16", 2e-2)):
        for k in outs["tf32"]:
            a, b = outs[prec][k], outs["tf32"][k]
            assert torch.isfinite(a).all(), k
            if k == "HCLO":
                continue
            err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
            print("%s vs tf32 loop %-6s max-rel %.3e" % (prec, k, err))
            assert err <= (max(tol, 1.5e-3) if k in ("XA", "XD") else tol), (prec, k, err)


@pytest.mark.parametrize("B,Ti,packed", [(64, 120, True), (5, 17, True), (3, 9, False), (1, 1, True)])
def test_persistent_bilstm_matches_per_step_launches(L, B, Ti, packed):
    """rnn_persist.cu (one resident kernel for all Ti steps of both directions, forward and backward) against the per-step launches
    of rnn.cu: same FFMA loops in the same order, so outputs, saved activations and every encoder gradient agree bit for bit
    (gradients that go through atomics: 1e-5 of the max)."""
    from oracle import port
    from t2v import engine
    dev = torch.device("cuda")
    P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
    ops = engine.Ops("fp32")
    g = torch.Generator().manual_seed(B * 100 + Ti)
    text = torch.randint(1, 79, (B, Ti), generator=g).to(dev)
    in_len = torch.randint(max(1, Ti // 2), Ti + 1, (B,), generator=g).sort(descending=True)[0]
    in_len[0] = Ti
    in_len = in_len.to(dev)
    dmem = torch.randn(B, Ti, 512, generator=g).to(dev)
    res = {}
    for mode in (True, False):
        engine._BILSTM_PERSIST = mode
        try:
            n0 = _launches()
            HoutP, ctx = engine.encoder_forward(ops, P, text, in_len if packed else None, True, None, 5, dev, packed=packed)
            n_fwd = _launches() - n0
            grads = {}
            engine.encoder_backward(ops, P, dmem.clone(), ctx, True, 5, dev, grads)
            torch.cuda.synchronize()
        finally:
            engine._BILSTM_PERSIST = True
        res[mode] = dict(HoutP=HoutP.clone(), GS=ctx["GS"].clone(), CS=ctx["CS"].clone(), n_fwd=n_fwd,
                         **{"g:" + k: v.clone() for k, v in grads.items() if k.startswith("encoder.lstm")})
    # Ti step launches became one (+ the zero-fill of its counters and exchange buffer, which are library kernels too)
    assert res[True]["n_fwd"] <= res[False]["n_fwd"] - (Ti - 1) + 4
    for k in res[True]:
        if k == "n_fwd":
            continue
        a, b = res[True][k], res[False][k]
        if k.startswith("g:"):
            assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-12, k
        else:
            assert torch.equal(a, b), k


@pytest.mark.parametrize("N,To", [(64, 800), (5, 130), (1, 64)])
def test_persistent_gru_matches_per_step_launches(L, N, To):
    """Reference-encoder GRU (modules.py:60-80) as one resident kernel per pass (rnn_persist.cu) against the per-step GEMM + pointwise
    launches: the forward state is bit-identical (same FFMA order); gradients differ only by summation order of the batched dW_hh."""
    from oracle import port
    from t2v import engine
    dev = torch.device("cuda")
    P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
    ops = engine.Ops("fp32")
    g = torch.Generator().manual_seed(N * 7 + To)
    mel = (torch.randn(N, 80, To, generator=g) * 2 - 5).to(dev)
    dh = torch.randn(N, 256, generator=g).to(dev)
    res = {}
    for mode in (True, False):
        engine._BILSTM_PERSIST = mode
        try:
            n0 = _launches()
            h, ctx = engine.refenc_forward(ops, P, mel, True, dev)
            n_fwd = _launches() - n0
            grads = {}
            engine.refenc_backward(ops, P, dh.clone(), ctx, True, dev, grads)
            torch.cuda.synchronize()
        finally:
            engine._BILSTM_PERSIST = True
        res[mode] = dict(h=h.clone(), HS=ctx["HS"].clone(), SV=ctx["SV"].clone(), n_fwd=n_fwd,
                         **{"g:" + k: v.clone() for k, v in grads.items()})
    Tq = res[True]["HS"].shape[0] - 1
    assert res[True]["n_fwd"] <= res[False]["n_fwd"] - (2 * Tq - 1) + 4      # + the zero-fills of its counters (library kernels too)
    for k in res[True]:
        if k == "n_fwd":
            continue
        a, b = res[True][k], res[False][k]
        if k.startswith("g:"):
            assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-12, k
        else:
            assert torch.equal(a, b), k


def _launches():
    from t2v import _lib
    return _lib.launch_count()


@pytest.mark.parametrize("B,Ti,To,gscale", [(5, 23, 12, 0.1), (64, 120, 20, 1e-7), (3, 128, 7, 3e-4)])
def test_persistent_backward_fp16_operands_match_tf32_loop(L, B, Ti, To, gscale):
    """The fp16-operand instantiation of the persistent backward loop (W^T tiles and gate gradients as fp16 copies, gradients scaled by a
    power of two derived from max |dO|) against the tf32 instantiation on the same saved forward state, for upstream gradients of very
    different magnitudes (1e-7 is what a C3-sized loss produces: far below the fp16 range without the scale).  fp16 and tf32 share
    the 11-bit significand: tolerance 2e-3 of each tensor's max."""
    from oracle import port
    from t2v import engine
    dev = torch.device("cuda")
    P = {k: v.to(dev) for k, v in port.init_params(1234).items()}
    g = torch.Generator().manual_seed(B * 1000 + Ti)
    memory = (torch.randn(B, Ti, 512, generator=g) * 0.5).to(dev)
    mel = (torch.randn(B, 80, To, generator=g) * 2 - 5).to(dev)
    in_len = torch.randint(max(1, Ti // 2), Ti + 1, (B,), generator=g).sort(descending=True)[0]
    in_len[0] = Ti
    in_len = in_len.to(dev)
    dO = (torch.randn(To * B, 84, generator=g) * gscale).to(dev)
    dO[:, 81:] = 0
    ops = engine.Ops("fp16")
    O, align, ctx = engine.decoder_forward(ops, P, memory, mel, in_len, True, None, None, 77, -float("inf"), dev)
    outs = {}
    for mode in (True, False):
        engine._BWD16 = mode
        try:
            grads = {}
            dmem, br = engine.decoder_backward(ops, P, dO, ctx, dev, grads)
            br.join()
            torch.cuda.synchronize()
        finally:
            engine._BWD16 = True
        t, D = br.keep
        assert int(D.op16) == (1 if mode else 0)
        res = dict(dmem=dmem.clone(), **{k: t[k].clone() for k in ("DXA", "DXD", "DGA", "DGD", "DCTX", "DQ", "dpmem")})
        res.update({"grad:" + k: v.clone() for k, v in grads.items()})
        if mode:
            sc = t["scale"].cpu()
            assert float(sc[0]) * float(sc[1]) == 1.0 and 2048.0 <= float(sc[0]) * float(dO.abs().max()) < 4096.0 + 1e-3
            dg16 = t["DGA16"].view(torch.float16).float() * float(sc[1])
            assert float((dg16 - t["DGA"]).abs().max()) <= 2e-3 * float(t["DGA"].abs().max())       # the copies are the same numbers
        outs[mode] = res
    for k in outs[False]:
        a, b = outs[True][k], outs[False][k]
        assert torch.isfinite(a).all(), k
        err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        print("bwd fp16 vs tf32 %-60s max-rel %.3e" % (k, err))
        assert err <= 2e-3, (k, err)


@pytest.mark.parametrize("rows,n_a,n_b,splits", [(4096, 4096, 1792, 2), (1000, 512, 256, 1), (51200, 4096, 2560, 5)])
def test_gemm_tc_rowred16_mn_major_fp16(L, rows, n_a, n_b, splits):
    """t2v_gemm_tc_rowred16: D = alpha_dev * A16^T B16 from row-major fp16 operands (MN-major kind::f16 UMMA, 256 x 256 tiles, 64
    reduction rows per stage) against an fp64 reference of the same fp16 values: only the fp32 accumulation order differs."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(rows + n_a)
    A = torch.randn(rows + 3, n_a, generator=g).to(dev).half()
    Bm = torch.randn(rows + 3, n_b, generator=g).to(dev).half()
    D = torch.zeros(n_a, n_b, device=dev)
    alpha = torch.tensor([0.25], device=dev)
    L("t2v_gemm_tc_rowred16", A.view(torch.int16), n_a, n_a, 0, Bm.view(torch.int16), n_b, n_b, 0, D, n_b, rows, splits, 1, 1.0, alpha, 1, 1, None, 0, 0)
    torch.cuda.synchronize()
    ref = 0.25 * (A[:rows].double().t() @ Bm[:rows].double()).float()
    err = float((D - ref).abs().max() / ref.abs().max())
    print("rowred16 rows=%d %dx%d splits=%d max-rel %.2e" % (rows, n_a, n_b, splits, err))
    assert err < 1e-4, err


@pytest.mark.parametrize("bits,rows,n_a,n_b,splits", [(32, 2000, 512, 512, 3), (32, 1500, 512, 128, 2), (32, 900, 80, 512, 4),
                                                      (16, 6000, 512, 512, 5), (16, 700, 256, 256, 1)])
def test_gemm_tc_rowred_taps(L, bits, rows, n_a, n_b, splits):
    """Conv1d weight gradient in one launch (taps = 5): D[:, t*n_b:(t+1)*n_b] = A[2:2+rows]^T B[t:t+rows] for every tap t, through
    t2v_gemm_tc_rowred (tf32) and t2v_gemm_tc_rowred16 (fp16), against fp64 of the same rounded operands."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(rows + n_a + bits)
    A = torch.randn(rows + 4, n_a, generator=g).to(dev)
    Bm = torch.randn(rows + 4, n_b, generator=g).to(dev)
    D = torch.zeros(n_a, 5 * n_b, device=dev)
    if bits == 32:
        L("t2v_round_tf32", A, A.numel())
        L("t2v_round_tf32", Bm, Bm.numel())
        L("t2v_gemm_tc_rowred", A, n_a, n_a, 2, Bm, n_b, n_b, 0, D, 5 * n_b, rows, splits, 0, 1, 1.0, 5)
    else:
        A, Bm = A.half(), Bm.half()
        L("t2v_gemm_tc_rowred16", A.view(torch.int16), n_a, n_a, 2, Bm.view(torch.int16), n_b, n_b, 0, D, 5 * n_b, rows, splits, 1, 1.0,
          None, 1, 5, None, 0, 0)
    torch.cuda.synchronize()
    ref = torch.cat([(A[2:2 + rows].double().t() @ Bm[t:t + rows].double()).float() for t in range(5)], dim=1)
    err = float((D - ref).abs().max() / ref.abs().max())
    print("rowred taps bits=%d rows=%d %dx%d splits=%d max-rel %.2e" % (bits, rows, n_a, n_b, splits, err))
    assert err < 1e-4, err


@pytest.mark.parametrize("batch,rows,n_a,n_b", [(3, 100, 120, 512), (64, 800, 120, 512), (2, 37, 40, 256)])
def test_gemm_tc_rowred_batched(L, batch, rows, n_a, n_b):
    """t2v_gemm_tc_rowred_batched: D[z] = A[z*rows:(z+1)*rows]^T B[:, z*n_b:(z+1)*n_b] (the d(memory) GEMM of the attention backward:
    alignments [B,To,Ti] x dctx [To,B,512]) against fp64 of the same tf32-rounded operands; rows % 32 != 0 exercises the zero-filled
    reduction tail (A's tail rows belong to the next batch)."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(batch * 1000 + rows)
    A = torch.rand(batch, rows, n_a, generator=g).to(dev)
    Bm = torch.randn(rows, batch, n_b, generator=g).to(dev)
    L("t2v_round_tf32", A, A.numel())
    L("t2v_round_tf32", Bm, Bm.numel())
    D = torch.full((batch, n_a, n_b), float("nan"), device=dev)
    L("t2v_gemm_tc_rowred_batched", A, n_a, n_a, rows, Bm, batch * n_b, n_b, n_b, D, n_b, n_a * n_b, rows, batch, 1.0)
    torch.cuda.synchronize()
    ref = torch.einsum("zri,rzj->zij", A.double(), Bm.double()).float()
    err = float((D - ref).abs().max() / ref.abs().max())
    print("rowred batched %d x [%d -> %d x %d] max-rel %.2e" % (batch, rows, n_a, n_b, err))
    assert err < 1e-4, err


def test_gemm_tc_rowred16_column_split(L):
    """n_split: columns [0, 768) of dW = A^T B land in D, columns [768, 1792) in D2 (the [weight_ih | weight_hh] gradients of an LSTMCell)"""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    rows, n_a, n_b = 3000, 512, 1792
    A = torch.randn(rows, n_a, generator=g).to(dev).half()
    Bm = torch.randn(rows, n_b, generator=g).to(dev).half()
    D1, D2 = torch.zeros(n_a, 768, device=dev), torch.zeros(n_a, 1024, device=dev)
    L("t2v_gemm_tc_rowred16", A.view(torch.int16), n_a, n_a, 0, Bm.view(torch.int16), n_b, n_b, 0, D1, 768, rows, 3, 1, 1.0, None, 1, 1,
      D2, 1024, 768)
    torch.cuda.synchronize()
    ref = (A.double().t() @ Bm.double()).float()
    err = float((torch.cat([D1, D2], dim=1) - ref).abs().max() / ref.abs().max())
    assert err < 1e-4, err


@pytest.mark.parametrize("M,N,Ci,taps", [(3000, 512, 512, 5), (1000, 512, 80, 5), (2000, 80, 512, 5), (700, 256, 192, 1)])
def test_gemm_tc_split3_16_matches_fp64(L, M, N, Ci, taps):
    """t2v_split16 + t2v_gemm_tc_split3_16: x = hi + lo fp16 pairs (weights scaled by 16), three kind::f16 products through one
    accumulator -> fp32-level accuracy (the Postnet forward of the fp16 mode, model.py:105-148, as a 5-tap row-shifted GEMM)."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(M + N)
    R = M + taps - 1
    X = (torch.rand(R, Ci, generator=g) * 2 - 1).to(dev)
    W = (torch.randn(N, taps * Ci, generator=g) * 0.05).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    Xh, Xl = (torch.empty(R, Ci, device=dev, dtype=torch.int16) for _ in range(2))
    Wh, Wl = (torch.empty(N, taps * Ci, device=dev, dtype=torch.int16) for _ in range(2))
    L("t2v_split16", X, Xh, Xl, X.numel(), 1.0)
    L("t2v_split16", W, Wh, Wl, W.numel(), 16.0)
    # the pair reproduces x to ~2^-22
    xr = Xh.view(torch.float16).float() + Xl.view(torch.float16).float()
    assert float((xr - X).abs().max()) < 2e-7
    D = torch.zeros(M, N, device=dev)
    L("t2v_gemm_tc_split3_16", Xh, Xl, Ci, R, Ci, Wh, Wl, taps * Ci, N, taps * Ci, D, N, bias, M, N, Ci, taps, 1, Ci, 0, 0, 1.0 / 16.0,
      256 if N % 256 == 0 else 128)
    torch.cuda.synchronize()
    ref = bias.double()[None, :].repeat(M, 1)
    for t in range(taps):
        ref += X[t:t + M].double() @ W[:, t * Ci:(t + 1) * Ci].double().t()
    err = float((D - ref.float()).abs().max() / ref.abs().max())
    print("split3_16 M=%d N=%d Ci=%d taps=%d max-rel %.2e" % (M, N, Ci, taps, err))
    # what is left is the tensor core's own accumulation (one truncating fp32 add per K = 16 instruction, 3 x K / 16 of them): ~1e-5
    # at K = 2560, against 5e-4 for a single fp16 / tf32 product
    assert err < 5e-5, err
