"""Forward / backward orchestration of the sm_100a kernels behind model.Tacotron2.

Every arithmetic step of the hot path is a call into libt2v_b200.so (t2v._lib); torch is used only to allocate device
buffers and to carry pointers.  Two precision modes:
   "fp32"  every GEMM on the exact FFMA kernel (t2v_gemm_f32)                 -- tight parity mode
   "tf32"  the large GEMMs on the tcgen05/TMA kernel (t2v_gemm_tc, tf32 math) -- the fast mode
Layouts: Conv1d stacks use padded channels-last rows [B*(T+4), C]; the decoder uses time-major rows (t*B+b).

Reference map (file:line in the reference tree):
   encoder_forward   model.py:151-203      refenc_forward   modules.py:34-85 + CoordConv.py:37-74,142-161
   vae_head          modules.py:8-31       prenet           model.py:91-102
   decoder           model.py:206-464      postnet          model.py:105-148
   outputs/masking   model.py:509-547      (loss: loss_function.py -> t2v/functions.py)
"""
import math

import torch

from . import _lib
from ._lib import call as L

F32 = torch.float32
_BN256 = __import__("os").environ.get("T2V_BN256", "1") != "0"
_TRACE = bool(int(__import__("os").environ.get("T2V_TRACE", "0")))
_t_last = [None]


_NVTX = bool(int(__import__("os").environ.get("T2V_NVTX", "0")))


def _trace(label):
    """phase boundary of the forward / backward orchestration.  T2V_NVTX=1: an NVTX mark per boundary (eager launches: the marks
    line up with the kernels in an Nsight timeline; a captured graph replays without them).  T2V_TRACE=1: synchronise and print
    the wall time since the previous trace point (debug / coarse profiling)"""
    if _NVTX and label:
        torch.cuda.nvtx.mark("t2v: end of " + label.strip())
    if not _TRACE:
        return
    import time
    host = time.perf_counter()
    torch.cuda.synchronize()
    now = time.perf_counter()
    if _t_last[0] is not None and label:
        print("[t2v] %-28s %9.3f ms  (host issue %8.3f ms)" % (label, (now - _t_last[0]) * 1e3, (host - _t_last[0]) * 1e3), flush=True)
    _t_last[0] = now
SITE_ENC, SITE_PRENET, SITE_POST, SITE_EPS = 0, 3, 20, 30          # dropout RNG site ids (decoder uses 10..13 in decoder.cu)
_ctypes = _lib.ctypes


_SIDE = {}
_OVERLAP = __import__("os").environ.get("T2V_OVERLAP", "1") != "0"
# Postnet weight gradients on a low-priority side branch beside the dX chain (and, if they last that long, beside the persistent
# decoder-backward kernel on the SMs it leaves free).  Round 1 kept this off: launch failures next to the two-MMA-issuer backward
# kernel.  Since then the persistent kernels are launched cooperatively (placed as a whole or not yet at all), the single-issuer
# kernel soaked clean for 200 steps with the branch (profiles/r02_soak_200_steps_post_dw_branch.json), and with all five taps in one
# launch the branch finishes long before the backward loop starts.  T2V_POST_DW_BRANCH=0 puts the GEMMs back on the main chain.
_POST_DW_BRANCH = __import__("os").environ.get("T2V_POST_DW_BRANCH", "1") == "1"
_BILSTM_PERSIST = __import__("os").environ.get("T2V_BILSTM_PERSIST", "1") != "0"
_DW16 = __import__("os").environ.get("T2V_DW16", "1") != "0"          # decoder weight gradients from the fp16 copies (fp16 mode)
_PRIO = __import__("os").environ.get("T2V_PRIO", "1") != "0"          # high-priority side streams for critical-path branches
_BIAS_IN_LOOP = __import__("os").environ.get("T2V_BIAS_IN_LOOP", "1") != "0"   # LSTM bias gradients from the backward loop kernel
_DMEM_TC = __import__("os").environ.get("T2V_DMEM_TC", "1") != "0"    # d(memory) = alignments^T dctx on the tensor core (tf32)
_SPLIT16 = __import__("os").environ.get("T2V_SPLIT16", "1") != "0"    # fp16 mode: Postnet forward on fp16 hi / lo pairs
_GRU_BWD_TC = __import__("os").environ.get("T2V_GRU_BWD_TC", "1") != "0"   # reference-encoder GRU: backward GEMMs on the tensor core
_REFENC_LATE = __import__("os").environ.get("T2V_REFENC_LATE", "0") == "1"   # fork the reference-encoder branch after the encoder convs (measured: +0.26 ms)
_TAPS1 = __import__("os").environ.get("T2V_TAPS1", "1") != "0"        # Conv1d weight gradients: all taps in one row-reduction launch
_BWD16 = __import__("os").environ.get("T2V_BWD16", "1") != "0"        # fp16 operand copies in the persistent backward loop (op16 modes)


class _Branch(object):
    """`with _Branch(i):` runs the block on side stream i, ordered after everything already enqueued on the current stream
    (a parallel branch of the train-step CUDA graph when captured); `.join()` makes the current stream wait for it.
    Tensors created inside belong to the side stream's allocator pool; every later use of a side stream starts with a
    new fork, so block reuse stays stream-ordered.  T2V_OVERLAP=0 runs everything in order on one stream."""

    def __init__(self, idx, urgent=False):
        """urgent: the branch is part of the step's critical path -> same (high) priority as the graph's capture stream; other
        branches (weight gradients nobody waits for until the optimizer step) keep the default priority, so their CTAs only take
        the SMs the critical chains leave free (kernel-node priorities survive graph capture)"""
        self.keep = None
        self.main = torch.cuda.current_stream()
        if _OVERLAP:
            prio = -1 if (urgent and _PRIO) else 0
            key = (self.main.device_index, idx, prio)
            if key not in _SIDE:
                _SIDE[key] = torch.cuda.Stream(device=self.main.device, priority=prio)
            self.side = _SIDE[key]
        else:
            self.side = None
        self._ctx = None

    def __enter__(self):
        if self.side is not None:
            self.side.wait_stream(self.main)
            self._ctx = torch.cuda.stream(self.side)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
            self._ctx = None
        return False

    def join(self):
        if self.side is not None:
            self.main.wait_stream(self.side)


def _p(t, off_elems=0):
    """raw device pointer (int) of tensor `t` advanced by off_elems fp32 elements"""
    return t.data_ptr() + 4 * off_elems


def _zeros(*shape, device, dtype=F32):
    """zero-initialised buffer, filled by a KERNEL of the library: memset / memcpy graph nodes (torch.zeros of some dtypes, tensor
    clones) do not carry the priority of the stream they were captured on -- inside the train-step graph they queue behind every
    pending CTA of the low-priority weight-gradient GEMMs (observed: a 1 us copy at the head of the reference-encoder backward
    waited 0.8 ms for a 960-CTA GEMM to finish dispatching), kernel nodes do not."""
    t = torch.empty(*shape, device=device, dtype=dtype)
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return t
    if nbytes % 4 or not t.is_cuda:
        return t.zero_()
    L("t2v_fill", t, nbytes // 4, 0.0)
    return t


def _clone(x):
    """contiguous fp32 copy by a kernel (see _zeros: no memcpy nodes on the step's chains)"""
    x = x.detach()
    if x.dtype != F32 or not x.is_contiguous() or x.numel() == 0:
        return torch.clone(x)
    y = torch.empty_like(x)
    L("t2v_axpby", x, 1.0, y, 0.0, x.numel())
    return y


def _empty(*shape, device, dtype=F32):
    return torch.empty(*shape, device=device, dtype=dtype)


def _ceil4(n):
    return (n + 3) // 4 * 4


class Ops(object):
    """GEMM front-end that picks the kernel for the precision mode."""

    def __init__(self, precision="fp32"):
        assert precision in ("fp32", "tf32", "fp16", "bf16")
        self.precision = precision
        self.tc = precision != "fp32"
        self.R = 1 if self.tc else 0       # producers round tensor-core operands to the tf32 grid on store
        # the persistent decoder loops stream 16-bit operand copies in the fp16 / bf16 modes (kind::f16 MMAs): fp16 keeps the
        # 11-bit significand of tf32 at half the bytes; everything outside the two loops runs as in the tf32 mode
        self.op16 = {"fp16": 1, "bf16": 2}.get(precision, 0)
        self.RX = (self.op16 + 1) if self.op16 else self.R     # rounding grid of the decoder-loop activations (ABI `rnd` code)
        # split (error-compensated) tensor-core GEMMs x = x_hi + x_lo for the two places where tf32 rounding dominates the
        # error of the outputs: the deferred mel / gate projection and the Postnet forward (DESIGN.md "Precision modes")
        self.split = self.tc and __import__("os").environ.get("T2V_SPLIT", "1") != "0"
        self.p_att = self.p_dec = 0.1      # hparams.p_attention_dropout / p_decoder_dropout (model.py sets them)

    @staticmethod
    def bn(M, N):
        """N-tile of the tcgen05 GEMM: 256 for the large GEMMs (halves the activation-tile re-reads per FLOP)"""
        return 256 if (N >= 512 and N % 256 == 0 and M >= 4096 and _BN256) else 128

    def lo(self, W, Whi):
        """low part of a split weight: tf32(W - W_hi)"""
        Wl = _clone(W)
        L("t2v_axpby", Whi, -1.0, Wl, 1.0, Wl.numel())
        L("t2v_round_tf32", Wl, Wl.numel())
        return Wl

    def wr(self, W):
        """weights consumed directly by a tensor-core GEMM: tf32-rounded copy (tcgen05 truncates otherwise)"""
        if not self.tc:
            return W
        Wc = _clone(W)
        L("t2v_round_tf32", Wc, Wc.numel())
        return Wc

    # ---- generic strided fp32 GEMM (always exact) ----
    @staticmethod
    def gemm(A, a_rs, a_cs, Bm, b_rs, b_cs, C, c_rs, M, N, K, alpha=1.0, beta=0.0, bias=None):
        L("t2v_gemm_f32", A, a_rs, a_cs, Bm, b_rs, b_cs, C, c_rs, M, N, K, alpha, beta, bias, 1, 0, 0, 0)

    @staticmethod
    def _tc_ok(*ptr_ld):
        for ptr, ld in ptr_ld:
            p = ptr if isinstance(ptr, int) else ptr.data_ptr()
            if p % 16 or (ld * 4) % 16:
                return False
        return True

    # out[M,N] = x[M,K] @ W[N,K]^T (+bias) (+= if accumulate)
    def linear(self, x, lda, W, ldw, out, ldd, M, N, K, bias=None, accumulate=False, a_rows=None, force_exact=False):
        if self.tc and not force_exact and self._tc_ok((x, lda), (W, ldw)) and M >= 1:
            # accumulate: the atomic epilogue with one split = exactly one add per element (deterministic); the read-modify-write
            # epilogue (epi 2) walks D row by row per thread and measured 10x slower
            L("t2v_gemm_tc", x, lda, a_rows or M, K, W, ldw, N, K, out, ldd, bias, M, N, K, 1, 0, 0, 0, 0, 4, 1, 0,
              1 if accumulate else 0, 1.0, self.bn(M, N))
        else:
            self.gemm(x, lda, 1, W, ldw, 1, out, ldd, M, N, K, 1.0, 1.0 if accumulate else 0.0, bias)

    # dx[M,K] = dy[M,N] @ W[N,K]
    def linear_dx(self, dy, ldy, W, ldw, dx, lddx, M, N, K, accumulate=False, force_exact=False):
        dev_ok = self.tc and not force_exact and self._tc_ok((dy, ldy)) and isinstance(W, torch.Tensor)
        if dev_ok:
            Np = _ceil4(N)
            WT = _zeros(K, Np, device=W.device)
            L("t2v_transpose", W, ldw, WT, Np, N, K, 1)
            L("t2v_gemm_tc", dy, ldy, M, N, WT, Np, K, N, dx, lddx, None, M, K, N, 1, 0, 0, 0, 0, 4, 1, 0,
              1 if accumulate else 0, 1.0, 128)
        else:
            self.gemm(dy, ldy, 1, W, 1, ldw, dx, lddx, M, K, N, 1.0, 1.0 if accumulate else 0.0, None)

    # dW[N,K] = dy[M,N]^T @ x[M,K]   (reduction over the M rows); dW must be zero-initialised unless accumulate
    def linear_dw(self, dy, ldy, x, ldx, dW, lddw, M, N, K, accumulate=False, device=None, force_exact=False):
        if self.tc and not force_exact and M >= 256 and self._tc_ok((dy, ldy), (x, ldx)):
            self.rowred(dy, ldy, N, 0, x, ldx, K, 0, dW, lddw, M)
        else:
            self.gemm(dy, 1, ldy, x, 1, ldx, dW, lddw, N, K, M, 1.0, 1.0 if accumulate else 0.0, None)

    @staticmethod
    def pick_splits(tiles, iters):
        """split count of a row-reduction GEMM: fill whole waves of the 148 SMs (each extra split costs one more pass of atomics over D)"""
        if tiles * 8 < 148:            # few output tiles (conv / small-N weight gradients): as many splits as it takes to fill two waves
            return max(1, min(iters // 8 if iters >= 8 else 1, (296 + tiles - 1) // tiles))
        best, splits = 0.0, 1
        for sp in range(1, 9):
            if sp > iters or iters // sp < 16:
                break
            n = tiles * sp
            eff = n / (((n + 147) // 148) * 148.0) - 0.02 * (sp - 1)
            if eff > best + 1e-9:
                best, splits = eff, sp
        return splits

    @staticmethod
    def rowred16(A16, lda, n_a, B16, ldb, n_b, D, ldd, rows, alpha_dev, fmt=1, a_row0=0, b_row0=0, taps=1, D2=None, ldd2=0, n_split=0):
        """D[n_a, taps*n_b] += (*alpha_dev) * A16[a_row0 + r]^T B16[b_row0 + tap + r] over fp16 copies (kind::f16, 256 x 256 tiles);
        D zero-initialised.  taps > 1: all taps of a Conv1d weight gradient in one launch"""
        iters = (rows + 63) // 64
        tiles = ((n_a + 255) // 256) * (n_b * taps // 256)
        L("t2v_gemm_tc_rowred16", A16, lda, n_a, a_row0, B16, ldb, n_b, b_row0, D, ldd, rows, Ops.pick_splits(tiles, iters), 1, 1.0,
          alpha_dev, fmt, taps, D2, ldd2, n_split)

    @staticmethod
    def rowred(A, lda, n_a, a_row0, Bm, ldb, n_b, b_row0, D, ldd, rows, taps=1):
        """D[n_a, taps*n_b] += sum_r A[a_row0+r, :]^T B[b_row0+tap+r, :] on the MN-major tcgen05 kernel (no transposed copies);
        split over the reduction so that ~2 CTAs per SM exist.  D must be zero-initialised (or hold the running sum)."""
        iters = (rows + 31) // 32
        wide = n_b > 128 and n_b % 256 == 0
        bm = 256 if (wide and n_a >= 512) else 128            # the tile shapes t2v_gemm_tc_rowred picks
        tiles = ((n_a + bm - 1) // bm) * ((n_b + 255) // 256 if wide else (n_b + 127) // 128) * taps
        # split the reduction so that the CTAs fill whole waves of the 148 SMs (each extra split costs one more pass of atomics over D)
        splits = Ops.pick_splits(tiles, iters)
        L("t2v_gemm_tc_rowred", A, lda, n_a, a_row0, Bm, ldb, n_b, b_row0, D, ldd, rows, splits, 0, 1, 1.0, taps)

    @staticmethod
    def _tc_reduce_rows(AT, lda, n_a, a_k0, BT, ldb, n_b, b_k0, D, ldd, Mred, accumulate, a_inner=None, b_inner=None):
        """D[n_a, n_b] (+)= sum_r AT[:, a_k0+r] * BT[:, b_k0+r], r < Mred, with split-K sized to fill the SMs."""
        iters = (Mred + 31) // 32
        tiles = ((n_a + 127) // 128) * ((n_b + 127) // 128)
        want = max(1, min(iters, int(round(296.0 / tiles))))
        splits = 1
        for s in range(want, 0, -1):
            if iters % s == 0:
                splits = s
                break
        atomic = 1 if (splits > 1 or accumulate) else 0     # destination must be zero-initialised by the caller
        L("t2v_gemm_tc", AT, lda, n_a, a_inner or (a_k0 + Mred), BT, ldb, n_b, b_inner or (b_k0 + Mred), D, ldd, None, n_a, n_b,
          Mred, 1, 0, 0, a_k0, b_k0, 4, splits, 0, atomic, 1.0, 128)
        return splits


# ======================================================================================================= helpers
def _bn_forward(ops, Y, Xout, rows, C, period, lo, hi, n_valid, P, pre, training, act, mask, seed, site, p, T, dev,
                update_running=True, rnd=0, out_lo=None, out16=None):
    """BatchNorm (batch stats in training, running stats in eval) + activation + dropout.  Returns (mean, invstd)."""
    mean = _empty(C, device=dev)
    invstd = _empty(C, device=dev)
    if training:
        sums = _zeros(2, C, device=dev, dtype=torch.float64)
        L("t2v_col_stats", Y, rows, C, period, lo, hi, 0, sums[0], sums[1])
        L("t2v_bn_finalize", sums[0], sums[1], float(n_valid), C, 1e-5, 0.1, mean, invstd,
          P[pre + ".running_mean"] if update_running else None, P[pre + ".running_var"] if update_running else None,
          P[pre + ".num_batches_tracked"] if update_running else None)
    else:
        L("t2v_bn_eval_prepare", P[pre + ".running_mean"], P[pre + ".running_var"], C, 1e-5, mean, invstd)
    L("t2v_bn_act_fwd", Y, Xout, out_lo, rows, C, period, lo, hi, mean, invstd, P[pre + ".weight"], P[pre + ".bias"], act,
      mask, seed, site, p, T, rnd, out16[0] if out16 else None, out16[1] if out16 else None)
    return mean, invstd


def _bn_backward(dOut, Y, dY, rows, C, period, lo, hi, n_valid, P, pre, training, act, mask, seed, site, p, T, dev, grads,
                 rnd=0):
    sums = _zeros(2, C, device=dev, dtype=torch.float64)
    mean, invstd = Y.mean_invstd
    L("t2v_bn_act_bwd_reduce", dOut, Y.t, rows, C, period, lo, hi, mean, invstd, P[pre + ".weight"], P[pre + ".bias"], act,
      mask, seed, site, p, T, sums[0], sums[1])
    gw = _empty(C, device=dev)
    gb = _empty(C, device=dev)
    L("t2v_double_to_float", sums[1], gw, C, 0.0)
    L("t2v_double_to_float", sums[0], gb, C, 0.0)
    grads[pre + ".weight"] = gw
    grads[pre + ".bias"] = gb
    L("t2v_bn_act_bwd_apply", dOut, Y.t, dY, rows, C, period, lo, hi, mean, invstd, P[pre + ".weight"], P[pre + ".bias"],
      act, mask, seed, site, p, T, sums[0], sums[1], float(n_valid), 1 if training else 0, rnd)


def _colsum(x, rows, C, period, lo, hi, dev, ld=None):
    acc = _zeros(C, device=dev, dtype=torch.float64)
    L("t2v_col_stats", x, rows, C, period, lo, hi, 2, acc, None)
    out = _empty(C, device=dev)
    L("t2v_double_to_float", acc, out, C, 0.0)
    return out


class Grads(dict):
    """name -> gradient tensor.  `dst` (optional) maps parameter names to pre-allocated destinations (the slices of the model's flat
    gradient buffer, t2v.optim.FlatGrads): producers of the large gradients write there directly, so the step needs no
    concatenation pass over them afterwards."""

    def __init__(self, dst=None):
        dict.__init__(self)
        self.dst = dst or {}

    def out(self, name, *shape, dev, zero=False):
        v = self.dst.get(name)
        if v is not None and tuple(v.shape) == tuple(shape) and v.is_contiguous():
            if zero:
                v.zero_()
            return v
        return _zeros(*shape, device=dev) if zero else _empty(*shape, device=dev)


def _gout(grads, name, *shape, dev, zero=False):
    """destination of a large gradient: the flat-buffer slice when `grads` carries one (Grads.dst), else a fresh tensor"""
    if isinstance(grads, Grads):
        return grads.out(name, *shape, dev=dev, zero=zero)
    return _zeros(*shape, device=dev) if zero else _empty(*shape, device=dev)


class _Saved(object):
    """pre-BN tensor + its statistics"""

    def __init__(self, t, mean_invstd):
        self.t = t
        self.mean_invstd = mean_invstd


# ======================================================================================================= conv1d stacks
def conv_stack_forward(ops, P, prefix, X, B, T, chans, acts, training, masks, seed, site0, dev, round_last=True, Xlo=None, X16=None):
    """k=5/p=2 Conv1d + BatchNorm1d + act + dropout(.5) layers over padded channels-last rows (Encoder
    model.py:159-177, Postnet model.py:105-148).  X: [B*(T+4), chans[0]].  Returns (out, saved).
    Xlo: the low part of the split input (x = X + Xlo, both on the tf32 grid) -> every layer runs as the three-term split
    x_hi W_hi + x_lo W_hi + x_hi W_lo on the tensor cores (fp32-level accuracy at 3x the tf32 GEMM time)."""
    Tp = T + 4
    R = B * Tp
    M = R - 4
    saved = []
    split = Xlo is not None and ops.tc
    # fp16 mode: the split operands are fp16 hi / lo pairs (X16 = the pair of the stack's input) -- the same 22 significant bits as
    # the tf32 pairs at half the bytes; the split GEMMs are bound by the L2 -> shared-memory operand stream
    split16 = X16 is not None and ops.tc
    WS = 16.0            # power-of-two weight scale of the fp16 pairs: lifts W_lo (~2^-11 |W|) out of the fp16 subnormals
    for i in range(len(chans) - 1):
        Ci, Co = chans[i], chans[i + 1]
        pre = "%s.%d" % (prefix, i)
        W = P[pre + ".0.conv.weight"]
        Wk = _empty(Co, 5 * Ci, device=dev)
        L("t2v_conv1d_pack", W, Wk, Co, Ci, 5, 0, 0 if split16 else ops.R)
        Y = _zeros(R, Co, device=dev)
        if ops.tc:
            bn = ops.bn(M, Co)
            if split16:
                Whi = torch.empty(Co, 5 * Ci, device=dev, dtype=torch.int16)
                Wlo = torch.empty(Co, 5 * Ci, device=dev, dtype=torch.int16)
                L("t2v_split16", Wk, Whi, Wlo, Co * 5 * Ci, WS)
                L("t2v_gemm_tc_split3_16", X16[0], X16[1], Ci, R, Ci, Whi, Wlo, 5 * Ci, Co, 5 * Ci, _p(Y, 2 * Co), Co,
                  P[pre + ".0.conv.bias"], M, Co, Ci, 5, 1, Ci, 0, 0, 1.0 / WS, bn)
            elif split:
                Wf = _empty(Co, 5 * Ci, device=dev)
                L("t2v_conv1d_pack", W, Wf, Co, Ci, 5, 0, 0)
                Wl = ops.lo(Wf, Wk)
                L("t2v_gemm_tc_split3", X, Xlo, Ci, R, Ci, Wk, Wl, 5 * Ci, Co, 5 * Ci, _p(Y, 2 * Co), Co, P[pre + ".0.conv.bias"],
                  M, Co, Ci, 5, 1, Ci, 0, 0, 1.0, bn)
            else:
                L("t2v_gemm_tc", X, Ci, R, Ci, Wk, 5 * Ci, Co, 5 * Ci, _p(Y, 2 * Co), Co, P[pre + ".0.conv.bias"], M, Co, Ci, 5, 1,
                  Ci, 0, 0, 4, 1, 0, 0, 1.0, bn)
        else:
            ops.gemm(X, Ci, 1, Wk, 5 * Ci, 1, _p(Y, 2 * Co), Co, M, Co, 5 * Ci, 1.0, 0.0, P[pre + ".0.conv.bias"])
        Xn = _empty(R, Co, device=dev)
        p = 0.5 if training else 0.0
        mask = None if masks is None else masks[i]
        last = i == len(chans) - 2
        Xnlo = _empty(R, Co, device=dev) if (split and not split16 and not last) else None
        Xn16 = None
        if split16 and not last:
            Xn16 = (torch.empty(R, Co, device=dev, dtype=torch.int16), torch.empty(R, Co, device=dev, dtype=torch.int16))
        rnd = ops.R if (round_last or not last) else 0
        if Xn16 is not None:
            rnd = 2          # the fp32 copy sits on the fp16 grid: identical to the hi part (operand of the backward's weight gradient)
        mi = _bn_forward(ops, Y, Xn, R, Co, Tp, 2, 2 + T, B * T, P, pre + ".1", training, acts[i], mask, seed, site0 + i, p, T, dev,
                         rnd=rnd, out_lo=Xnlo, out16=Xn16)
        saved.append(dict(X=X, Y=_Saved(Y, mi), mask=mask, p=p, Ci=Ci, Co=Co, act=acts[i], W=W, X16=X16[0] if (split16 and i > 0) else None))
        X, Xlo, X16 = Xn, Xnlo, Xn16
    return X, saved


class _NoBranch(object):
    keep = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def conv_stack_backward(ops, P, prefix, dOut, saved, B, T, training, seed, site0, dev, grads, need_dx, dw_branch=None):
    """dw_branch: a _Branch on which the weight gradients (transposes + row-reduction GEMMs: off the dX chain) are enqueued;
    the caller joins it.  Tensors the branch reads are kept alive in dw_branch.keep until then."""
    Tp = T + 4
    R = B * Tp
    M = R - 4
    br = dw_branch if dw_branch is not None else _NoBranch()
    if dw_branch is not None and dw_branch.keep is None:
        dw_branch.keep = []
    for i in range(len(saved) - 1, -1, -1):
        s = saved[i]
        Ci, Co = s["Ci"], s["Co"]
        pre = "%s.%d" % (prefix, i)
        dY = _empty(R, Co, device=dev)
        _bn_backward(dOut, s["Y"], dY, R, Co, Tp, 2, 2 + T, B * T, P, pre + ".1", training, s["act"], s["mask"], seed,
                     site0 + i, s["p"], T, dev, grads, rnd=ops.R)
        grads[pre + ".0.conv.bias"] = _colsum(dY, R, Co, Tp, 2, 2 + T, dev)
        if dw_branch is not None:
            dw_branch.keep.append(dY)
        with br:
            # weight gradient in tap-major form: dWk[co, tap*Ci+ci] = sum_r dY[r+2, co] * X[r+tap, ci]
            dWk = _zeros(Co, 5 * Ci, device=dev)
            if ops.tc and ops.op16 == 1 and _DW16 and Co % 256 == 0 and Ci % 256 == 0 and Co >= 512 and M >= 16384:
                # large layers (Postnet 512 -> 512): fp16 copies of both operands (the gradient scaled by a power of two into the
                # fp16 range: same 11-bit significands as tf32), kind::f16 row-reduction GEMMs at twice the tf32 rate
                sc = _empty(4, device=dev)
                L("t2v_grad_scale", dY, R * Co, 14, sc)
                dY16 = torch.empty(R, Co, device=dev, dtype=torch.int16)
                L("t2v_cvt16_scaled", dY, dY16, R * Co, 1, sc)
                X16 = s.get("X16")            # the forward's fp16 hi part (== X) when the stack ran on fp16 pairs
                if X16 is None:
                    X16 = torch.empty(R, Ci, device=dev, dtype=torch.int16)
                    L("t2v_cvt16_scaled", s["X"], X16, R * Ci, 1, None)
                Ops.rowred16(dY16, Co, Co, X16, Ci, Ci, dWk, 5 * Ci, M, sc.data_ptr() + 4, 1, a_row0=2, b_row0=0, taps=5)
            elif ops.tc and _TAPS1 and (Ci % 256 == 0 or Ci in (64, 128)):
                # the tap is a row offset of the X operand (MN-major operands, no transposes): all five taps in one launch
                Ops.rowred(dY, Co, Co, 2, s["X"], Ci, Ci, 0, dWk, 5 * Ci, M, taps=5)
            elif ops.tc:
                for tap in range(5):
                    Ops.rowred(dY, Co, Co, 2, s["X"], Ci, Ci, tap, _p(dWk, tap * Ci), 5 * Ci, M)
            else:
                ops.gemm(_p(dY, 2 * Co), 1, Co, s["X"], 1, Ci, dWk, 5 * Ci, Co, 5 * Ci, M, 1.0, 0.0, None)
            gW = _gout(grads, pre + ".0.conv.weight", Co, Ci, 5, dev=dev)
            L("t2v_conv1d_unpack_grad", dWk, gW, Co, Ci, 5, 0.0)
            grads[pre + ".0.conv.weight"] = gW
        if i > 0 or need_dx:
            Wd = _empty(Ci, 5 * Co, device=dev)
            L("t2v_conv1d_pack", s["W"], Wd, Co, Ci, 5, 1, ops.R)
            dX = _zeros(R, Ci, device=dev)
            if ops.tc:
                L("t2v_gemm_tc", dY, Co, R, Co, Wd, 5 * Co, Ci, 5 * Co, _p(dX, 2 * Ci), Ci, None, M, Ci, Co, 5, 1, Co, 0, 0, 4,
                  1, 0, 0, 1.0, ops.bn(M, Ci))
            else:
                ops.gemm(dY, Co, 1, Wd, 5 * Co, 1, _p(dX, 2 * Ci), Ci, M, Ci, 5 * Co, 1.0, 0.0, None)
            dOut = dX
    return dOut


# ======================================================================================================= encoder
def embedding_forward(P, text, dev, rnd=0):
    """nn.Embedding gather (model.py:474,528) into padded channels-last rows [B*(Ti+4),512]; bit exact."""
    B, Ti = text.shape
    X0 = _zeros(B * (Ti + 4), 512, device=dev)
    L("t2v_embedding_fwd", text, P["transcript_embedding.weight"], X0, B, Ti, 512, P["transcript_embedding.weight"].shape[0], rnd)
    return X0


def encoder_forward(ops, P, text, in_len, training, masks, seed, dev, packed=True, X0=None, shape=None, after_convs=None):
    """Embedding + 3x(conv,BN,ReLU,dropout) + BiLSTM (model.py:151-203, 528-531).
    text [B,Ti] int64 (or X0 = already-embedded padded rows with shape=(B,Ti)).
    Returns (HoutP [B*(Ti+4),512] padded LSTM outputs, ctx)."""
    B, Ti = shape if X0 is not None else text.shape
    Tp = Ti + 4
    R = B * Tp
    if X0 is None:
        X0 = embedding_forward(P, text, dev, ops.R)
    X3, conv_saved = conv_stack_forward(ops, P, "encoder.convolutions", X0, B, Ti, [512] * 4, [1, 1, 1], training, masks,
                                        seed, SITE_ENC, dev)
    if after_convs is not None:       # hook: work the caller wants enqueued beside the BiLSTM (which leaves 84 SMs free) rather than
        after_convs()                 # beside the convolution stack
    Hh = 256
    lens = in_len if packed else None
    HoutP = _zeros(R, 512, device=dev)
    GS = _zeros(2, Ti, B, 4 * Hh, device=dev)
    CS = _zeros(2, Ti + 2, B, Hh, device=dev)          # slot t+1 = cell after time t; slots 0 and Ti+1 stay zero
    GX = []
    for d, sfx in enumerate(("", "_reverse")):
        gx = _empty(R, 4 * Hh, device=dev)
        ops.linear(X3, 512, ops.wr(P["encoder.lstm.weight_ih_l0" + sfx]), 512, gx, 4 * Hh, R, 4 * Hh, 512,
                   bias=P["encoder.lstm.bias_ih_l0" + sfx])
        GX.append(gx)
    hbuf = _zeros(2, 2, B, Hh, device=dev)             # [ping-pong][direction]
    cst = _zeros(2, B, Hh, device=dev)
    Whh = [P["encoder.lstm.weight_hh_l0"], P["encoder.lstm.weight_hh_l0_reverse"]]
    bhh = [P["encoder.lstm.bias_hh_l0"], P["encoder.lstm.bias_hh_l0_reverse"]]
    persist = B <= 64 and _BILSTM_PERSIST
    if persist:        # one resident kernel for all Ti steps of both directions (rnn_persist.cu)
        cnt = _zeros(64, device=dev, dtype=torch.int32)
        L("t2v_bilstm_seq_fwd", GX[0], GX[1], Whh[0], Whh[1], bhh[0], bhh[1], HoutP, GS, CS, hbuf, cnt, lens, B, Hh, Ti)
    for s in range(0 if persist else Ti):
        t0, t1 = s, Ti - 1 - s
        a, b = s & 1, (s + 1) & 1
        L("t2v_bilstm_step_fwd", _p(GX[0], (2 + t0) * 4 * Hh), _p(GX[1], (2 + t1) * 4 * Hh), Tp * 4 * Hh, Whh[0], Whh[1],
          bhh[0], bhh[1], hbuf[a, 0], hbuf[a, 1], hbuf[b, 0], hbuf[b, 1], cst[0], cst[1],
          _p(HoutP, (2 + t0) * 512), _p(HoutP, (2 + t1) * 512 + Hh), Tp * 512, GS[0, t0], GS[1, t1], CS[0, t0 + 1], CS[1, t1 + 1],
          lens, t0, t1, B, Hh)
    ctx = dict(B=B, Ti=Ti, text=text, in_len=lens, conv=conv_saved, X3=X3, GS=GS, CS=CS, HoutP=HoutP)
    return HoutP, ctx


def encoder_backward(ops, P, dmem, ctx, training, seed, dev, grads):
    """dmem: [B,Ti,512] gradient wrt the (compact) encoder outputs."""
    B, Ti = ctx["B"], ctx["Ti"]
    Tp, Hh = Ti + 4, 256
    R = B * Tp
    lens = ctx["in_len"]
    HoutP, GS, CS, X3 = ctx["HoutP"], ctx["GS"], ctx["CS"], ctx["X3"]
    dX3 = _zeros(R, 512, device=dev)
    DGs = [_zeros(R, 4 * Hh, device=dev), _zeros(R, 4 * Hh, device=dev)]
    WT = []
    for sfx in ("", "_reverse"):
        wt = _empty(Hh, 4 * Hh, device=dev)
        L("t2v_transpose", P["encoder.lstm.weight_hh_l0" + sfx], Hh, wt, 4 * Hh, 4 * Hh, Hh, 0)
        WT.append(wt)
    dcb = _zeros(2, B, Hh, device=dev)
    persist = B <= 64 and _BILSTM_PERSIST
    if persist:
        cnt = _zeros(64, device=dev, dtype=torch.int32)
        L("t2v_bilstm_seq_bwd", WT[0], WT[1], dmem, GS, CS, DGs[0], DGs[1], cnt, lens, B, Hh, Ti)
    for s in range(0 if persist else Ti):
        t0, t1 = Ti - 1 - s, s                    # each direction walks its own forward order backwards
        L("t2v_bilstm_step_bwd", _p(DGs[0], (3 + t0) * 4 * Hh) if s > 0 else None, _p(DGs[1], (1 + t1) * 4 * Hh) if s > 0 else None,
          Tp * 4 * Hh, WT[0], WT[1], _p(dmem, t0 * 512), _p(dmem, t1 * 512 + Hh), Ti * 512, dcb[0], dcb[1], GS[0, t0], GS[1, t1],
          CS[0, t0 + 1], CS[1, t1 + 1], CS[0, t0], CS[1, t1 + 2], _p(DGs[0], (2 + t0) * 4 * Hh), _p(DGs[1], (2 + t1) * 4 * Hh),
          lens, t0, t1, B, Hh)
    if ops.tc:      # operands of the tensor-core weight-gradient GEMMs go onto the tf32 grid (tcgen05 would truncate)
        L("t2v_round_tf32", HoutP, HoutP.numel())
    for d, sfx in enumerate(("", "_reverse")):
        DG = DGs[d]
        gb = _colsum(DG, R, 4 * Hh, 1, 0, 1, dev)
        if ops.tc:
            L("t2v_round_tf32", DG, DG.numel())
        # batched weight grads; h_prev of row r is HoutP[r -/+ 1] (zero pad rows make the boundaries right)
        gWhh = _zeros(4 * Hh, Hh, device=dev)
        if d == 0:
            ops.linear_dw(_p(DG, 4 * Hh), 4 * Hh, HoutP, 512, gWhh, Hh, R - 1, 4 * Hh, Hh, device=dev)
        else:
            ops.linear_dw(DG, 4 * Hh, _p(HoutP, 512 + Hh), 512, gWhh, Hh, R - 1, 4 * Hh, Hh, device=dev)
        grads["encoder.lstm.weight_hh_l0" + sfx] = gWhh
        gWih = _zeros(4 * Hh, 512, device=dev)
        ops.linear_dw(DG, 4 * Hh, X3, 512, gWih, 512, R, 4 * Hh, 512, device=dev)
        grads["encoder.lstm.weight_ih_l0" + sfx] = gWih
        grads["encoder.lstm.bias_ih_l0" + sfx] = gb
        grads["encoder.lstm.bias_hh_l0" + sfx] = _clone(gb)
        ops.linear_dx(DG, 4 * Hh, P["encoder.lstm.weight_ih_l0" + sfx], 512, dX3, 512, R, 4 * Hh, 512, accumulate=(d == 1))
    dX0 = conv_stack_backward(ops, P, "encoder.convolutions", dX3, ctx["conv"], B, Ti, training, seed, SITE_ENC, dev, grads,
                              need_dx=True)
    gE = _zeros(*P["transcript_embedding.weight"].shape, device=dev)
    L("t2v_embedding_bwd", ctx["text"], dX0, gE, B, Ti, 512)
    grads["transcript_embedding.weight"] = gE


# ======================================================================================================= reference encoder
_REF = "vae_gst.ref_encoder."


def refenc_forward(ops, P, mel, training, dev):
    """mel [N,80,T] contiguous, reinterpreted as NHWC [N,T,80,1] (the reference's .view, quirk Q1);
    CoordConv + 6x(conv3x3 s2, BN2d, ReLU) + GRU -> last hidden [N,256] (modules.py:65-80)."""
    N, n_mel, T = mel.shape
    Hc, Wc, Ci = T, n_mel, 1
    x = mel
    filters = [P[_REF + "convs.%d.%s" % (i, "conv.weight" if i == 0 else "weight")].shape[0] for i in range(6)]
    layers = []
    for i in range(6):
        Ho, Wo = (Hc - 1) // 2 + 1, (Wc - 1) // 2 + 1
        Ct = 4 if i == 0 else Ci
        Co = filters[i]
        rows = N * Ho * Wo
        # The reference encoder's output (mu, logvar -> z -> style) is added to EVERY encoder position, so its error does not
        # average out in the attention context: a 3e-4 style error alone costs 3.6e-4 on the mel frames (measured with the oracle).
        # Its forward convolutions are 11.6 GFLOP at C3 (0.5 % of the step): they stay on the exact FFMA GEMM.
        exact = True
        rl = 0 if exact else 1
        col = _empty(rows, 9 * Ct, device=dev)
        L("t2v_im2col_3x3s2", x, col, N, Hc, Wc, Ci, 1 if i == 0 else 0, rl)
        wname = _REF + ("convs.0.conv" if i == 0 else "convs.%d" % i)
        Wk = _empty(Co, 9 * Ct, device=dev)
        L("t2v_conv1d_pack", P[wname + ".weight"], Wk, Co, Ct, 9, 0, rl)
        Y = _empty(rows, Co, device=dev)
        ops.linear(col, 9 * Ct, Wk, 9 * Ct, Y, Co, rows, Co, 9 * Ct, bias=P[wname + ".bias"], force_exact=exact)
        Xn = _empty(rows, Co, device=dev)
        mi = _bn_forward(ops, Y, Xn, rows, Co, 1, 0, 1, rows, P, _REF + "bns.%d" % i, training, 1, None, 0, 0, 0.0, 1, dev)
        layers.append(dict(col=col, Wk=Wk, Y=_Saved(Y, mi), rows=rows, Ct=Ct, Co=Co, H=Hc, W=Wc, Ci=Ci, wname=wname, exact=exact))
        x, Hc, Wc, Ci = Xn, Ho, Wo, Co
    Tq, Wq, Cq = Hc, Wc, Ci                       # GRU sequence length, remaining mel bins, channels
    Fin = Wq * Cq
    Hh = P[_REF + "gru.weight_hh_l0"].shape[1]
    # reference feature order is c*W'+w (modules.py:73-76); ours is w*C+c -> permute the input weight columns
    Wih = _empty(3 * Hh, Fin, device=dev)
    L("t2v_conv1d_pack", P[_REF + "gru.weight_ih_l0"], Wih, 3 * Hh, Cq, Wq, 0, 0)
    GI = _empty(N * Tq, 3 * Hh, device=dev)
    ops.linear(x, Fin, Wih, Fin, GI, 3 * Hh, N * Tq, 3 * Hh, Fin, force_exact=True)
    HS = _zeros(Tq + 1, N, Hh, device=dev)
    SV = _empty(Tq, N, 4 * Hh, device=dev)
    gh = _empty(N, 3 * Hh, device=dev)
    persist = N <= 64 and _BILSTM_PERSIST
    if persist:        # all Tq steps in one resident kernel (rnn_persist.cu)
        cnt = _zeros(32, device=dev, dtype=torch.int32)
        L("t2v_gru_seq_fwd", GI, Tq * 3 * Hh, P[_REF + "gru.weight_hh_l0"], P[_REF + "gru.bias_ih_l0"], P[_REF + "gru.bias_hh_l0"],
          HS, SV, cnt, N, Hh, Tq)
    for t in range(0 if persist else Tq):
        ops.gemm(HS[t], Hh, 1, P[_REF + "gru.weight_hh_l0"], Hh, 1, gh, 3 * Hh, N, 3 * Hh, Hh)
        L("t2v_gru_pointwise_fwd", _p(GI, t * 3 * Hh), Tq * 3 * Hh, gh, P[_REF + "gru.bias_ih_l0"], P[_REF + "gru.bias_hh_l0"],
          HS[t], HS[t + 1], SV[t], N, Hh)
    ctx = dict(N=N, layers=layers, X6=x, Tq=Tq, Wq=Wq, Cq=Cq, Fin=Fin, Hh=Hh, Wih=Wih, HS=HS, SV=SV)
    return HS[Tq], ctx


def refenc_backward(ops, P, dh_last, ctx, training, dev, grads):
    N, Tq, Hh, Fin = ctx["N"], ctx["Tq"], ctx["Hh"], ctx["Fin"]
    HS, SV = ctx["HS"], ctx["SV"]
    Whh = P[_REF + "gru.weight_hh_l0"]
    DGI = _empty(N * Tq, 3 * Hh, device=dev)
    dgh = _empty(N, 3 * Hh, device=dev)
    dh = _clone(dh_last)
    dhp = _empty(N, Hh, device=dev)
    gWhh = _zeros(3 * Hh, Hh, device=dev)
    bh_acc = _zeros(3 * Hh, device=dev, dtype=torch.float64)
    persist = N <= 64 and _BILSTM_PERSIST
    # the GRU's three backward GEMMs (dW_hh, dW_ih, dX) have few output tiles and a long reduction: on the exact FFMA kernel they
    # are 6 .. 14 CTAs walking 400 .. 800 k-steps each (~0.2 ms apiece, on the chain that ends the step).  In the tensor-core modes
    # they run as tf32 GEMMs over rounded copies of the (small) operands; the forward stays exact (DESIGN.md "Precision modes").
    fast = ops.tc and _GRU_BWD_TC and Tq * N >= 256

    def r32(x):
        y = _clone(x.contiguous())
        L("t2v_round_tf32", y, y.numel())
        return y
    if persist:
        DGH = _empty(Tq, N, 3 * Hh, device=dev)
        cnt = _zeros(32, device=dev, dtype=torch.int32)
        L("t2v_gru_seq_bwd", Whh, dh_last, SV, HS, DGI, Tq * 3 * Hh, DGH, cnt, N, Hh, Tq)
        if fast:
            ops.linear_dw(r32(DGH), 3 * Hh, r32(HS[:Tq]), Hh, gWhh, Hh, Tq * N, 3 * Hh, Hh, device=dev)
        else:
            ops.linear_dw(DGH, 3 * Hh, HS, Hh, gWhh, Hh, Tq * N, 3 * Hh, Hh, device=dev, force_exact=True)   # sum_t dgh[t]^T h[t-1]
        L("t2v_col_stats", DGH, Tq * N, 3 * Hh, 1, 0, 1, 2, bh_acc, None)
    for t in range(-1 if persist else Tq - 1, -1, -1):
        L("t2v_gru_pointwise_bwd", dh, SV[t], HS[t], _p(DGI, t * 3 * Hh), Tq * 3 * Hh, dgh, dhp, N, Hh, 0)
        ops.gemm(dgh, 3 * Hh, 1, Whh, 1, Hh, dhp, Hh, N, Hh, 3 * Hh, 1.0, 1.0)          # dh_prev += dgh @ W_hh
        ops.gemm(dgh, 1, 3 * Hh, HS[t], 1, Hh, gWhh, Hh, 3 * Hh, Hh, N, 1.0, 1.0)        # dW_hh += dgh^T h_prev
        L("t2v_col_stats", dgh, N, 3 * Hh, 1, 0, 1, 2, bh_acc, None)
        dh, dhp = dhp, dh
    grads[_REF + "gru.weight_hh_l0"] = gWhh
    gbh = _empty(3 * Hh, device=dev)
    L("t2v_double_to_float", bh_acc, gbh, 3 * Hh, 0.0)
    grads[_REF + "gru.bias_hh_l0"] = gbh
    grads[_REF + "gru.bias_ih_l0"] = _colsum(DGI, N * Tq, 3 * Hh, 1, 0, 1, dev)
    gWihp = _zeros(3 * Hh, Fin, device=dev)
    DGIr = r32(DGI) if fast else DGI
    if fast:
        ops.linear_dw(DGIr, 3 * Hh, r32(ctx["X6"]), Fin, gWihp, Fin, N * Tq, 3 * Hh, Fin, device=dev)
    else:
        ops.linear_dw(DGI, 3 * Hh, ctx["X6"], Fin, gWihp, Fin, N * Tq, 3 * Hh, Fin, device=dev, force_exact=True)
    gWih = _empty(3 * Hh, Fin, device=dev)
    L("t2v_conv1d_unpack_grad", gWihp, gWih, 3 * Hh, ctx["Cq"], ctx["Wq"], 0.0)
    grads[_REF + "gru.weight_ih_l0"] = gWih
    dX = _empty(N * Tq, Fin, device=dev)
    if fast:
        ops.linear_dx(DGIr, 3 * Hh, r32(ctx["Wih"]), Fin, dX, Fin, N * Tq, 3 * Hh, Fin)
    else:
        ops.linear_dx(DGI, 3 * Hh, ctx["Wih"], Fin, dX, Fin, N * Tq, 3 * Hh, Fin, force_exact=True)
    for i in range(5, -1, -1):
        ly = ctx["layers"][i]
        rows, Co, Ct = ly["rows"], ly["Co"], ly["Ct"]
        dY = _empty(rows, Co, device=dev)
        _bn_backward(dX, ly["Y"], dY, rows, Co, 1, 0, 1, rows, P, _REF + "bns.%d" % i, training, 1, None, 0, 0, 0.0, 1, dev, grads,
                     rnd=ops.R)
        grads[ly["wname"] + ".bias"] = _colsum(dY, rows, Co, 1, 0, 1, dev)
        dWk = _zeros(Co, 9 * Ct, device=dev)
        if ops.tc and rows >= 256:      # the forward kept the patches exact; the tensor-core dW GEMM wants them on the tf32 grid
            L("t2v_round_tf32", ly["col"], ly["col"].numel())
        ops.linear_dw(dY, Co, ly["col"], 9 * Ct, dWk, 9 * Ct, rows, Co, 9 * Ct, device=dev)
        gW = torch.empty_like(P[ly["wname"] + ".weight"])
        L("t2v_conv1d_unpack_grad", dWk, gW, Co, Ct, 9, 0.0)
        grads[ly["wname"] + ".weight"] = gW
        if i > 0:
            dcol = _empty(rows, 9 * Ct, device=dev)
            ops.linear_dx(dY, Co, ly["Wk"], 9 * Ct, dcol, 9 * Ct, rows, Co, 9 * Ct)
            dX = _empty(N * ly["H"] * ly["W"], ly["Ci"], device=dev)
            L("t2v_col2im_3x3s2", dcol, dX, N, ly["H"], ly["W"], ly["Ci"])


def vae_forward(ops, P, mel, training, eps, dev):
    """VAE_GST.forward (modules.py:24-31): returns style [N,512], mulv [N,64] = [mu|logvar], z [N,32], ctx."""
    h, rctx = refenc_forward(ops, P, mel, training, dev)
    N = mel.shape[0]
    Z = P["vae_gst.fc1.weight"].shape[0]
    Hh = h.shape[1]
    Wmulv = torch.cat((P["vae_gst.fc1.weight"], P["vae_gst.fc2.weight"]), 0).contiguous()      # pointer plumbing only
    bmulv = torch.cat((P["vae_gst.fc1.bias"], P["vae_gst.fc2.bias"]), 0).contiguous()
    mulv = _empty(N, 2 * Z, device=dev)
    ops.linear(h, Hh, Wmulv, Hh, mulv, 2 * Z, N, 2 * Z, Hh, bias=bmulv, force_exact=True)
    z = _empty(N, Z, device=dev)
    L("t2v_vae_reparam_fwd", mulv, eps, z, N, Z, 1 if training else 0)
    E = P["vae_gst.fc3.weight"].shape[0]
    style = _empty(N, E, device=dev)
    ops.linear(z, Z, P["vae_gst.fc3.weight"], Z, style, E, N, E, Z, bias=P["vae_gst.fc3.bias"], force_exact=True)
    ctx = dict(N=N, Z=Z, Hh=Hh, E=E, h=h, Wmulv=Wmulv, mulv=mulv, z=z, eps=eps, ref=rctx, training=training)
    return style, mulv, z, ctx


def fc3_forward(ops, P, z, dev):
    """model.vae_gst.fc3(z) (inference.ipynb cell 22 / synthesizer.py:131)."""
    N, Z = z.shape
    E = P["vae_gst.fc3.weight"].shape[0]
    style = _empty(N, E, device=dev)
    ops.linear(z.contiguous(), Z, P["vae_gst.fc3.weight"], Z, style, E, N, E, Z, bias=P["vae_gst.fc3.bias"], force_exact=True)
    return style


def vae_backward(ops, P, dstyle, dmu, dlogvar, ctx, dev, grads):
    N, Z, Hh, E = ctx["N"], ctx["Z"], ctx["Hh"], ctx["E"]
    dz = _empty(N, Z, device=dev)
    ops.gemm(dstyle, E, 1, P["vae_gst.fc3.weight"], 1, Z, dz, Z, N, Z, E)
    g3 = _empty(E, Z, device=dev)
    ops.gemm(dstyle, 1, E, ctx["z"], 1, Z, g3, Z, E, Z, N)
    grads["vae_gst.fc3.weight"] = g3
    grads["vae_gst.fc3.bias"] = _colsum(dstyle, N, E, 1, 0, 1, dev)
    dmulv = _empty(N, 2 * Z, device=dev)
    L("t2v_vae_reparam_bwd", ctx["mulv"], ctx["eps"], dz, dmu, dlogvar, dmulv, N, Z, 1 if ctx["training"] else 0)
    gW = _empty(2 * Z, Hh, device=dev)
    ops.gemm(dmulv, 1, 2 * Z, ctx["h"], 1, Hh, gW, Hh, 2 * Z, Hh, N)
    gb = _colsum(dmulv, N, 2 * Z, 1, 0, 1, dev)
    grads["vae_gst.fc1.weight"], grads["vae_gst.fc2.weight"] = gW[:Z].contiguous(), gW[Z:].contiguous()
    grads["vae_gst.fc1.bias"], grads["vae_gst.fc2.bias"] = gb[:Z].contiguous(), gb[Z:].contiguous()
    dh = _empty(N, Hh, device=dev)
    ops.gemm(dmulv, 2 * Z, 1, ctx["Wmulv"], 1, Hh, dh, Hh, N, Hh, 2 * Z)
    refenc_backward(ops, P, dh, ctx["ref"], ctx["training"], dev, grads)


# ======================================================================================================= decoder
_D = "decoder."
_A = "decoder.attention_layer."


def conv_weight_T(P, dev):
    """location conv weight [32,2,31] -> [2*31, 32] (what attention2.cu stages into shared memory)"""
    wt = _empty(62, 32, device=dev)
    L("t2v_transpose", P[_A + "location_layer.location_conv.conv.weight"], 62, wt, 32, 32, 62, 0)
    return wt


def pack_decoder_weights(P, dev, R=0):
    Wa = _empty(4096, 1792, device=dev)
    L("t2v_copy2d", P[_D + "attention_rnn.weight_ih"], 768, 1, Wa, 1792, 4096, 768, 0.0, R)
    L("t2v_copy2d", P[_D + "attention_rnn.weight_hh"], 1024, 1, _p(Wa, 768), 1792, 4096, 1024, 0.0, R)
    Wd = _empty(4096, 2560, device=dev)
    L("t2v_copy2d", P[_D + "decoder_rnn.weight_ih"], 1536, 1, Wd, 2560, 4096, 1536, 0.0, R)
    L("t2v_copy2d", P[_D + "decoder_rnn.weight_hh"], 1024, 1, _p(Wd, 1536), 2560, 4096, 1024, 0.0, R)
    Wpg = _empty(81, 1536, device=dev)
    L("t2v_copy2d", P[_D + "linear_projection.linear_layer.weight"], 1536, 1, Wpg, 1536, 80, 1536, 0.0, R)
    L("t2v_copy2d", P[_D + "gate_layer.linear_layer.weight"], 1536, 1, _p(Wpg, 80 * 1536), 1536, 1, 1536, 0.0, R)
    bpg = _empty(81, device=dev)
    L("t2v_copy2d", P[_D + "linear_projection.linear_layer.bias"], 80, 1, bpg, 80, 1, 80, 0.0, 0)
    L("t2v_copy2d", P[_D + "gate_layer.linear_layer.bias"], 1, 1, _p(bpg, 80), 1, 1, 1, 0.0, 0)
    return Wa, Wd, Wpg, bpg


def pack_projection_lo(ops, P, W, dev):
    """W_lo = tf32(W - W_hi) of [linear_projection ; gate_layer] for the split projection"""
    if not (ops.tc and ops.split):
        return
    Wf = _empty(81, 1536, device=dev)
    L("t2v_copy2d", P[_D + "linear_projection.linear_layer.weight"], 1536, 1, Wf, 1536, 80, 1536, 0.0, 0)
    L("t2v_copy2d", P[_D + "gate_layer.linear_layer.weight"], 1536, 1, _p(Wf, 80 * 1536), 1536, 1, 1536, 0.0, 0)
    W["Wpg_x"] = Wf                         # exact copy (the in-kernel projection of the free-running loop uses it)
    W["Wpg_lo"] = ops.lo(Wf, W["Wpg"])


def _fill_seq_struct(S, ops, P, W, B, Ti, To, training, seed, drop_masks, mask_value, in_len, mem, pmem, buf):
    S.B, S.Ti, S.To = B, Ti, To
    S.use_tc = 1 if ops.tc else 0
    S.training = 1 if training else 0
    S.p_att, S.p_dec = float(getattr(ops, "p_att", 0.1)), float(getattr(ops, "p_dec", 0.1))
    S.seed = seed
    S.drop_masks = _lib.ptr(drop_masks)
    S.mask_value = mask_value
    S.in_lens = _lib.ptr(in_len)
    S.Wa, S.Wd = W["Wa"].data_ptr(), W["Wd"].data_ptr()
    S.ba1, S.ba2 = P[_D + "attention_rnn.bias_ih"].data_ptr(), P[_D + "attention_rnn.bias_hh"].data_ptr()
    S.bd1, S.bd2 = P[_D + "decoder_rnn.bias_ih"].data_ptr(), P[_D + "decoder_rnn.bias_hh"].data_ptr()
    S.WaP, S.WdP = _lib.ptr(W.get("WaP")), _lib.ptr(W.get("WdP"))
    S.Wq = W["Wq"].data_ptr()
    S.Wconv = W["WconvT"].data_ptr()          # [2*31, 32]: taps-major transpose of the location conv weight
    S.Wloc = P[_A + "location_layer.location_dense.linear_layer.weight"].data_ptr()
    S.v = P[_A + "v.linear_layer.weight"].data_ptr()
    S.mem, S.pmem = mem.data_ptr(), pmem.data_ptr()
    for k in ("XA", "XD", "CA", "CD", "CUM", "align", "GA", "GD", "CPA", "CPD", "ASAVE", "parts", "qparts", "ebuf", "HCHI", "HCLO",
              "XA16", "XD16"):
        setattr(S, k, _lib.ptr(buf.get(k)))
    S.op16 = ops.op16 if buf.get("XA16") is not None else 0
    S.WaP16, S.WdP16 = _lib.ptr(W.get("WaP16")), _lib.ptr(W.get("WdP16"))
    if S.op16:       # fp16 copy of the encoder memory for the context reduction of the persistent loop (exact: mem is on the tf32 grid)
        buf["mem16"] = torch.empty(B * Ti, 512, device=mem.device, dtype=torch.int16)
        L("t2v_cvt16_2d", mem, 512, buf["mem16"], 512, B * Ti, 512, 1)
    S.mem16 = _lib.ptr(buf.get("mem16"))


def alloc_decoder_buffers(B, Ti, To, dev, save=True, op16=0, split=False):
    buf = dict(XA=_zeros((To + 1) * B, 1792, device=dev), XD=_zeros((To + 1) * B, 2560, device=dev),
               CA=_zeros((To + 1) * B, 1024, device=dev), CD=_zeros((To + 1) * B, 1024, device=dev),
               CUM=_zeros((To + 1) * B, Ti, device=dev), align=_zeros(B, To, Ti, device=dev),
               parts=_empty(32 * B * 4096, device=dev), qparts=_empty(8 * B * 128, device=dev),
               ebuf=_empty(B * Ti, 129, device=dev))
    if save:
        buf.update(GA=_empty(To * B, 4096, device=dev), GD=_empty(To * B, 4096, device=dev),
                   CPA=_empty(To * B, 1024, device=dev), CPD=_empty(To * B, 1024, device=dev),
                   ASAVE=_empty(To * B * Ti, 128, device=dev))
    if op16:        # 16-bit operand copies of XA / XD for the persistent loop (zero rows = the initial h / ctx / go frame)
        buf.update(XA16=_zeros((To + 1) * B, 1792, device=dev, dtype=torch.int16),
                   XD16=_zeros((To + 1) * B, 2560, device=dev, dtype=torch.int16))
    if split:       # [h_dec_t | ctx_t] as hi + lo parts: the operands of the split mel / gate projection
        buf["HCHI"] = _empty(To * B, 1536, device=dev)
        buf["HCLO"] = _empty(To * B, 1536, device=dev)
    return buf


def pack_step_weights(ops, W, dev):
    """tile-contiguous copies of Wa / Wd for the persistent loop kernels (decoder_persist.cu): fp32 tiles (tf32 math) and, in the
    fp16 / bf16 modes, 16-bit tiles made from the same (already rounded) fp32 matrices"""
    if not ops.tc:
        return
    W["WaP"], W["WdP"] = _empty(4096, 1792, device=dev), _empty(4096, 2560, device=dev)
    L("t2v_pack_step_tiles", W["Wa"], 0, W["WaP"])
    L("t2v_pack_step_tiles", W["Wd"], 1, W["WdP"])
    if ops.op16:
        W["WaP16"] = torch.empty(4096, 1792, device=dev, dtype=torch.int16)
        W["WdP16"] = torch.empty(4096, 2560, device=dev, dtype=torch.int16)
        L("t2v_pack_step_tiles16", W["Wa"], 0, W["WaP16"], ops.op16)
        L("t2v_pack_step_tiles16", W["Wd"], 1, W["WdP16"], ops.op16)


def project_mel_gate(ops, W, buf, O, B, To, persistent):
    """deferred linear_projection + gate_layer of [h_dec_t | ctx_t] for all steps (model.py:383-388) -> O [To*B, 84].
    The mel frames are sums with heavy cancellation: plain tf32 rounding of this one GEMM was 90 % of the mel error of the whole
    decoder.  When the persistent loop kernel ran it left [h_dec_t | ctx_t] as hi + lo parts (HCHI / HCLO) and the projection is
    ONE split GEMM x_hi W_hi + x_lo W_hi + x_hi W_lo; otherwise (per-step launches) two plain GEMMs over the XD rows: h_dec_t sits in
    XD row t+1 (cols 1536..), ctx_t in XD row t (cols 1024..1535)."""
    n = To * B
    XD = buf["XD"]
    if persistent and ops.tc and ops.split and buf.get("HCHI") is not None and W.get("Wpg_lo") is not None:
        L("t2v_gemm_tc_split3", buf["HCHI"], buf["HCLO"], 1536, n, 1536, W["Wpg"], W["Wpg_lo"], 1536, 81, 1536, O, 84, W["bpg"],
          n, 81, 1536, 1, 0, 0, 0, 0, 1.0, 128)
        return
    ops.linear(_p(XD, B * 2560 + 1536), 2560, W["Wpg"], 1536, O, 84, n, 81, 1024, bias=W["bpg"], a_rows=n)
    ops.linear(_p(XD, 1024), 2560, _p(W["Wpg"], 1024), 1536, O, 84, n, 81, 512, accumulate=True, a_rows=n)


def decoder_prepare(ops, P, mel_tgt, B, Ti, prenet_masks, seed, dev):
    """Everything of Decoder.forward that does not need the encoder outputs: weight packing, the sequence buffers and the prenet over
    the go frame + all teacher frames (model.py:406-409; dropout always on, model.py:101).  The train step runs it on a side branch
    beside the encoders."""
    To = mel_tgt.shape[2]
    W = {}
    W["Wa"], W["Wd"], W["Wpg"], W["bpg"] = pack_decoder_weights(P, dev, ops.R)
    W["Wq"] = ops.wr(P[_A + "query_layer.linear_layer.weight"])
    W["WconvT"] = conv_weight_T(P, dev)
    W["Wm"] = ops.wr(P[_A + "memory_layer.linear_layer.weight"])
    pack_step_weights(ops, W, dev)
    pack_projection_lo(ops, P, W, dev)
    buf = alloc_decoder_buffers(B, Ti, To, dev, save=True, op16=ops.op16, split=ops.split)
    Fr = _empty((To + 1) * B, 80, device=dev)
    L("t2v_bct_to_rows_tb_shift", mel_tgt, Fr, B, 80, To, ops.R)
    n = (To + 1) * B
    P1pre = _empty(n, 256, device=dev)
    ops.linear(Fr, 80, ops.wr(P[_D + "prenet.layers.0.linear_layer.weight"]), 80, P1pre, 256, n, 256, 80)
    P1 = _empty(n, 256, device=dev)
    m0 = None if prenet_masks is None else prenet_masks[0]
    m1 = None if prenet_masks is None else prenet_masks[1]
    L("t2v_relu_drop_fwd", P1pre, P1, 256, n, 256, m0, seed, SITE_PRENET, 0.5, 0, ops.R)
    P2pre = _empty(n, 256, device=dev)
    ops.linear(P1, 256, ops.wr(P[_D + "prenet.layers.1.linear_layer.weight"]), 256, P2pre, 256, n, 256, 256)
    L("t2v_relu_drop_fwd", P2pre, buf["XA"], 1792, n, 256, m1, seed, SITE_PRENET + 1, 0.5, 0, ops.RX)
    if ops.op16:
        L("t2v_cvt16_2d", buf["XA"], 1792, buf["XA16"], 1792, n, 256, ops.op16)
    return dict(W=W, buf=buf, Fr=Fr, P1pre=P1pre, P1=P1, P2pre=P2pre, m0=m0, m1=m1)


def decoder_forward(ops, P, memory, mel_tgt, in_len, training, prenet_masks, drop_masks, seed, mask_value, dev, prep=None):
    """Teacher-forced Decoder.forward (model.py:391-426).  memory [B,Ti,512]; mel_tgt [B,80,To].
    Returns O [To*B,84] (mel|gate rows, time-major), alignments [B,To,Ti], ctx."""
    B, Ti, _ = memory.shape
    To = mel_tgt.shape[2]
    if prep is None:
        prep = decoder_prepare(ops, P, mel_tgt, B, Ti, prenet_masks, seed, dev)
    W, buf = prep["W"], prep["buf"]
    Fr, P1pre, P1, P2pre, m0, m1 = (prep[k] for k in ("Fr", "P1pre", "P1", "P2pre", "m0", "m1"))
    pmem = _empty(B * Ti, 128, device=dev)
    ops.linear(memory, 512, W["Wm"], 512, pmem, 128, B * Ti, 128, 512)
    S = _lib.T2VDecoderSeq()
    _fill_seq_struct(S, ops, P, W, B, Ti, To, training, seed, drop_masks, mask_value, in_len, memory, pmem, buf)
    _trace("  fwd prenet+pack")
    L("t2v_decoder_fwd_steps", S, 0, To)
    persistent = bool(_lib.lib().t2v_decoder_last_path())      # 1: the persistent loop kernel was enqueued (HCHI / HCLO are valid)
    _trace("  fwd decoder loop")
    # deferred mel/gate projection of [h_dec_t | ctx_t] for all steps (model.py:383-388)
    O = _zeros(To * B, 84, device=dev)
    project_mel_gate(ops, W, buf, O, B, To, persistent)
    ctx = dict(B=B, Ti=Ti, To=To, W=W, pmem=pmem, memory=memory, buf=buf, S=S, Fr=Fr, P1pre=P1pre, P1=P1, P2pre=P2pre,
               m0=m0, m1=m1, seed=seed, O=O)
    return O, buf["align"], ctx


def decoder_backward(ops, P, dO, ctx, dev, grads):
    """dO [To*B,84] grad wrt the mel|gate rows.  Returns (dmemory [B,Ti,512], the weight-gradient branch to join)."""
    B, Ti, To, W, buf = ctx["B"], ctx["Ti"], ctx["To"], ctx["W"], ctx["buf"]
    n = To * B
    XA, XD = buf["XA"], buf["XD"]
    # projection backward
    DHC = _empty(n, 1536, device=dev)
    ops.linear_dx(dO, 84, W["Wpg"], 1536, DHC, 1536, n, 81, 1536)
    gWpg = _zeros(81, 1536, device=dev)
    ops.linear_dw(dO, 84, _p(XD, B * 2560 + 1536), 2560, gWpg, 1536, n, 81, 1024, device=dev)
    ops.linear_dw(dO, 84, _p(XD, 1024), 2560, _p(gWpg, 1024), 1536, n, 81, 512, device=dev)
    gbpg = _colsum(dO, n, 84, 1, 0, 1, dev)
    grads[_D + "linear_projection.linear_layer.weight"] = gWpg[:80].contiguous()
    grads[_D + "gate_layer.linear_layer.weight"] = gWpg[80:81].contiguous()
    grads[_D + "linear_projection.linear_layer.bias"] = gbpg[:80].contiguous()
    grads[_D + "gate_layer.linear_layer.bias"] = gbpg[80:81].contiguous()
    # reverse time loop
    Wq = W["Wq"]
    WaT = _empty(1792, 4096, device=dev)
    WdT = _empty(2560, 4096, device=dev)
    WqT = _empty(1024, 128, device=dev)
    L("t2v_transpose", W["Wa"], 1792, WaT, 4096, 4096, 1792, 0)
    L("t2v_transpose", W["Wd"], 2560, WdT, 4096, 4096, 2560, 0)
    L("t2v_transpose", Wq, 1024, WqT, 128, 128, 1024, 0)
    WaTP = WdTP = None
    if ops.tc:      # tile-contiguous copies for the persistent reverse loop (decoder_persist_bwd.cu)
        WaTP, WdTP = _empty(1792, 4096, device=dev), _empty(2560, 4096, device=dev)
        L("t2v_pack_step_tiles", WaT, 2, WaTP)
        L("t2v_pack_step_tiles", WdT, 3, WdTP)
    _trace("  bwd proj")
    nck = int(_lib.lib().t2v_attn2_chunks(Ti))
    D = _lib.T2VDecoderBwd()
    _ctypes.memmove(_ctypes.addressof(D.f), _ctypes.addressof(ctx["S"]), _ctypes.sizeof(_lib.T2VDecoderSeq))
    t = dict(WaT=WaT, WdT=WdT, WqT=WqT, DHC=DHC, DGA=_empty(n, 4096, device=dev), DGD=_empty(n, 4096, device=dev),
             DXA=_empty(n, 1792, device=dev), DXD=_empty(n, 2560, device=dev), dCa=_zeros(B, 1024, device=dev),
             dCd=_zeros(B, 1024, device=dev), dwprev=_zeros(2 * B, Ti, device=dev), gcum=_zeros(2 * B, Ti, device=dev),
             dpmem=_zeros(B * Ti, 128, device=dev), DCTX=_empty(n, 512, device=dev), dw_part=_empty(5 * B, Ti, device=dev),
             DQ=_zeros(n, 128, device=dev), dHq=_empty(B, 1024, device=dev), dv_part=_zeros(B * nck, 128, device=dev),
             dwloc_part=_zeros(B * nck, 128 * 32, device=dev), dwconv_part=_zeros(B * nck, 32 * 2 * 31, device=dev))
    for k, v in t.items():
        setattr(D, k, v.data_ptr())
    D.WaTP, D.WdTP = _lib.ptr(WaTP), _lib.ptr(WdTP)
    t["_packed"] = (WaTP, WdTP)
    D.op16 = 0
    if ops.tc and ops.op16 and _BWD16:
        # fp16 copies of the W^T tiles and of the gate gradients for the two dX GEMMs of the loop (kind::f16).  Gradients are
        # ~1e-8: a power-of-two scale derived from max |dO| keeps them inside the fp16 range (saturating conversion), the dX rows
        # are unscaled in fp32 -- same 11-bit significand as the tf32 path, half the L2 -> SM bytes that bound these GEMMs.
        t["scale"] = _zeros(4, device=dev)
        L("t2v_grad_scale", dO, dO.numel(), 12, t["scale"])
        t["WaTP16"] = torch.empty(1792, 4096, device=dev, dtype=torch.int16)
        t["WdTP16"] = torch.empty(2560, 4096, device=dev, dtype=torch.int16)
        L("t2v_pack_step_tiles16", WaT, 2, t["WaTP16"], 1)
        L("t2v_pack_step_tiles16", WdT, 3, t["WdTP16"], 1)
        t["DGA16"] = torch.empty(n, 4096, device=dev, dtype=torch.int16)
        t["DGD16"] = torch.empty(n, 4096, device=dev, dtype=torch.int16)
        D.op16 = 1
        D.DGA16, D.DGD16 = t["DGA16"].data_ptr(), t["DGD16"].data_ptr()
        D.WaTP16, D.WdTP16 = t["WaTP16"].data_ptr(), t["WdTP16"].data_ptr()
        D.dg_scale = t["scale"].data_ptr()
    t["gba"], t["gbd"] = _zeros(4096, device=dev), _zeros(4096, device=dev)
    D.gb_att, D.gb_dec = t["gba"].data_ptr(), t["gbd"].data_ptr()
    L("t2v_decoder_bwd_steps", D, To, 0)
    persistent = bool(_lib.lib().t2v_decoder_last_bwd_path())
    dw16 = bool(D.op16) and ops.op16 == 1 and bool(_lib.lib().t2v_decoder_last_bwd_path()) and buf.get("XA16") is not None and _DW16
    _trace("  bwd decoder loop")
    if ops.tc:
        L("t2v_round_tf32", t["dpmem"], t["dpmem"].numel())
    # every weight gradient of the decoder is off the path to d(memory): a side branch, joined by the caller after the
    # encoder backward
    br = _Branch(0)
    br.keep = (t, D)                                     # the branch reads these after this function returns
    with br:
        # batched weight gradients over all steps
        if dw16:     # the two big weight gradients straight from the fp16 copies the loops left behind (gradient copy x 1 / scale);
            # each GEMM writes its [weight_ih | weight_hh] column blocks into the two parameters' own gradient tensors
            inv = t["scale"].data_ptr() + 4
            ga_ih = _gout(grads, _D + "attention_rnn.weight_ih", 4096, 768, dev=dev, zero=True)
            ga_hh = _gout(grads, _D + "attention_rnn.weight_hh", 4096, 1024, dev=dev, zero=True)
            gd_ih = _gout(grads, _D + "decoder_rnn.weight_ih", 4096, 1536, dev=dev, zero=True)
            gd_hh = _gout(grads, _D + "decoder_rnn.weight_hh", 4096, 1024, dev=dev, zero=True)
            Ops.rowred16(t["DGA16"], 4096, 4096, buf["XA16"], 1792, 1792, ga_ih, 768, n, inv, D2=ga_hh, ldd2=1024, n_split=768)
            Ops.rowred16(t["DGD16"], 4096, 4096, buf["XD16"], 2560, 2560, gd_ih, 1536, n, inv, D2=gd_hh, ldd2=1024, n_split=1536)
            grads[_D + "attention_rnn.weight_ih"], grads[_D + "attention_rnn.weight_hh"] = ga_ih, ga_hh
            grads[_D + "decoder_rnn.weight_ih"], grads[_D + "decoder_rnn.weight_hh"] = gd_ih, gd_hh
        else:
            gWa = _zeros(4096, 1792, device=dev)
            gWd = _zeros(4096, 2560, device=dev)
            ops.linear_dw(t["DGA"], 4096, XA, 1792, gWa, 1792, n, 4096, 1792, device=dev)
            ops.linear_dw(t["DGD"], 4096, XD, 2560, gWd, 2560, n, 4096, 2560, device=dev)
            grads[_D + "attention_rnn.weight_ih"] = gWa[:, :768].contiguous()
            grads[_D + "attention_rnn.weight_hh"] = gWa[:, 768:].contiguous()
            grads[_D + "decoder_rnn.weight_ih"] = gWd[:, :1536].contiguous()
            grads[_D + "decoder_rnn.weight_hh"] = gWd[:, 1536:].contiguous()
        if persistent and _BIAS_IN_LOOP:      # the persistent kernel summed the gate gradients while it produced them
            gba, gbd = t["gba"], t["gbd"]
        else:
            gba = _colsum(t["DGA"], n, 4096, 1, 0, 1, dev)
            gbd = _colsum(t["DGD"], n, 4096, 1, 0, 1, dev)
        grads[_D + "attention_rnn.bias_ih"], grads[_D + "attention_rnn.bias_hh"] = gba, _clone(gba)
        grads[_D + "decoder_rnn.bias_ih"], grads[_D + "decoder_rnn.bias_hh"] = gbd, _clone(gbd)
        gWq = _zeros(128, 1024, device=dev)
        ops.linear_dw(t["DQ"], 128, XD, 2560, gWq, 1024, n, 128, 1024, device=dev)
        grads[_A + "query_layer.linear_layer.weight"] = gWq
        for name, part, cols in ((_A + "v.linear_layer.weight", t["dv_part"], 128),
                                 (_A + "location_layer.location_dense.linear_layer.weight", t["dwloc_part"], 128 * 32),
                                 (_A + "location_layer.location_conv.conv.weight", t["dwconv_part"], 32 * 2 * 31)):
            g = _empty(cols, device=dev)
            L("t2v_sum_rows_per_batch", part, g, 1, B * nck, cols, 0.0)
            if name.endswith("location_conv.conv.weight"):          # partials are kept in the [2*31, 32] layout
                gt = _empty(cols, device=dev)
                L("t2v_transpose", g, 32, gt, 62, 62, 32, 0)
                g = gt
            grads[name] = g.view_as(P[name])
        _trace("  bwd decoder dW")
        # prenet backward (model.py:91-102)
        dP2pre = _empty(n, 256, device=dev)
        L("t2v_relu_drop_bwd", ctx["P2pre"], t["DXA"], 1792, dP2pre, n, 256, ctx["m1"], ctx["seed"], SITE_PRENET + 1, 0.5, 0, ops.R)
        g2 = _zeros(256, 256, device=dev)
        ops.linear_dw(dP2pre, 256, ctx["P1"], 256, g2, 256, n, 256, 256, device=dev)
        grads[_D + "prenet.layers.1.linear_layer.weight"] = g2
        dP1 = _empty(n, 256, device=dev)
        ops.linear_dx(dP2pre, 256, P[_D + "prenet.layers.1.linear_layer.weight"], 256, dP1, 256, n, 256, 256)
        dP1pre = _empty(n, 256, device=dev)
        L("t2v_relu_drop_bwd", ctx["P1pre"], dP1, 256, dP1pre, n, 256, ctx["m0"], ctx["seed"], SITE_PRENET, 0.5, 0, ops.R)
        g1 = _zeros(256, 80, device=dev)
        ops.linear_dw(dP1pre, 256, ctx["Fr"], 80, g1, 80, n, 256, 80, device=dev)
        grads[_D + "prenet.layers.0.linear_layer.weight"] = g1
        Wm = P[_A + "memory_layer.linear_layer.weight"]
        gWm = _zeros(128, 512, device=dev)
        ops.linear_dw(t["dpmem"], 128, ctx["memory"], 512, gWm, 512, B * Ti, 128, 512, device=dev)
        grads[_A + "memory_layer.linear_layer.weight"] = gWm
    # d(memory)[b,ti,:] = sum_t w_t[b,ti] * dctx_t[b,:]  (one batched GEMM over the saved alignments)  + dpmem @ W_m
    Wm = P[_A + "memory_layer.linear_layer.weight"]
    dmem = _empty(B * Ti, 512, device=dev)
    if ops.tc and _DMEM_TC and Ti % 4 == 0 and Ti <= 128:
        # on the tensor core (this GEMM heads the chain to the encoder backward): tf32-rounded copies of both operands, one 128 x 256
        # tile pair per utterance, the decoder steps are the reduction rows of the MN-major kernel
        al = _clone(buf["align"])
        L("t2v_round_tf32", al, al.numel())
        L("t2v_round_tf32", t["DCTX"], t["DCTX"].numel())
        L("t2v_gemm_tc_rowred_batched", al, Ti, Ti, To, t["DCTX"], B * 512, 512, 512, dmem, 512, Ti * 512, To, B, 1.0)
    else:
        L("t2v_gemm_f32", buf["align"], 1, Ti, t["DCTX"], 1, B * 512, dmem, 512, Ti, 512, To, 1.0, 0.0, None, B, To * Ti, 512, Ti * 512)
    ops.linear_dx(t["dpmem"], 128, Wm, 512, dmem, 512, B * Ti, 128, 512, accumulate=True)
    return dmem.view(B, Ti, 512), br


# ======================================================================================================= postnet + outputs
def postnet_forward(ops, P, X0p, B, To, training, masks, seed, dev, X0lo=None, X16=None):
    return conv_stack_forward(ops, P, "postnet.convolutions", X0p, B, To, [80, 512, 512, 512, 512, 80], [2, 2, 2, 2, 0],
                              training, masks, seed, SITE_POST, dev, round_last=False, Xlo=X0lo, X16=X16)


def split_lo(ops, x, x_hi):
    """tf32(x - x_hi): the low part of a split tensor-core operand"""
    return ops.lo(x, x_hi)


class TrainContext(object):
    pass


def forward_train(ops, P, text, in_len, mel_tgt, out_len, training=True, rand=None, seed=0, mask_padding=True,
                  mask_value=-float("inf")):
    """Tacotron2.forward + parse_output (model.py:509-547).  `rand`: None (RNG dropout from `seed`) or an object with
    .enc/.prenet/.dec/.post/.eps explicit keep masks (reference layouts, see oracle/port.py Rand)."""
    dev = text.device
    B, Ti = text.shape
    To = mel_tgt.shape[2]
    c = TrainContext()
    c.B, c.Ti, c.To, c.training, c.seed, c.rand = B, Ti, To, training, seed, rand
    g = (lambda name: None) if rand is None else (lambda name: getattr(rand, name))
    _trace("")
    # the decoder's weight packing, buffer initialisation and prenet only need the mel: a branch of their own
    br_prep = _Branch(2)
    with br_prep:
        prep = decoder_prepare(ops, P, mel_tgt, B, Ti, g("prenet"), seed, dev)
    # the reference encoder / VAE head only needs the mel: it runs beside the text encoder (both are chains of small kernels).
    # T2V_REFENC_LATE=1 forks the branch only after the encoder's convolution stack (its wide kernels then share the GPU with the
    # BiLSTM kernel instead of the convolution chain): measured 42.50 vs 42.24 ms per step, so the early fork stays the default.
    vae_out = {}

    def run_vae():
        br = _Branch(0, urgent=True)
        with br:
            eps = g("eps")
            if training and eps is None:
                eps = torch.empty(B, P["vae_gst.fc1.weight"].shape[0], device=dev, dtype=F32)
                L("t2v_randn", eps, eps.numel(), seed, SITE_EPS)
            vae_out["r"] = vae_forward(ops, P, mel_tgt, training, eps, dev)
        vae_out["br"] = br
    if not _REFENC_LATE:
        run_vae()
    HoutP, c.enc = encoder_forward(ops, P, text, in_len, training, g("enc"), seed, dev, packed=True,
                                   after_convs=run_vae if _REFENC_LATE else None)
    style, mulv, z, c.vae = vae_out["r"]
    vae_out["br"].join()
    _trace("fwd encoder + vae/ref-encoder")
    memory = _empty(B, Ti, 512, device=dev)
    L("t2v_unpad_add", HoutP, style, memory, B, Ti, 512, ops.R)                  # model.py:536-537
    br_prep.join()
    O, align, c.dec = decoder_forward(ops, P, memory, mel_tgt, in_len, training, g("prenet"), g("dec"), seed, mask_value, dev,
                                      prep=prep)
    _trace("fwd decoder")
    X0p = _zeros(B * (To + 4), 80, device=dev)
    L("t2v_rows_tb_to_padded", O, 84, X0p, B, 80, To, 0)
    X0r = X0p                                             # Postnet conv-0 operand (tf32-rounded copy in the tensor-core mode)
    X0lo = X016 = None
    if ops.tc:
        X0r = _zeros(B * (To + 4), 80, device=dev)
        L("t2v_rows_tb_to_padded", O, 84, X0r, B, 80, To, 1)
        if ops.split and ops.op16 == 1 and _SPLIT16:
            X016 = (torch.empty(B * (To + 4), 80, device=dev, dtype=torch.int16), torch.empty(B * (To + 4), 80, device=dev, dtype=torch.int16))
            L("t2v_split16", X0p, X016[0], X016[1], X0p.numel(), 1.0)
        elif ops.split:
            X0lo = split_lo(ops, X0p, X0r)
    Y5, c.post = postnet_forward(ops, P, X0r, B, To, training, g("post"), seed, dev, X0lo=X0lo, X16=X016)
    _trace("fwd postnet")
    lens = out_len if mask_padding else None
    mel = _empty(B, 80, To, device=dev)
    mel_post = _empty(B, 80, To, device=dev)
    gate = _empty(B, To, device=dev)
    L("t2v_padded_to_bct", X0p, None, mel, B, 80, To, lens, 0.0)
    L("t2v_padded_to_bct", X0p, Y5, mel_post, B, 80, To, lens, 0.0)
    L("t2v_gate_from_rows", O, 84, 80, gate, B, To, lens, 1000.0)
    if mask_padding:
        L("t2v_mask_padded_rows", X0r, B, 80, To, out_len)      # Postnet conv-0's saved input becomes the masked mel (Q10)
    Z = c.vae["Z"]
    mu = mulv[:, :Z].contiguous()
    logvar = mulv[:, Z:].contiguous()
    return [mel, mel_post, gate, align, mu, logvar, z], c


def backward_train(ops, P, c, dmel, dpost, dgate, dmu, dlogvar, grad_dst=None):
    """Gradients of every live parameter given the output gradients (reference layouts).  grad_dst: optional name -> destination
    tensors (slices of the flat gradient buffer) that the producers of the large gradients fill directly."""
    dev = dmel.device
    B, Ti, To = c.B, c.Ti, c.To
    grads = Grads(grad_dst)
    R = B * (To + 4)
    _trace("")
    dY5 = _zeros(R, 80, device=dev)
    L("t2v_bct_to_padded", dpost, dY5, B, 80, To, 0.0)
    # the Postnet weight gradients are off the path to the decoder: a side branch that runs beside the dX chain and then on
    # the ~20 SMs the persistent decoder-backward kernel leaves free
    br_post = _Branch(2) if _POST_DW_BRANCH else None
    dX0 = conv_stack_backward(ops, P, "postnet.convolutions", dY5, c.post, B, To, c.training, c.seed, SITE_POST, dev, grads,
                              need_dx=True, dw_branch=br_post)
    _trace("bwd postnet")
    dres = _zeros(R, 80, device=dev)                      # dmel + dpost (residual, model.py:543)
    L("t2v_bct_to_padded", dmel, dres, B, 80, To, 0.0)
    L("t2v_bct_to_padded", dpost, dres, B, 80, To, 1.0)
    dO = _empty(To * B, 84, device=dev)
    L("t2v_padded_to_rows_tb", dX0, dres, dgate, dO, 84, B, 80, To, ops.R)
    dmem, br_dw = decoder_backward(ops, P, dO, c.dec, dev, grads)
    _trace("bwd decoder")
    dstyle = _empty(B, 512, device=dev)
    L("t2v_sum_rows_per_batch", dmem, dstyle, B, Ti, 512, 0.0)
    # three independent tails: decoder weight gradients (branch started in decoder_backward), VAE / reference encoder,
    # text encoder
    br_vae = _Branch(1, urgent=True)
    with br_vae:
        vae_backward(ops, P, dstyle, dmu, dlogvar, c.vae, dev, grads)
    encoder_backward(ops, P, dmem, c.enc, c.training, c.seed, dev, grads)
    br_vae.join()
    br_dw.join()
    if br_post is not None:
        br_post.join()
    _trace("bwd encoder + vae/ref-encoder + decoder dW")
    return grads
