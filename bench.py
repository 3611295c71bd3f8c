#!/usr/bin/env python
"""bench.py -- padded mel-frames/s of one Tacotron2-VAE train step (fwd + loss + bwd + clip + Adam) on N B200s.

Workload (BASELINE.json config 3 / SURVEY.md C3): batch 64 per GPU, text <= 120, mel 80 x <= 800, synthetic inputs and
random-init weights of the reference architecture, data-parallel over N ranks (weak scaling: the per-GPU batch is
fixed).  One JSON line on rank 0:
  value     : device-resident throughput (inputs already in HBM)
  e2e       : the same step through the public API with the batch copied from pinned host memory and the loss read back
  roofline  : the decoder step (the path's dominant kernel group) against the measured HBM peak
  cpu_baseline : the CPU oracle (oracle/port.py, a restatement of the reference's PyTorch path) on a bounded sample
`--impl reference` times that CPU oracle instead (the reference's own CPU path; /root/reference does not travel).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "tacotron2-vae_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ti", type=int, default=120)
    ap.add_argument("--to", type=int, default=800)
    ap.add_argument("--precision", default=os.environ.get("T2V_PRECISION", "fp16"), choices=["fp16", "bf16", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """samples nvidia-smi SM clocks / throttle reasons during the timed region"""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _oracle_state():
    from oracle import port
    P = port.init_params(1234)
    params = {k: v.clone().requires_grad_(True) for k, v in P.items() if v.dtype.is_floating_point and "running" not in k}
    state = dict(P)
    state.update(params)
    return port, state, params


def _oracle_step(port, state, params, opt, batch, rand):
    opt.zero_grad()
    out = port.tacotron2_forward(state, batch[0], batch[1], batch[2], batch[4], True, rand)
    loss, _, _ = port.vae_loss(out, batch[2], batch[3], 0.001)
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.grad is not None], 1.0)
    opt.step()


def pick_cpu_threads(cores):
    """PyTorch's CPU path stops scaling (and collapses) well below the core count of a GPU host for these small per-step
    ops; calibrate on a tiny sample and use the fastest setting.  The value chosen is what `cores` reports."""
    port, state, params = _oracle_state()
    opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-6)
    batch = port.synthetic_batch(4, 40, 24, seed=0)
    rand = port.Rand.draw(4, 40, 24, seed=1)
    best, best_t = None, None
    for n in [c for c in (8, 16, 32) if c <= max(cores, 8)]:
        torch.set_num_threads(n)
        _oracle_step(port, state, params, opt, batch, rand)
        t0 = time.perf_counter()
        _oracle_step(port, state, params, opt, batch, rand)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    return best


def cpu_oracle_throughput(B, Ti, To, steps, warmup, threads):
    """frames/s of the CPU oracle (fwd + loss + bwd + clip + Adam) on a [B,Ti,To] sample"""
    torch.set_num_threads(threads)
    port, state, params = _oracle_state()
    opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-6)
    batch = port.synthetic_batch(B, Ti, To, seed=0)
    rand = port.Rand.draw(B, Ti, To, seed=1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _oracle_step(port, state, params, opt, batch, rand)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return B * To * len(times) / sum(times), sum(times) / len(times)


def _progress(msg):
    if os.environ.get("T2V_BENCH_VERBOSE"):
        print("[bench %.1fs] %s" % (time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def main():
    a = parse_args()
    if os.environ.get("T2V_STALL_DUMP"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["T2V_STALL_DUMP"]), repeat=True, file=sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return
        # bounded sample of the C3 workload: the SAME batch and text length (BatchNorm statistics and every GEMM shape depend on
        # the batch), only the number of mel frames per utterance is cut (the decoder loop is linear in it) so that K + W steps
        # end within a few minutes
        B_s = a.batch
        To_s = max(16, min(a.to, int(120000.0 / max(1, a.batch) / max(1, a.steps + min(a.warmup, 2)))))   # ~17 k frames / step at the default K=5, W=2
        threads = pick_cpu_threads(cores)
        fps, sec = cpu_oracle_throughput(B_s, a.ti, To_s, a.steps, min(a.warmup, 2), threads)
        sample = "B=%d,Ti=%d,To=%d (%d padded frames/step) of the B=%d,To=%d workload: frames per utterance cut, batch kept" % (
            B_s, a.ti, To_s, B_s * To_s, a.batch, a.to)
        print(json.dumps({
            "impl": "reference", "metric": "padded mel-frames/s (train fwd+bwd+opt)", "value": fps, "unit": "frames/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": min(a.warmup, 2), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: Tacotron2-VAE train step, batch %d/GPU, Ti<=%d, To<=%d" % (a.batch, a.ti, a.to),
                       "sample": sample, "parallelism": "dp1", "global_batch": a.batch,
                       "note": "CPU oracle (oracle/port.py, restatement of the reference's PyTorch CPU path) on a bounded sample"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "host_cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("T2V_NCCL_DEBUG"):      # NCCL's own init log (communicators, rings / trees / NVLS) into a file per rank
            os.environ.setdefault("NCCL_DEBUG", "INFO")
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,GRAPH")
            os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(ROOT, "gpurun_out", "nccl_init_%h_%p.log"))
        dist.init_process_group("nccl", device_id=dev)
    import model as t2v_model
    from hparams import create_hparams
    from loss_function import Tacotron2Loss_VAE
    from oracle import port                      # synthetic_batch only (input generator), never on the timed path
    from t2v import _lib, optim

    hp = create_hparams("anneal_function=constant,batch_size=%d" % a.batch)
    torch.manual_seed(hp.seed)
    m = t2v_model.Tacotron2(hp).to(dev).train()
    m.precision = a.precision
    crit = Tacotron2Loss_VAE(hp)
    opt = optim.FusedAdamClip(m, lr=hp.learning_rate, weight_decay=hp.weight_decay, max_norm=hp.grad_clip_thresh)
    flat = opt.flat
    if world > 1:
        for t in m.state_dict().values():
            dist.broadcast(t, 0)
    B, Ti, To = a.batch, a.ti, a.to
    host = [t.pin_memory() for t in port.synthetic_batch(B, Ti, To, seed=rank)]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    loss_host = torch.zeros(1).pin_memory()

    def to_device():
        return tuple(t.to(dev, non_blocking=True) for t in host)

    def step(batch_dev, it, read_loss):
        x = (batch_dev[0], batch_dev[1], batch_dev[2], Ti, batch_dev[4], batch_dev[5].float(), batch_dev[6].float())
        y = (batch_dev[2], batch_dev[3])
        opt.zero_grad()
        out = m(x)
        loss, _, _, _ = crit(out, y, it)
        loss.backward()
        if world > 1:
            flat.adopt_grads()
            dist.all_reduce(flat.buffer)
        opt.step(grad_scale=1.0 / world)
        if read_loss:
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = to_device()                                 # the device-resident leg's batch: lives for the whole run

    def timed(n, e2e):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.reset_launch_count()
        ev0.record()
        for it in range(n):
            b = to_device() if e2e else resident
            step(b, it, e2e)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, _lib.launch_count()

    _progress("model built, starting warm-up")
    # warm-up (>= 3 steps) in the end-to-end form: a superset of the resident step (the first per-step batch allocation next to the
    # resident batch is a cudaMalloc, which stalls the host once and would otherwise land inside the e2e timed region)
    timed(max(a.warmup, 3), True)
    _progress("warm-up done")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(a.steps, False)
    _progress("device-resident timing done: %.1f ms/step" % (ms / a.steps))
    ms_e2e, _ = timed(a.steps, True)
    sampler.stop_flag = True
    _progress("e2e timing done")
    frames = B * To * world * a.steps
    value = frames / (ms * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)

    # roofline of the dominant kernel group: one teacher-forced decoder step (GEMM_a, cell, q-GEMM, fused attention,
    # GEMM_d, cell), timed with CUDA events over a fresh forward's time loop
    roof = decoder_step_roofline(m, hp, B, Ti, To, dev, a.precision)
    infer_step = inference_decoder_step(m, dev, a.precision) if rank == 0 else None
    bwd_step = decoder_step_backward(m, B, Ti, To, dev, a.precision) if rank == 0 else None

    stft = stft_mel_roofline(dev, B, To) if rank == 0 else None
    _progress("roofline done")
    comm = None
    if world > 1:      # the gradient exchange alone: one ncclAllReduce(SUM) over the flat fp32 gradient buffer, device-timed, max over ranks
        times = []
        for _ in range(5):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dist.all_reduce(flat.buffer); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        t = torch.tensor([min(times[1:])], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = flat.numel * 4
        comm = {"backend": "nccl", "nccl_version": ".".join(str(v) for v in torch.cuda.nccl.version()), "world_size": dist.get_world_size(),
                "allreduce_bytes": nbytes, "allreduce_ms": float(t), "allreduce_busbw_gbs": nbytes * 2 * (world - 1) / world / (float(t) * 1e-3) / 1e9,
                "share_of_step": float(t) / (ms / a.steps),
                "note": "one blocking all-reduce after the backward graph; not overlapped (it is %.1f %% of the step)" % (100 * float(t) / (ms / a.steps))}
    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        Bs, Tos = B, 48                                   # same batch, frames per utterance cut (see --impl reference)
        threads = pick_cpu_threads(cores)
        fps, sec = cpu_oracle_throughput(Bs, Ti, Tos, 2, 1, threads)
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "host_cores": cores, "kind": "port",
               "sample": "B=%d,Ti=%d,To=%d (%d padded frames/step, 2 steps + 1 warm-up) of the C3 workload" % (Bs, Ti, Tos, Bs * Tos)}
    _progress("cpu baseline done")
    if rank == 0:
        print(json.dumps({
            "metric": "padded mel-frames/s (train fwd+bwd+opt)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32"}.get(a.precision, a.precision), "data": "synthetic",
            "config": {"workload": "C3: Tacotron2-VAE train step, batch %d/GPU, Ti<=%d, To<=%d, fp32 master weights / state, %s" %
                       (B, Ti, To, {"fp16": "fp16 operands in the decoder loops (tcgen05 kind::f16), tf32 GEMMs elsewhere",
                                    "bf16": "bf16 operands in the decoder loops (tcgen05 kind::f16), tf32 GEMMs elsewhere",
                                    "tf32": "tcgen05 tf32 GEMMs", "fp32": "FFMA fp32 GEMMs"}[a.precision]),
                       "parallelism": "dp%d" % world, "global_batch": B * world,
                       "l2": "per-step working set (>3 GB of saved activations) exceeds the 126 MB L2"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": launches, "comm": comm, "roofline": roof, "decoder_step_inference": infer_step, "decoder_step_backward": bwd_step, "stft_mel": stft, "cpu_baseline": cpu,
            "clocks": sampler.summary()}))
    if world > 1:
        dist.destroy_process_group()


def stft_mel_roofline(dev, B, To):
    """The mel front-end of one batch (reference TacotronSTFT.mel_spectrogram, layers.py:75-92): B waveforms of To * 256 samples ->
    [B, 80, To + 1] log-mel, ONE fused kernel (stft_fused.cu).  Algorithmic bytes: 1024 B in + 320 B out per frame."""
    from layers import TacotronSTFT
    st = TacotronSTFT(1024, 256, 1024, 80, 16000, 0.0, 8000.0).to(dev)
    wav = torch.rand(B, To * 256, device=dev) * 2 - 1
    for _ in range(3):
        mel = st.mel_spectrogram(wav)
    times = []
    flush = torch.empty(64 << 20, device=dev)            # 256 MB > the 126 MB L2: the waveforms (52 MB) are re-read from HBM every time
    for _ in range(5):
        flush.fill_(0.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mel = st.mel_spectrogram(wav); e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    frames = B * (To + 1)
    ms = min(times)
    peak = 6514.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    gbs = frames * 1344 / (ms * 1e-3) / 1e9
    return {"kernel": "stft_mel_fused_kernel", "frames": frames, "ms": ms, "frames_per_s": frames / (ms * 1e-3), "bound": "hbm",
            "algorithmic_bytes_per_frame": 1344, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
            "finite": bool(torch.isfinite(mel).all()),
            "l2": "flushed before every timed launch",
            "note": "one warp per 1024-point FFT (two real frames, 32 x 32 in registers): bound by instruction issue / latency of the "
                    "register FFT and the band-limited mel loop (16 warps per SM at 98 registers), not by HBM"}


def decoder_step_backward(m, B, Ti, To, dev, precision):
    """Times the reverse-time decoder loop alone (t2v_decoder_bwd_steps over To steps: dec_persist_bwd_kernel when it applies)
    with CUDA events around the launch inside engine.decoder_backward.  Algorithmic bytes per step: the same fp32 weights as the
    forward step (72.4 MB, read once) + saved activations of the step + gate-gradient rows written."""
    from t2v import engine
    import t2v.engine as E
    with torch.no_grad():
        ops = engine.Ops(precision)
        P = m._state()
        mem = torch.randn(B, Ti, 512, device=dev)
        mel = torch.randn(B, 80, To, device=dev)
        in_len = torch.full((B,), Ti, device=dev, dtype=torch.long)
        dO = torch.randn(To * B, 84, device=dev) * 0.1
        _, _, ctx = engine.decoder_forward(ops, P, mem, mel, in_len, True, None, None, 1, -float("inf"), dev)
        times = []
        orig = E.L

        def timed_L(name, *args):
            if name != "t2v_decoder_bwd_steps":
                return orig(name, *args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            r = orig(name, *args)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3 / To)
            return r
        E.L = timed_L
        try:
            for _ in range(3):
                _, br = engine.decoder_backward(ops, P, dO, ctx, dev, {})
                br.join()
                torch.cuda.synchronize()
        finally:
            E.L = orig
    us = min(times[1:])
    s4 = 4
    alg = 18103953 * s4 + (B * Ti * 640 + B * Ti * 128 + B * (4096 * 4 + 1024 * 4 + 1792 + 2560 + 1536)) * s4
    return {"value": us, "unit": "us/step", "batch": B, "text_len": Ti, "steps": To, "algorithmic_bytes_per_step": alg,
            "achieved_gbs": alg / (us * 1e-6) / 1e9,
            "what": "reverse-time Decoder.decode step (both LSTM cell backwards, attention backward incl. location-layer weight "
                    "gradients, dX GEMMs): dec_persist_bwd_kernel time / To"}


def inference_decoder_step(m, dev, precision, B=16, Ti=120, n=1000):
    """Config 5: batch-16 free-running autoregressive decode for `n` steps (prenet -> decode -> mel/gate projection ->
    feedback, per-row stop bookkeeping on the device), replayed as one CUDA graph.  Returns mean us per decoder step."""
    from t2v import engine, infer
    with torch.no_grad():
        ops = engine.Ops(precision)
        P = m._state()
        mem = torch.randn(B, Ti, 512, device=dev)

        def run():
            sess = infer.DecoderSession(ops, P, mem, None, n, training=False, seed=7)
            sess.run_free(n, 0.5, seed=7)
            return sess
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            keep = run()
        times = []
        for _ in range(4):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        mel, gate, align = keep.outputs(n)
        ok = bool(torch.isfinite(mel).all())
    return {"value": min(times[1:]) * 1e3 / n, "unit": "us/step", "batch": B, "text_len": Ti, "steps": n, "finite": ok,
            "what": "free-running Decoder.inference step incl. prenet, attention, both LSTM cells, mel/gate projection, stop flags: "
                         "dec_persist_fwd_kernel<INFER> time / steps (one persistent launch; per-step launches: ~88 us)"}


def decoder_step_roofline(m, hp, B, Ti, To, dev, precision):
    """Times the teacher-forced decoder time loop alone (t2v_decoder_fwd_steps over To steps) with CUDA events."""
    from t2v import _lib, engine
    from t2v._lib import call as L
    with torch.no_grad():
        ops = engine.Ops(precision)
        P = m._state()
        mem = torch.randn(B, Ti, 512, device=dev)
        mel = torch.randn(B, 80, To, device=dev)
        in_len = torch.full((B,), Ti, device=dev, dtype=torch.long)
        _, _, ctx = engine.decoder_forward(ops, P, mem, mel, in_len, True, None, None, 1, -float("inf"), dev)   # warm
        S = ctx["S"]
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()                      # same replay mechanism as the train step
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            L("t2v_decoder_fwd_steps", S, 0, To)
        times = []
        for _ in range(4):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times = times[1:]
    us = min(times) * 1e3 / To
    s = 2 if precision in ("fp16", "bf16") else 4
    # weights 18 103 953 x s + activations x s, s = bytes per operand element (2: fp16 / bf16 -- the north star's
    # own denominator, 47.32 MB at B=64 -- or 4: fp32 storage)
    alg_bytes = 18103953 * s + (B * Ti * 640 + 4 * B * Ti + B * 8192 + B * 1361) * s
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (us * 1e-6) / 1e9
    persist = os.environ.get("T2V_PERSIST", "1") != "0" and precision != "fp32" and B <= 64 and Ti <= 128
    kernel = ("dec_persist_fwd_kernel: ONE persistent launch for all To steps of Decoder.decode (128 CTAs = 32 clusters x 4; "
              "TMA weight streaming, tcgen05 %s split-K over the cluster + DSMEM exchange, LSTM cells in the epilogue, query "
              "projection as a per-CTA UMMA, attention by CTA pairs); per-step time = kernel time / To" %
              ("kind::f16 (%s operands)" % precision if s == 2 else "kind::tf32")) if persist else (
              "decoder step = gemm_tc<128,4,8,64> x3 (attention_rnn gates, query, decoder_rnn gates) + lstm_pointwise_fwd x2 + "
              "attn3_row (fused energy/softmax/context): 6 launches per step in two concurrent chains, CUDA-graph replay")
    traffic, traffic_src = None, None
    try:      # DRAM bytes per step of the same kernel from the committed ncu --set full capture (NOT measured in this run)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_persist_dram.json")))
        if persist and tj.get("precision") == precision:
            traffic = tj["dram_bytes_per_step"]
            traffic_src = "profiles/r02_ncu_persist_dram.json: ncu --set full of dec_persist_fwd_kernel, B=%d Ti=%d To=%d, %s" % (
                tj.get("B", 0), tj.get("Ti", 0), tj.get("To", 0), tj.get("precision"))
    except Exception:  # noqa: BLE001
        pass
    return {"kernel": kernel,
            "bound": "hbm",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src,
            "frac_on_dram_traffic": (traffic / (us * 1e-6) / 1e9 / peak) if traffic else None,
            "us_per_step": us, "algorithmic_bytes_per_step": alg_bytes,
            "algorithmic_bytes_note": "SURVEY 8(d): 18 103 953 weights + step activations, x %d bytes per operand element (%s)" % (
                s, "the north star's bf16 denominator" if s == 2 else "fp32 storage"),
            "target_us_for_half_of_peak": alg_bytes / (0.5 * peak * 1e9) * 1e6,
            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
            "note": "frac = algorithmic bytes per step (every weight and activation byte a step must touch, the north star's "
                    "denominator) / measured time / HBM peak.  In the 16-bit modes the 36 MB of weights stay L2-resident, so DRAM itself "
                    "only sees `traffic` (the saved activations being written): the kernel is bound by the recurrence's dependency "
                    "chain over an L2 -> SM operand stream, not by HBM (DESIGN.md section 3)"}


if __name__ == "__main__":
    main()
