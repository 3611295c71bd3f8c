"""Import-compatibility module (reference fp16_optimizer.py:21-41 are imported by model.py:8 / train.py:13).

The legacy fp16 + dynamic-loss-scale path is superseded on B200: master weights and state stay fp32 and the tensor
cores consume tf32/bf16 operands, so no loss scaling exists here.  The converters keep their semantics; constructing
FP16_Optimizer says why it is not needed."""
import torch


def _convert(val, pred, fn):
    if isinstance(val, (tuple, list)):
        return type(val)(_convert(v, pred, fn) for v in val)
    if isinstance(val, torch.Tensor) and pred(val):
        return fn(val)
    return val


def fp32_to_fp16(val):
    return _convert(val, lambda t: t.dtype == torch.float32, lambda t: t.half())


def fp16_to_fp32(val):
    return _convert(val, lambda t: t.dtype == torch.float16, lambda t: t.float())


class FP16_Optimizer(object):
    def __init__(self, *a, **k):
        raise RuntimeError("fp16_run is not needed on the B200 engine: run with fp16_run=False; reduced-precision "
                           "tensor-core operands (tf32/bf16) are handled inside the kernels with fp32 master weights")
