"""ctypes binding of libt2v_b200.so.  The argtypes are generated from include/t2v_b200.h itself, so the header is
the single source of truth for the ABI (tests check that every declared symbol is exported).

There is NO fallback: if the library is missing this module raises, and every op of the model fails loudly."""
import ctypes
import os
import re

import torch

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG, "libt2v_b200%s.so" % os.environ.get("T2V_LIB_SUFFIX", ""))
HEADER = os.path.join(os.path.dirname(PKG), "include", "t2v_b200.h")

_P = ctypes.c_void_p
_CT = {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "long long": ctypes.c_longlong,
       "unsigned long long": ctypes.c_ulonglong, "unsigned int": ctypes.c_uint, "cudaStream_t": _P,
       "unsigned long long n": ctypes.c_ulonglong}


class T2VDecoderSeq(ctypes.Structure):
    _fields_ = ([("B", ctypes.c_int), ("Ti", ctypes.c_int), ("To", ctypes.c_int), ("use_tc", ctypes.c_int),
                 ("training", ctypes.c_int), ("p_att", ctypes.c_float), ("p_dec", ctypes.c_float),
                 ("seed", ctypes.c_ulonglong), ("drop_masks", _P), ("mask_value", ctypes.c_float), ("in_lens", _P)] +
                [(n, _P) for n in ("Wa", "ba1", "ba2", "Wd", "bd1", "bd2", "Wq", "Wconv", "Wloc", "v", "mem", "pmem",
                                   "XA", "XD", "CA", "CD", "CUM", "align", "GA", "GD", "CPA", "CPD", "ASAVE", "parts",
                                   "qparts", "ebuf", "WaP", "WdP", "HCHI", "HCLO")] +
                [("op16", ctypes.c_int)] + [(n, _P) for n in ("XA16", "XD16", "WaP16", "WdP16", "mem16")])


class T2VDecoderBwd(ctypes.Structure):
    _fields_ = [("f", T2VDecoderSeq)] + [(n, _P) for n in (
        "WaT", "WdT", "WqT", "DHC", "DGA", "DGD", "DXA", "DXD", "dCa", "dCd", "dwprev", "gcum", "dpmem", "DCTX", "dw_part",
        "DQ", "dHq", "dv_part", "dwloc_part", "dwconv_part", "WaTP", "WdTP")] + [("op16", ctypes.c_int)] + [(n, _P) for n in (
        "DGA16", "DGD16", "WaTP16", "WdTP16", "dg_scale", "gb_att", "gb_dec")]


class T2VDecoderInfer(ctypes.Structure):
    _fields_ = ([("f", T2VDecoderSeq)] + [(n, _P) for n in ("Wp1", "Wp2", "Wpg", "bpg", "prenet_masks", "O", "P1")] +
                [("gate_threshold", ctypes.c_float), ("n_frames", _P)])


def parse_header(path=HEADER):
    """-> {name: (restype, [argtype, ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|unsigned long long|int|void)\s+(t2v_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(_P)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CT[ty])
        restype = {"const char*": ctypes.c_char_p, "unsigned long long": ctypes.c_ulonglong, "int": ctypes.c_int,
                   "void": None}[ret]
        protos[name] = (restype, argtypes)
    return protos


def load_library(path=LIB_PATH):
    if not os.path.isfile(path):
        raise RuntimeError("libt2v_b200.so not found at %s -- build it with __graft_entry__.build() "
                           "(python tacotron2-vae_b200/t2v/build.py); there is no CPU/PyTorch fallback" % path)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)          # AttributeError here == header/ABI mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


_lib = None
_DEBUG_SYNC = bool(int(os.environ.get("T2V_DEBUG_SYNC", "0")))


def lib():
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


def ptr(t):
    """device pointer of a tensor (None -> NULL); refuses anything that is not a CUDA tensor"""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    if not t.is_cuda:
        raise RuntimeError("t2v kernels take CUDA tensors only (got a %s tensor)" % t.device)
    return t.data_ptr()


def call(name, *args):
    """Call a stream-taking entry point: tensors -> device pointers, current torch stream appended."""
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor):
            conv.append(ptr(a))
        elif isinstance(a, ctypes.Structure):
            conv.append(ctypes.addressof(a))
        else:
            conv.append(a)
    conv.append(torch.cuda.current_stream().cuda_stream)
    r = getattr(lib(), name)(*conv)
    if r != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, r, lib().t2v_last_error().decode()))
    if _DEBUG_SYNC:
        try:
            torch.cuda.synchronize()
        except Exception as e:       # noqa: BLE001
            raise RuntimeError("%s%r -> device fault: %s" % (name, tuple(a for a in args if not isinstance(a, torch.Tensor)), e))


def launch_count():
    return int(lib().t2v_launch_count())


def reset_launch_count():
    lib().t2v_reset_launch_count()


def add_launch_count(n):
    lib().t2v_add_launch_count(int(n))
