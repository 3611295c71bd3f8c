"""Builds libt2v_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, next to the package)."""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG, "csrc")
# T2V_LIB_SUFFIX selects a side build (e.g. "_san": long wait limits for compute-sanitizer runs) next to the product library
SUFFIX = os.environ.get("T2V_LIB_SUFFIX", "")
LIB = os.path.join(PKG, "libt2v_b200%s.so" % SUFFIX)
SOURCES = ["api.cu", "gemm_simt.cu", "gemm_tc.cu", "pointwise.cu", "attention.cu", "attention2.cu", "misc.cu", "stft_fused.cu", "rnn.cu", "rnn_persist.cu", "decoder.cu", "decoder_persist.cu", "decoder_persist_bwd.cu"]


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(PKG), "include", "t2v_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, "build" + SUFFIX), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(PKG, "build" + SUFFIX, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC,-fvisibility=hidden", "-c", os.path.join(CSRC, src), "-o", obj]
        cmd[1:1] = os.environ.get("T2V_NVCC_FLAGS", "").split()     # tuning builds (-DT2V_...=...)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
