"""In-tree build of the C-ABI CUDA library (``gcc_b200/libgcc_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels with the
repo snapshot to the GPU box.  ``build()`` is incremental (mtime based).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgcc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(verbose=False, force=False):
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "gcc_b200.h"))
    hdr_mtime = max((os.path.getmtime(h) for h in headers if os.path.exists(h)), default=0.0)
    objs, procs = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_mtime):
            cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(os.path.dirname(HERE), "include"), "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append((src, out.decode(errors="replace")))
        elif verbose and out:
            print(out.decode(errors="replace"), file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join("== %s ==\n%s" % f for f in failed))
    if procs or not os.path.exists(LIB):
        cmd = [nvcc, "-arch=sm_100a", "-shared", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return LIB


def build_check(verbose=False):
    """CUDA-core cross-check kernels (scripts/csrc/conv_direct.cu): development probes only, NOT part of the product
    library or its public header.  Same argument lists as gcc_conv_gemm_bf16 / gcc_wgrad_gemm_bf16."""
    root = os.path.dirname(HERE)
    src = [os.path.join(root, "scripts", "csrc", f) for f in ("conv_direct.cu", "check_stub.cu")]
    out = os.path.join(root, "scripts", "libgcc_b200_check.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(s) for s in src):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-I", CSRC, "-I", os.path.join(root, "include"), "-shared", "-o", out] + src
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
