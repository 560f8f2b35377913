"""Developer tool: SM cycles per phase of the map kernel (K3).

Builds a private copy of the library with -DJX_K3_CLOCKS (build/libjoxsz_b200_clk.so), runs the shipped
geometry with W walkers and prints the cycles per walker of phases A0 / A1 / B / C as seen by thread 0
of each CTA.  Usage (on a GPU box): python scripts/k3_phase_clocks.py [W]
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from joxsz_b200 import build as jb, _lib  # noqa: E402


EXTRA = os.environ.get("JX_CLK_DEFS", "").split()      # e.g. "-DJX_K3W_NO_M": experiment variants of the map kernel
TAG = os.environ.get("JX_CLK_TAG", "")


def build_variant():
    out = os.path.join(jb.HERE, "build", f"libjoxsz_b200_clk{TAG}.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    flags = [f for f in jb.NVCC_FLAGS if not f.startswith("--use_fast_math")]
    srcs = [os.path.join(jb.CSRC, s) for s in jb.SOURCES]
    subprocess.check_call(["nvcc", *flags, "-DJX_K3_CLOCKS", *EXTRA, "-shared", "-o", out, *srcs, "-lcudart"])
    return out


if __name__ == "__main__":
    if "--build-only" in sys.argv:
        print(build_variant())
        sys.exit(0)
    path = os.path.join(jb.HERE, "build", f"libjoxsz_b200_clk{TAG}.so")
    if not os.path.exists(path):
        build_variant()
    _lib.LIB_PATH = path
    import numpy as np
    import torch
    from joxsz_b200 import cluster
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.mb import mb
    from joxsz_b200.synthetic import draw_parameters

    W = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32768
    mb.fit.debugfit = False
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    wl = os.environ.get("JX_CLK_WORKLOAD", "cl1226")         # synth255 / synth511: the large-map kernel (K3L2)
    if wl == "synth255":
        inp = cluster.synthetic_inputs(map_half=127, nr=512, base=inp)
    elif wl == "synth511":
        inp = cluster.synthetic_inputs(map_half=255, nr=1024, base=inp)
    fit, _ = cluster.build_fit(inp, savedir=None)
    eng = BatchedLikelihood(fit, max_walkers=W, device=0)
    theta = torch.from_numpy(draw_parameters(fit.thawed, n=W, seed=4, spread=0.03, frac_bad=0.0)).cuda()
    lib = _lib.load()
    lib.jx_debug_k3_clocks.argtypes = [C.POINTER(C.c_ulonglong)]
    out = (C.c_ulonglong * 8)()
    for _ in range(3):
        ll = eng(theta)
    torch.cuda.synchronize()
    lib.jx_debug_k3_clocks(out)
    n = 5
    for _ in range(n):
        ll = eng(theta)
    torch.cuda.synchronize()
    lib.jx_debug_k3_clocks(out)
    nfin = int(torch.isfinite(ll).sum()) if hasattr(ll, "sum") else int(np.isfinite(ll).sum())
    names = ["A0 synth", "A1 rows", "B cols", "C rows+store"]
    tot = sum(out[i] for i in range(4))
    if tot:
        print(f"k3_szmap_kernel: walkers {W} (finite {nfin}), {n} launches; cycles per evaluated walker (thread 0 of its CTA):")
        for i, nm in enumerate(names):
            print(f"  {nm:10s} {out[i] / (n * W):10.0f}  {100.0 * out[i] / tot:5.1f} %")
        print(f"  total      {tot / (n * W):10.0f}")
    lib.jx_debug_k3l2_clocks.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.jx_debug_k3l2_clocks(out)
    if sum(out):
        print(f"k3l2_szmap_kernel ({wl}): walkers {W}, {n + 3} launches; cycles per walker (thread 0 of its CTA)")
        for i, nm in enumerate(["A0 synth", "A1 rows", "B cols", "C rows+store"]):
            print(f"  {nm:16s} {out[i] / ((n + 3) * W):10.0f}")
        print(f"  total            {sum(out[i] for i in range(4)) / ((n + 3) * W):10.0f}")
    # warp-specialised kernel (k3w_szmap.cu): thread 0 of the F group and thread 0 of the M group
    lib.jx_debug_k3w_clocks.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.jx_debug_k3w_clocks(out)
    if sum(out):
        print(f"k3w_szmap_kernel: walkers {W}, {n + 3} launches; cycles per walker")
        nm = ["F A0 synth", "F A1 rows", "M B compute", "M B write-back", "F wait bdone", "F C rows+store", "M wait full", "-"]
        for i in range(7):
            print(f"  {nm[i]:16s} {out[i] / ((n + 3) * W):10.0f}")
        print(f"  F total {sum(out[i] for i in (0, 1, 4, 5)) / ((n + 3) * W):10.0f}   M total {sum(out[i] for i in (2, 3, 6)) / ((n + 3) * W):10.0f}")
