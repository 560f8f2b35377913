"""The C-ABI library builds, loads, and exports every symbol include/joxsz_b200.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "joxsz_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jx_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from joxsz_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _header_functions()
    assert "jx_loglike" in names and "jx_create" in names and len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.PROTOTYPES), "python prototypes and header disagree"
    _lib.load()
    info = _lib.load().jx_build_info()
    assert b"sm_100a" in info


def test_setup_struct_layout_matches_header():
    """sizeof/offsets of the ctypes mirror equal what the C compiler lays out."""
    from joxsz_b200 import _lib
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "joxsz_b200.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %d\n", sizeof(jx_setup), offsetof(jx_setup, slot_val),
               offsetof(jx_setup, prior_const), offsetof(jx_setup, r_pp), offsetof(jx_setup, tmin),
               offsetof(jx_setup, bkgterm), (int)JX_NPAR);
        return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        vals = [int(v) for v in subprocess.check_output([exe]).split()]
    S = _lib.JxSetup
    assert vals == [ctypes.sizeof(S), S.slot_val.offset, S.prior_const.offset, S.r_pp.offset, S.tmin.offset,
                    S.bkgterm.offset, _lib.JX_NPAR]


def test_no_gpu_means_loud_failure(cl1226_fit):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from joxsz_b200 import _lib
    from joxsz_b200.batched import BatchedLikelihood
    with pytest.raises(_lib.JxError):
        BatchedLikelihood(cl1226_fit)
    with pytest.raises(_lib.JxError):
        cl1226_fit.press.press_fun(cl1226_fit.pars, cl1226_fit.data.sz.r_pp)
    # jx_create itself refuses without a device
    from joxsz_b200.packer import PackedSetup
    pk = PackedSetup(cl1226_fit, max_walkers=8)
    h = ctypes.c_void_p()
    rc = _lib.load().jx_create(ctypes.byref(pk.struct()), ctypes.byref(h))
    assert rc == -3 and b"no CUDA device" in _lib.load().jx_last_error(None)


def test_fft_codelets_on_host():
    """Radix-16 codelet, twiddles and exchange layout of the map kernel, run on the CPU."""
    exe = os.path.join(ROOT, "tests", "host", "fft_host_test.bin")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "fft_host_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
