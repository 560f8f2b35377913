"""Host-side operator builders of the product (joxsz_b200/operators.py, packer.py) against the oracle,
including a numpy model of the CUDA map pipeline driven by the same packed tables."""
import numpy as np
import pytest

from helpers import orc, rel_err, rel_err_max
from kernel_model import model_stages, model_tail


@pytest.fixture(scope="module")
def packed(cl1226_fit):
    from joxsz_b200.packer import PackedSetup
    return PackedSetup(cl1226_fit, max_walkers=64)


def _pp_and_params(golden, s, idx):
    thetas = golden["thetas"][idx]
    ps = [s.full_params(t) for t in thetas]
    return np.stack([orc.press_fun(p, s.r_pp) for p in ps]), ps


def test_abel_matrix_equals_pyabel(cl1226_oracle):
    from joxsz_b200 import operators as ops
    s = cl1226_oracle
    rng = np.random.default_rng(5)
    for r in (s.r_pp, 16.0 * np.arange(1, 120)):          # non-uniform (x=r) and uniform (dx) branches
        A = ops.abel_forward_matrix(r)
        f = rng.random((3, r.size)) * np.exp(-r / 700.0)
        ref = orc.pyabel_direct_forward(f, r)
        assert np.max(np.abs(f @ A.T - ref)) < 1e-12 * np.max(np.abs(ref))
        assert np.all(A[-1] == 0.0) and np.all(np.tril(A, k=-1) == 0.0)
    assert ops.is_uniform_sampling(16.0 * np.arange(1, 120)) and not ops.is_uniform_sampling(s.r_pp)


def test_packed_geometry(packed):
    mo = packed.map_ops
    assert (mo.N, mo.H, mo.B, mo.P, mo.Q) == (171, 86, 55, 256, 129)
    assert packed.nr == 313 and packed.sep == 85 and packed.ndim == 13
    assert packed.proj_op.shape == (4 * mo.nseg, 313) and mo.seg.max() == mo.nseg - 1
    assert packed.slot_src[3] == -1 and packed.slot_val[3] == 0.014       # c frozen
    assert packed.prior_const == 0.0


def test_model_pipeline_matches_oracle_stages(golden, cl1226_oracle, packed):
    s = cl1226_oracle
    idx = [0, 1, 4, 9]
    pp, ps = _pp_and_params(golden, s, idx)
    m = model_stages(packed, pp)
    c = s.d_mat.shape[0] // 2
    for k, p in enumerate(ps):
        with np.errstate(all="ignore"):
            st = orc.sz_stages(p, s)
        assert rel_err_max(m["Z"][k], st["y_2d"][c:, c:]) < 1e-12
        assert rel_err_max(m["conv"][k], st["conv_2d"][c:, c:]) < 1e-12
        assert rel_err_max(m["row"][k], st["map_out"][c, c:]) < 1e-11
        tail = model_tail(packed, m["row"][k:k + 1], st["t_prof"][None, :], np.array([p["calibration"]]))
        assert rel_err_max(tail["bright"][0], st["bright"]) < 1e-11
        assert abs(tail["chisq"][0] - st["chisq"]) < 1e-8 * max(1.0, st["chisq"])


def test_direct_y_convolution_equals_fft_path(golden, cl1226_oracle, packed):
    """The mixed-domain beam table `bmix` (direct convolution along y, K3 phase B) gives the same convolved map as the
    cyclic FFT path with the beam spectrum `bhat`."""
    pp, _ = _pp_and_params(golden, cl1226_oracle, [0, 3])
    a = model_stages(packed, pp)
    b = model_stages(packed, pp, direct_b=True)
    assert packed.bmix.shape == (28, 129)
    assert rel_err_max(b["conv"], a["conv"]) < 1e-13
    assert rel_err_max(b["row"], a["row"]) < 1e-11


def test_filter_row_operator_is_the_circular_filter(cl1226_oracle, packed):
    """K7's constant operator applied to the distinct pixels of a D4- and transpose-symmetric map equals the central
    half row of real(ifft2(fft2(map) * filtering)) (reference joxsz_funcs.py:466-467, :472)."""
    from kernel_model import filter_row_operator
    s = cl1226_oracle
    H = packed.map_ops.H
    c = H - 1
    R = filter_row_operator(packed)
    assert R.shape == (H * (H + 1) // 2, H)
    rng = np.random.default_rng(11)
    q = rng.standard_normal((H, H))
    q = q + q.T                                                   # conv_c[u, v] = conv_c[v, u]
    full = np.empty((2 * c + 1, 2 * c + 1))
    a = np.abs(np.arange(2 * c + 1) - c)
    full[:] = q[a[:, None], a[None, :]]
    ref = np.real(np.fft.ifft2(np.fft.fft2(full) * s.filtering))[c, c:]
    iu, iv = np.triu_indices(H)
    got = q[iu, iv] @ R
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))


def test_y_operator(golden, cl1226_oracle, packed):
    s = cl1226_oracle
    pp, ps = _pp_and_params(golden, s, [0, 2])
    with np.errstate(all="ignore"):
        y_ref = np.stack([orc.sz_stages(p, s)["y"] for p in ps])
    assert rel_err_max(pp @ packed.y_op.T, y_ref) < 1e-13


def test_unsupported_geometry_is_rejected(cl1226_fit):
    import copy
    from joxsz_b200 import operators as ops
    from joxsz_b200.packer import PackedSetup, PackError
    sz = cl1226_fit.data.sz
    bad = copy.copy(cl1226_fit)
    bad.data = copy.copy(cl1226_fit.data)
    bad.data.sz = copy.copy(sz)
    bad.data.sz.d_mat = sz.d_mat[:-1, :-1]                  # even-sized map: the reference cannot build one
    with pytest.raises(ops.GeometryError):
        PackedSetup(bad)
    bad.data.sz = copy.copy(sz)
    bad.data.sz.beam_2d = sz.beam_2d + np.arange(sz.beam_2d.shape[0])[:, None] * 1e-3   # not symmetric
    with pytest.raises(ops.GeometryError):
        PackedSetup(bad)
    bad.data.sz = copy.copy(sz)
    bad.data.sz.calc_integ = True
    bad.data.sz.integ_sig = 0.0
    with pytest.raises(PackError):
        PackedSetup(bad)


def test_packing_does_not_depend_on_blas_threads(cl1226_fit):
    """The operators must come out bit-identical whatever thread count BLAS runs with: a `python` process (all cores)
    and the ranks torchrun starts (OMP_NUM_THREADS=1) otherwise disagree in the last bit of proj_op, and with it in
    the last bit of log-likelihoods -- found on hardware as a state_checksum mismatch between 1 and 2 GPUs."""
    import threadpoolctl
    from joxsz_b200.packer import PackedSetup
    packs = []
    for n in (1, 2, 5):
        with threadpoolctl.threadpool_limits(n):
            packs.append(PackedSetup(cl1226_fit, max_walkers=16, device=0))
    for k in ("proj_op", "y_op", "w_integ", "g_op", "w_t0", "bhat", "bmix", "hf", "cmat", "dinv"):
        a = np.asarray(getattr(packs[0], k))
        for other in packs[1:]:
            assert np.array_equal(a.view(np.int64), np.asarray(getattr(other, k)).view(np.int64)), k
