"""Ensemble sampler: sharding / index logic on CPU (numpy model of the K6 kernels, gloo world_size 2),
the CUDA K6 kernels against that model, and the device sampler's emcee semantics on the GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

import kernel_model as km
from joxsz_b200.sampler import EnsembleSampler, shard_bounds
from kernel_model import split_permutation

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


# ------------------------------------------------------------------------------------------- host logic

def test_shard_bounds_cover_exactly():
    for ns in (0, 1, 7, 15, 16, 32768, 32769):
        for world in (1, 2, 3, 4, 8):
            seen = []
            per0 = shard_bounds(ns, world, 0)[0]
            for r in range(world):
                per, first, count = shard_bounds(ns, world, r)
                assert per == per0 and count <= per
                assert first == min(r * per, ns)
                seen += list(range(first, first + count))
            assert seen == list(range(ns))


def test_split_permutation_is_shared_and_balanced():
    p = split_permutation(101, 7, 3)
    assert np.array_equal(p, split_permutation(101, 7, 3))
    assert sorted(p.tolist()) == list(range(101))
    assert not np.array_equal(p, split_permutation(101, 7, 4))
    assert not np.array_equal(p, split_permutation(101, 8, 3))
    # colour of position q is q & 1: 51 walkers in colour 0, 50 in colour 1, like emcee's arange(n) % 2 shuffled
    assert len(p[0::2]) == 51 and len(p[1::2]) == 50


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    z = np.uint32(0)
    out = km.philox4x32_10(z, z, z, z, z, z)
    assert [int(v) for v in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.uint32(0xffffffff)
    out = km.philox4x32_10(f, f, f, f, f, f)
    assert [int(v) for v in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    out = km.philox4x32_10(np.uint32(0x243f6a88), np.uint32(0x85a308d3), np.uint32(0x13198a2e), np.uint32(0x03707344),
                           np.uint32(0xa4093822), np.uint32(0x299f31d0))
    assert [int(v) for v in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _run_cpu_sampler(W, ndim, steps, world=1, rank=0, group=None, seed=11):
    eng = km.GaussianToyLikelihood(ndim)
    s = EnsembleSampler(W, ndim, eng, seed=seed, world_size=world, rank=rank, group=group,
                        ops=km.NumpyStretchOps(), device="cpu")
    p0 = np.random.default_rng(5).normal(size=(W, ndim))
    s.initialize(p0)
    for _ in range(steps):
        s.step()
    return s, eng


def test_cpu_model_sampler_statistics():
    """Stretch move on a Gaussian target: acceptance in the usual range, variances recovered."""
    W, ndim = 256, 4
    eng = km.GaussianToyLikelihood(ndim)
    s = EnsembleSampler(W, ndim, eng, seed=3, ops=km.NumpyStretchOps(), device="cpu")
    p0 = np.random.default_rng(1).normal(size=(W, ndim))
    for _ in s.sample(p0, iterations=300, thin_by=1):
        pass
    acc = s.acceptance_fraction.mean()
    assert 0.3 < acc < 0.9
    chain = s.get_chain(discard=100, flat=True)
    assert np.allclose(chain.std(axis=0), eng.sig.numpy(), rtol=0.1)
    assert s.chain.shape == (W, 300, ndim)
    assert set(eng.calls[1:]) == {W // 2}          # every half-step evaluates half the ensemble in one call
    lp = s.get_log_prob()
    assert np.allclose(lp[-1], -0.5 * ((s.get_chain()[-1] / eng.sig.numpy()) ** 2).sum(axis=1))


def test_emcee_thin_conventions():
    W, ndim = 16, 3
    s = EnsembleSampler(W, ndim, km.GaussianToyLikelihood(ndim), seed=3, ops=km.NumpyStretchOps(), device="cpu")
    p0 = np.random.default_rng(1).normal(size=(W, ndim))
    n = sum(1 for _ in s.sample(p0, thin=5, iterations=20))
    assert n == 20 and s.get_chain().shape == (4, W, ndim)
    s.reset(W, ndim)
    n = sum(1 for _ in s.sample(p0, thin_by=5, iterations=4))
    assert n == 4 and s.get_chain().shape == (4, W, ndim)
    with pytest.raises(ValueError):
        EnsembleSampler(5, 3, km.GaussianToyLikelihood(3), ops=km.NumpyStretchOps(), device="cpu")
    with pytest.raises(ValueError):
        s.initialize(np.full((W, ndim), np.nan))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, W, ndim, steps, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, eng = _run_cpu_sampler(W, ndim, steps, world=world, rank=rank, group=dist.group.WORLD)
    q.put((rank, s.coords_host(), s.log_prob_host(), s.acceptance_fraction, eng.calls))
    dist.destroy_process_group()


@pytest.mark.parametrize("W", [64, 37])
def test_two_ranks_reproduce_single_rank_chain(W):
    """world_size 2 over gloo: both ranks end with the same ensemble, bit-identical to the 1-rank run
    (counter-based RNG), each evaluating only its slice."""
    import torch.multiprocessing as mp
    ndim, steps, world = 5, 6, 2
    ref, _ = _run_cpu_sampler(W, ndim, steps)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, W, ndim, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, coords, lp, acc, calls in res:
        assert np.array_equal(coords, ref.coords_host())
        assert np.array_equal(lp, ref.log_prob_host())
        assert np.array_equal(acc, ref.acceptance_fraction)
        assert max(calls) <= (W + 1) // 2 // world + 1 + (W // world)       # never the whole half-ensemble
    assert res[0][4][1] + res[1][4][1] == (W + 1) // 2                      # first half-step: slices add up


# ------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
def test_k6_kernels_match_numpy_model():
    from joxsz_b200.sampler import CudaStretchOps
    dev = torch.device("cuda", 0)
    ops = CudaStretchOps(dev)
    rng = np.random.default_rng(0)
    for nall, ndim in ((64, 13), (101, 7)):
        coords = rng.normal(size=(nall, ndim))
        lp = rng.normal(size=nall)
        perm = split_permutation(nall, 5, 9)
        perm_d = torch.zeros(nall, dtype=torch.int32, device=dev)
        ops.permutation(perm_d, 5, 9)
        assert np.array_equal(perm_d.cpu().numpy(), perm)
        big = torch.zeros(65536, dtype=torch.int32, device=dev)
        ops.permutation(big, (1 << 40) + 3, (1 << 35) + 1)
        assert np.array_equal(big.cpu().numpy(), split_permutation(65536, (1 << 40) + 3, (1 << 35) + 1))
        seed, it = 0x1234567890ABCDEF, (1 << 33) + 17
        for split in (0, 1):
            ns = (nall - split + 1) // 2
            first, count = 3, ns - 5
            prop_m, fac_m = km.stretch_propose(coords, perm, split, first, count, 2.0, seed, it)
            c_d, p_d = torch.from_numpy(coords).to(dev), torch.from_numpy(perm).to(dev)
            prop = torch.zeros((count, ndim), dtype=torch.float64, device=dev)
            fac = torch.zeros(count, dtype=torch.float64, device=dev)
            ops.propose(c_d, p_d, split, first, count, 2.0, seed, it, prop, fac)
            assert np.allclose(prop.cpu().numpy(), prop_m, rtol=0, atol=1e-14)
            assert np.allclose(fac.cpu().numpy(), fac_m, rtol=1e-14, atol=1e-15)
            lp_new = rng.normal(size=count)
            lp_new[::7] = -np.inf
            pk_m = km.stretch_accept(coords, lp, perm, split, first, count, prop.cpu().numpy(), lp_new,
                                     fac.cpu().numpy(), seed, it)
            packed = torch.zeros((count, ndim + 2), dtype=torch.float64, device=dev)
            ops.accept(c_d, torch.from_numpy(lp).to(dev), p_d, split, first, count, prop,
                       torch.from_numpy(lp_new).to(dev), fac, seed, it, packed)
            assert np.array_equal(packed.cpu().numpy(), pk_m)
            assert 0 < pk_m[:, -1].sum() < count
            # scatter of a full half
            full = rng.normal(size=(ns, ndim + 2))
            full[:, -1] = rng.integers(0, 2, ns)
            c2, l2, n2 = coords.copy(), lp.copy(), np.zeros(nall, dtype=np.int32)
            km.stretch_scatter(c2, l2, n2, perm, split, full, ns)
            cd, ld = torch.from_numpy(coords).to(dev), torch.from_numpy(lp).to(dev)
            nd = torch.zeros(nall, dtype=torch.int32, device=dev)
            ops.scatter(cd, ld, nd, p_d, split, torch.from_numpy(full).to(dev), ns)
            assert np.array_equal(cd.cpu().numpy(), c2) and np.array_equal(ld.cpu().numpy(), l2)
            assert np.array_equal(nd.cpu().numpy(), n2)


@pytest.mark.gpu
def test_device_sampler_on_the_cluster_likelihood(cl1226_fit, cl1226_oracle):
    """The device sampler on the real likelihood: log-probs carried by the chain equal the oracle's
    likelihood of the carried positions; detailed-balance bookkeeping (rejected walkers do not move)."""
    from helpers import orc
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.synthetic import draw_parameters
    eng = BatchedLikelihood(cl1226_fit, max_walkers=256)
    W = 64
    p0 = draw_parameters(cl1226_fit.thawed, n=W, seed=8, spread=0.01)
    s = EnsembleSampler(W, eng.ndim, eng, seed=21)
    s.initialize(p0)
    assert np.isfinite(s.log_prob_host()).all()
    before = s.coords_host().copy()
    s.step()
    after, lp = s.coords_host(), s.log_prob_host()
    acc = s.acceptance_fraction
    moved = np.any(after != before, axis=1)
    assert np.array_equal(moved, acc > 0)
    for _ in range(4):
        s.step()
    coords, lp = s.coords_host(), s.log_prob_host()
    ref = orc.BatchedOracle(cl1226_oracle).loglike(coords)
    assert np.max(np.abs(ref - lp)) < 1e-6
    assert 0.05 < acc.mean() <= 1.0
    # bound-method form, as joxsz_main.py:206 constructs the sampler
    s2 = EnsembleSampler(W, len(cl1226_fit.thawed), cl1226_fit.getLikelihood, pool=None, seed=21)
    s2.initialize(p0)
    for _ in range(5):
        s2.step()
    assert np.array_equal(s2.coords_host(), coords)
    eng.close()


@pytest.mark.gpu
def test_mcmc_run_schedule_end_to_end(cl1226_fit, cl1226_oracle, tmp_path):
    """The reference's driver sequence (joxsz_main.py:196-214): sampler on fit.getLikelihood, mcmc_run with its
    pre-fit / burn-in / sampling schedule, then the chain in emcee's layouts and the saved attributes."""
    from helpers import orc
    from joxsz_b200 import add_backend_attrs
    from joxsz_b200.sampler import mcmc_run
    from joxsz_b200.synthetic import FIDUCIAL
    fit = cl1226_fit
    saved = fit.thawedParVals()
    try:
        fit.updateThawed([FIDUCIAL[n] for n in fit.thawed])
        nwalkers, nburn, nlength, nthin = 30, 40, 60, 5            # joxsz_main.py:42-45 uses 30 / 2000 / 5000 / 5
        np.random.seed(7)
        mcmc = EnsembleSampler(nwalkers, len(fit.thawed), fit.getLikelihood, pool=None, backend=None, seed=7)
        mcmc.initspread = .1
        assert mcmc_run(mcmc, fit, nburn, nlength, nthin, max_prefit=1)
        cube_chain = mcmc.chain                                     # (nwalkers x niter x nparams)
        assert cube_chain.shape == (nwalkers, nlength // nthin, len(fit.thawed))
        flat_chain = cube_chain.reshape(-1, cube_chain.shape[2], order='F')
        assert np.isfinite(flat_chain).all()
        lp = mcmc.get_log_prob()
        assert lp.shape == (nlength // nthin, nwalkers) and np.isfinite(lp).all()
        # stored log-probs are the likelihood of the stored positions
        last = mcmc.get_chain()[-1]
        ref = orc.BatchedOracle(cl1226_oracle).loglike(last)
        assert np.max(np.abs(ref - lp[-1])) < 1e-6
        assert 0.02 < np.mean(mcmc.acceptance_fraction) < 0.9
        path = str(tmp_path / "chain.npz")
        mcmc.save_npz(path)
        add_backend_attrs(path, fit, nburn, nthin)
        z = np.load(path)
        assert z["chain"].shape == (nlength // nthin, nwalkers, len(fit.thawed))
        assert [k.decode() for k in z["param_names"]] == list(fit.thawed) and int(z["burn"]) == nburn
    finally:
        fit.updateThawed(saved)


def test_shuffle_and_proposal_streams_are_independent():
    """the colouring keys and the split-1 proposal draws once shared a Philox block, which made a walker's
    stretch factor a function of its rank in the permutation.  Every (purpose, split) pair has its own counter word."""
    k = np.arange(4096)
    seed, it = 99, 1234
    blocks = {}
    for purpose in (0, 1, 2):
        for split in (0, 1):
            if purpose == 2 and split == 1:
                continue
            r = km._draws(k, seed, it, split, purpose)
            blocks[(purpose, split)] = (r[0].astype(np.uint64) << np.uint64(32)) | r[1].astype(np.uint64)
    keys = list(blocks)
    for i in range(len(keys)):
        for j in range(i + 1, len(keys)):
            assert not np.any(blocks[keys[i]] == blocks[keys[j]]), (keys[i], keys[j])
    # z of the split-1 walkers does not depend on their position in the colouring permutation
    nall = 4096
    perm = split_permutation(nall, seed, it)
    pos = np.empty(nall, dtype=np.int64)
    pos[perm] = np.arange(nall)
    active = perm[1::2]
    u = blocks[(0, 1)][active].astype(np.float64) / 2.0 ** 64
    rho = np.corrcoef(u, pos[active])[0, 1]
    assert abs(rho) < 4.0 / np.sqrt(active.size)
    first = perm[1]                               # the smallest key of colour 1: its draw is an ordinary uniform
    assert blocks[(0, 1)][first] != blocks[(2, 0)][first]


@pytest.mark.gpu
def test_graph_replay_equals_eager_steps(cl1226_fit):
    """One iteration captured as a CUDA graph (device-side Philox counter) gives the chain of the launch-by-launch
    path bit for bit, including after the counter is moved and after a new ensemble is loaded."""
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.synthetic import draw_parameters
    eng = BatchedLikelihood(cl1226_fit, max_walkers=256)
    for W in (64, 51):
        p0 = draw_parameters(cl1226_fit.thawed, n=W, seed=8, spread=0.01)
        a = EnsembleSampler(W, eng.ndim, eng, seed=5, graph=True)
        b = EnsembleSampler(W, eng.ndim, eng, seed=5, graph=False)
        for s in (a, b):
            s.initialize(p0)
            for _ in range(7):
                s.step()
        assert a.graph_active and not b.graph_active and a._graph_failed is None
        assert np.array_equal(a.coords_host(), b.coords_host())
        assert np.array_equal(a.log_prob_host(), b.log_prob_host())
        assert np.array_equal(a.acceptance_fraction, b.acceptance_fraction)
        for s in (a, b):
            s.iteration += 1000                   # a jump of the counter (e.g. a restart that skips draws)
            s.initialize(p0 * (1 + 1e-4))
            for _ in range(3):
                s.step()
        assert np.array_equal(a.coords_host(), b.coords_host())
        assert np.array_equal(a.log_prob_host(), b.log_prob_host())
    eng.close()


def _nccl_worker(rank, world, port, W, steps, graph, exchange, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from joxsz_b200 import cluster
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.mb import mb
    from joxsz_b200.synthetic import draw_parameters
    mb.fit.debugfit = False
    inp = cluster.load_inputs_npz(os.path.join(ROOT, "tests", "golden", "cl1226_inputs.npz"))
    fit, _ = cluster.build_fit(inp, savedir=None)
    eng = BatchedLikelihood(fit, max_walkers=256, device=rank)
    p0 = draw_parameters(fit.thawed, n=W, seed=8, spread=0.01)
    s = EnsembleSampler(W, eng.ndim, eng, seed=5, world_size=world, rank=rank, group=dist.group.WORLD, graph=graph,
                        exchange=exchange)
    s.initialize(p0)
    for _ in range(steps):
        s.step()
    q.put((rank, s.coords_host(), s.log_prob_host(), s.acceptance_fraction, s.graph_active, s._graph_failed,
           s._px is not None))
    s.close()                  # the captured iteration holds NCCL collectives: it goes before the process group
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("graph", [False, True])
def test_two_nccl_ranks_reproduce_single_gpu_chain(cl1226_fit, graph, exchange):
    """Hardware proof of rank-count invariance: 2 processes / 2 GPUs end with the ensemble of the 1-GPU run, bit for
    bit -- eager launches and the captured graph, with the NCCL all-gather and with the accept kernel storing its rows
    into the peer's buffer over NVLink (odd half-ensembles: 35 walkers per colour over 2 ranks)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from joxsz_b200.batched import BatchedLikelihood
    from joxsz_b200.synthetic import draw_parameters
    W, steps, world = 70, 6, 2
    eng = BatchedLikelihood(cl1226_fit, max_walkers=256)
    ref = EnsembleSampler(W, eng.ndim, eng, seed=5, graph=False)
    ref.initialize(draw_parameters(cl1226_fit.thawed, n=W, seed=8, spread=0.01))
    for _ in range(steps):
        ref.step()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, W, steps, graph, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, coords, lp, acc, active, failed, p2p in res:
        assert active == graph, failed
        assert p2p == (exchange == "p2p")
        assert np.array_equal(coords, ref.coords_host())
        assert np.array_equal(lp, ref.log_prob_host())
        assert np.array_equal(acc, ref.acceptance_fraction)
    eng.close()
