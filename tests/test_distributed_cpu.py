"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sample sharding, the 'partial sums already
divided by the GLOBAL sample count' convention of qocvl_cost_grad, and the single all-reduce that
completes an evaluation.  The CUDA plan is replaced by a stub that mimics the C-ABI contract, so the
exact code path of quantum_optimal_control._engine (set_ensemble(shard=True) -> cost_and_grad_packed)
runs under torch.distributed without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quantum_optimal_control import _engine

N_GAUSS, N_CHAN, S_TOTAL = 3, 5, 7           # 7 samples over 2 ranks: ragged split 4 + 3


class StubPlan:
    """Same contract as qocvl.Plan for the calls _engine makes: per-sample value f(mul, add), outputs are
    sums over the LOCAL samples divided by the GLOBAL sample count."""
    device = torch.device("cpu")

    def __init__(self):
        self.mul = np.ones((1, 11))
        self.add = np.zeros((1, 11), complex)
        self.global_samples = 1

    def set_noise(self, mul=None, add=None, global_samples=None):
        self.mul = np.ones((1, 11)) if mul is None else np.asarray(mul)
        self.add = np.zeros((1, 11), complex) if add is None else np.asarray(add)
        self.global_samples = self.mul.shape[0] if global_samples is None else global_samples

    def _per_sample(self, abc):
        x = abc.numpy().ravel()
        rows = []
        for m, a in zip(self.mul, self.add):
            key = m[3] + 1e-6 * a[1].imag
            rows.append(np.concatenate([[np.sin(key), key, key ** 2], key * x]))
        return np.array(rows)

    def cost_grad(self, abc, out=None):
        res = torch.from_numpy(self._per_sample(abc).sum(axis=0) / self.global_samples)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def cost(self, abc, out=None):
        return self.cost_grad(abc)[:3].clone()

    def check(self):
        pass


class StubEngine(_engine.EngineBase):
    input_dim, num_ctrls, del_total = N_GAUSS, N_CHAN, 0.0

    def __init__(self):
        self._plan = StubPlan()
        self.device = torch.device("cpu")
        self._world = None

    def _noise_tables(self, del_total, rabi_scale):
        mul = np.ones((len(del_total), 11))
        mul[:, 3:7] = rabi_scale[:, None]
        add = np.zeros((len(del_total), 11), complex)
        add[:, 1] = -1j * del_total
        return mul, add


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    dl, rs = rng.normal(0, 1e5, S_TOTAL), 1 + rng.normal(0, 0.01, S_TOTAL)
    a, b, c = (rng.uniform(-1, 1, (N_GAUSS, N_CHAN)) for _ in range(3))
    eng = StubEngine()
    eng.set_ensemble(del_total=dl, rabi_scale=rs, shard=True)
    lo, hi = _engine.shard_bounds(S_TOTAL, world, rank)
    assert eng._plan.mul.shape[0] == hi - lo and eng._plan.global_samples == S_TOTAL
    out = eng.cost_and_grad_packed(eng._pack(a, b, c))
    cost, (ga, gb, gc) = eng.target_and_grad(a, b, c)
    infid, adiab = eng.metrics(a, b, c)
    # torch-autograd callers of a SHARDED engine get a differentiable cost whose gradient is the ensemble mean
    # (round-1 bug: target() silently fell through to the non-differentiable branch when sharded)
    ta, tb, tc = (torch.tensor(x, requires_grad=True) for x in (a, b, c))
    tcost = eng.target(ta, tb, tc)
    assert tcost.requires_grad
    tcost.backward()
    assert abs(float(tcost) - float(cost)) < 1e-15 and np.allclose(ta.grad.numpy(), ga.numpy(), rtol=1e-14)
    # fewer samples than ranks: the same ValueError on every rank, before any collective (no hang)
    with pytest.raises(ValueError, match="cannot shard"):
        StubEngine().set_ensemble(del_total=dl[:1], rabi_scale=rs[:1], shard=True)
    queue.put((rank, out.numpy().copy(), float(cost), float(infid), float(adiab), ga.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_mean_equals_single_process_mean():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: all 7 samples on one "device"
    rng = np.random.default_rng(11)
    dl, rs = rng.normal(0, 1e5, S_TOTAL), 1 + rng.normal(0, 0.01, S_TOTAL)
    a, b, c = (rng.uniform(-1, 1, (N_GAUSS, N_CHAN)) for _ in range(3))
    eng = StubEngine()
    eng.set_ensemble(del_total=dl, rabi_scale=rs)
    full = eng.cost_and_grad_packed(eng._pack(a, b, c)).numpy()
    for rank, out, cost, infid, adiab, ga in results:
        assert np.allclose(out, full, rtol=1e-14, atol=1e-15), rank       # every rank holds the full mean
        assert abs(cost - full[0]) < 1e-14 and abs(infid - full[1]) < 1e-14 and abs(adiab - full[2]) < 1e-14
        assert np.allclose(ga.ravel(), full[3:3 + N_GAUSS * N_CHAN], rtol=1e-14)


def _multistart_worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quantum_optimal_control.multi_start import MultiStart
    n_seeds = 5                                             # ragged: 3 + 2

    def make():
        eng = StubEngine()
        eng.set_ensemble(del_total=np.linspace(-1e5, 1e5, 4), rabi_scale=np.linspace(0.99, 1.01, 4))   # not sharded
        return eng

    ms = MultiStart(make, shard=True)
    mine = ms.local_seeds(n_seeds)
    seeds = torch.stack([torch.from_numpy(np.random.default_rng(40 + k).uniform(-1, 1, (3, N_GAUSS, N_CHAN)))
                         for k in mine])
    out = ms.cost_and_grad(seeds)
    hist, final = ms.optimize(seeds, num_iters=3, lr=0.1)
    best = ms.best(hist[-1], n_seeds)
    queue.put((rank, list(mine), out.numpy().copy(), hist.numpy().copy(), best))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_multi_start_shards_seeds_without_collective_in_the_loop():
    """MultiStart(shard=True): the ranks split the SEEDS (3 + 2 of 5), every seed sees the whole ensemble, per-seed
    results equal a single-process evaluation, and `best` agrees on both ranks (one MIN all-reduce at the end)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_multistart_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((queue.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == [0, 1, 2] and results[1][1] == [3, 4]
    eng = StubEngine()
    eng.set_ensemble(del_total=np.linspace(-1e5, 1e5, 4), rabi_scale=np.linspace(0.99, 1.01, 4))
    finals = []
    for rank, mine, out, hist, best in results:
        for row, k in zip(out, mine):
            abc = torch.from_numpy(np.random.default_rng(40 + k).uniform(-1, 1, (3, N_GAUSS, N_CHAN)))
            assert np.allclose(row, eng.cost_and_grad_packed(abc).numpy(), rtol=1e-14, atol=1e-15)
        finals.extend(hist[-1].tolist())
    assert results[0][4] == results[1][4] == (int(np.argmin(finals)), float(np.min(finals)))


def test_shard_bounds_cover_ragged_and_tiny_cases():
    for n, w in ((7, 2), (4096, 8), (3, 8), (0, 2)):
        spans = [_engine.shard_bounds(n, w, r) for r in range(w)]
        assert sum(hi - lo for lo, hi in spans) == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
