"""Multi-start optimisation (BASELINE.json configs[4]: many random pulse seeds x one robustness ensemble).

The reference runs its gradient descent from ONE random seed per notebook session (notebook 02 cell 13, notebook 04
cells 12-14: ``for i in range(num_iters): cost, a, b, c = single_optimization_step(...)``) and the user restarts it
by hand with another PRNG key.  Here the seeds are independent problems on the same propagator:

* across GPUs the seeds are block-sharded over the ranks (``shard_bounds``) with NO collective in the loop -- one
  ``all_gather`` of the final per-seed costs picks the winner;
* inside one GPU an evaluation of one seed may not fill the machine (a 512-sample ensemble of the 12-atom chain is
  64 CTAs of the samples-minor kernel on 148 SMs), so ``MultiStart`` keeps several replicas of the propagator (own
  plan and workspace each) and evaluates different seeds concurrently on separate CUDA streams.

Every evaluation goes through the same C-ABI call as a single-seed run (``qocvl_cost_grad``); results per seed are
bit-identical to evaluating the seeds one after the other.
"""
from __future__ import annotations

import numpy as np
import torch

from ._engine import shard_bounds


class MultiStart:
    """``make_propagator()`` must return a fresh, fully configured propagator (ensemble already set, NOT sharded:
    with multi-start the ranks split the seeds, not the samples)."""

    def __init__(self, make_propagator, n_concurrent=None, shard=False, sm_count=None):
        self._make = make_propagator
        self.replicas = [make_propagator()]
        self.device = self.replicas[0].device
        self._n_concurrent = n_concurrent
        self._sm_count = sm_count
        self.streams = []
        self.shard = shard
        if shard:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(), dist.get_rank()
        else:
            self.world, self.rank = 1, 0

    # -- replicas -------------------------------------------------------------------------------------------------
    def _ensure_replicas(self, abc0):
        if self.streams:
            return
        if self.device.type != "cuda":                         # host-logic tests: one replica, no streams
            self.streams = [None]
            return
        first = self.replicas[0]
        first.cost_and_grad_packed(abc0)                       # configures the launch: CTAs and workspace are known now
        torch.cuda.synchronize(self.device)
        k = self._n_concurrent
        if k is None:
            ctas = max(1, first._plan.chain_info()["ctas"])
            sms = self._sm_count or torch.cuda.get_device_properties(self.device).multi_processor_count
            k = max(1, min(4, -(-sms // ctas)))                 # enough CTAs in flight to cover every SM
            free, _total = torch.cuda.mem_get_info(self.device)
            need = max(1, first._plan.workspace_bytes)
            k = max(1, min(k, 1 + int(0.8 * free // need)))    # the first replica is already allocated
        for _ in range(k - 1):
            self.replicas.append(self._make())
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(k)]

    @property
    def n_concurrent(self):
        return len(self.streams) or (self._n_concurrent or 1)

    def local_seeds(self, n_seeds):
        """Seed indices this rank owns."""
        lo, hi = shard_bounds(n_seeds, self.world, self.rank)
        return range(lo, hi)

    # -- evaluation -----------------------------------------------------------------------------------------------
    def cost_and_grad(self, abc_seeds, out=None):
        """abc_seeds float64 CUDA [n_local, 3, N, C] -> [n_local, 3 + 3NC] (cost, infid, adiab, ga, gb, gc per seed).
        Stream-ordered with respect to the current stream; no host synchronisation."""
        n_local = abc_seeds.shape[0]
        if out is None:
            out = torch.empty((n_local, 3 + abc_seeds[0].numel()), dtype=torch.float64, device=self.device)
        if n_local == 0:
            return out
        self._ensure_replicas(abc_seeds[0])
        if self.streams[0] is None:
            for i in range(n_local):
                self.replicas[0].cost_and_grad_packed(abc_seeds[i], out[i])
            return out
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        for st in self.streams:
            st.wait_event(ready)
        for i in range(n_local):
            j = i % len(self.streams)
            with torch.cuda.stream(self.streams[j]):
                self.replicas[j].cost_and_grad_packed(abc_seeds[i], out[i])
        for st in self.streams:
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
        return out

    def optimize(self, abc_seeds, num_iters, lr=0.02):
        """Plain gradient descent of the notebooks from every seed at once; returns (cost history [num_iters, n_local],
        final parameters [n_local, 3, N, C]); all on the device."""
        abc = abc_seeds.clone()
        hist = torch.empty((num_iters, abc.shape[0]), dtype=torch.float64, device=self.device)
        out = None
        for it in range(num_iters):
            out = self.cost_and_grad(abc, out)
            hist[it] = out[:, 0]
            abc -= lr * out[:, 3:].view_as(abc)
        self.check()             # one sync after the loop: a seed whose pulses left the Taylor range raises (never NaN histories)
        return hist, abc

    def best(self, costs, n_seeds):
        """Global (seed index, cost) of the lowest final cost; one all_gather when sharded."""
        local = self.local_seeds(n_seeds)
        vals = costs.detach().to("cpu", torch.float64).numpy() if isinstance(costs, torch.Tensor) else np.asarray(costs)
        if len(vals) != len(local):
            raise ValueError(f"rank {self.rank} owns {len(local)} of {n_seeds} seeds, got {len(vals)} costs")
        table = np.full(n_seeds, np.inf)
        table[local.start:local.stop] = vals
        if self.shard and self.world > 1:
            import torch.distributed as dist
            t = torch.from_numpy(table).to(self.device if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            table = t.cpu().numpy()
        i = int(np.argmin(table))
        return i, float(table[i])

    def check(self):
        for r in self.replicas:
            r._plan.check()
