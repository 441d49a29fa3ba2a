"""BASELINE.json configs[4] (multi-start: 1024 pulse seeds x 512 noise samples, 12-atom blockade chain, n = 377,
M = 256) on ONE GPU's share of it: `n_seeds` seeds evaluated (cost + gradient) through MultiStart, which runs several
seeds concurrently on separate streams so that the 64-CTA launches of one seed fill the 148 SMs.
usage: python profiles/cfg5_multistart.py [n_seeds=12] [n_concurrent=auto] [iters=2]
       torchrun --nproc-per-node G profiles/cfg5_multistart.py 24      (seeds sharded over G GPUs)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL as Chain  # noqa: E402
from quantum_optimal_control.multi_start import MultiStart  # noqa: E402

import torch.distributed as dist  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:                        # torchrun: the ranks split the seeds, no collective until the final MIN
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n_conc = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "auto" else None
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
M, S = 256, 512


def make():
    prop = Chain(16, M, 2e-6 / M, 12, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    nrng = np.random.default_rng(1234)
    prop.set_ensemble(del_total=nrng.normal(0, 2 * np.pi * 0.1e6, S), rabi_scale=1 + nrng.normal(0, 0.01, S))
    return prop


ms = MultiStart(make, n_concurrent=n_conc, shard=world > 1)
mine = ms.local_seeds(n_seeds)
abc = []
for seed in mine:
    rng = np.random.default_rng(100 + seed)
    a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
    abc.append(ms.replicas[0]._pack(a, b, c))
abc = torch.stack(abc)
out = ms.cost_and_grad(abc)          # warm-up: configures every replica
torch.cuda.synchronize()
ms.check()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(iters):
    out = ms.cost_and_grad(abc, out)
e1.record()
torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / iters
if world > 1:
    t = torch.tensor([dev_ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)      # slowest rank
    dev_ms = float(t)
best_seed, best_cost = ms.best(out[:, 0], n_seeds)
if world > 1 and dist.get_rank() != 0:
    dist.destroy_process_group()
    sys.exit(0)
units = n_seeds * M * S
print(f"cfg5 multi-start: {n_seeds} seeds, {ms.n_concurrent} concurrent, {dev_ms:.1f} ms per sweep over the seeds "
      f"({dev_ms / n_seeds:.1f} ms / seed), {units / dev_ms * 1e3:.3e} slice-samples/s; 1024 seeds: "
      f"{1024 * dev_ms / n_seeds / 1e3:.1f} s on {world} GPU(s); best seed {best_seed} cost {best_cost:.6f}",
      ms.replicas[0]._plan.chain_info())
