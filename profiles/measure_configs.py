"""Times one cost+gradient evaluation of every BASELINE.json config on one GPU (CUDA events, 3 warm-up + 5 timed
calls) and prints a small table.  Not the headline benchmark (that is bench.py = cfg 3); this documents the
other configs for DESIGN.md / profiles/README.md.

    python profiles/measure_configs.py > gpurun_out/configs.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "optimal-control-rydberg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(prop, a, b, c, reps=5, warm=3):
    abc = prop._pack(a, b, c)
    plan = prop._plan
    plan.set_timing(True)
    for _ in range(warm):
        prop.cost_and_grad_packed(abc)
    plan.check()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = prop.cost_and_grad_packed(abc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    kms = plan.last_chain_ms()
    # the same evaluation as an optimiser loop issues it: same buffers again and again, no kernel timing -> the library
    # replays the launch sequence as one cached CUDA graph (api.cu::run_evaluation)
    plan.set_timing(False)
    buf = torch.empty_like(out)
    for _ in range(warm + 3):
        prop.cost_and_grad_packed(abc, buf)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(4 * reps):
        prop.cost_and_grad_packed(abc, buf)
    e1.record()
    torch.cuda.synchronize()
    timed.last_repeat_ms = e0.elapsed_time(e1) / (4 * reps)
    assert torch.equal(buf, out)
    return ms, kms, plan.flops_per_call(True), float(out[0])


def main():
    from quantum_optimal_control.state_transfer.propagator_vl import PropagatorVL as ST
    from quantum_optimal_control.two_qubit.propagator_vl import PropagatorVL as TQ
    from quantum_optimal_control.blockade_chain.propagator_vl import PropagatorVL as Chain
    gam = [0.1, 1.0 / 118e-9, 1.0 / 1.5e-3, 1.0 / 170e-6]
    rows = []
    # cfg 1: notebook 01
    rng = np.random.RandomState(1)
    a = rng.uniform(-1, 1, (10, 5))
    b = np.concatenate([rng.uniform(-1, 1, (10, 1)), rng.uniform(0, 1, (10, 2)), rng.uniform(-1, 0, (10, 2))], axis=1)
    c = rng.uniform(0, 0.1, (10, 5))
    st = ST(10, 500, 4e-10, 0, 0, 2 * np.pi * 127e6, 2 * np.pi * 127e6, 1 / 150e-6, 1 / 118e-9)
    rows.append(("cfg1 state transfer (n=5, M=500, 1 sample)", 500, timed(st, a, b, c)))
    # cfg 2: notebook 02
    rng = np.random.default_rng(0)
    a, b, c = rng.uniform(-1, 1, (10, 5)), rng.uniform(-1, 1, (10, 5)), rng.uniform(0, 0.5, (10, 5))
    tq = TQ(10, 500, 15, 50e6, 648e-9 / 530, 0.0, 2 * np.pi * 10e6, 2 * np.pi * (-35.7e6), 2 * np.pi * 100e6,
            2 * np.pi * 100e6, gam)
    rows.append(("cfg2 CZ notebook grid (n=16, M=530, 1 sample)", 530, timed(tq, a, b, c)))
    # cfg 4: 10-atom chain, n = 144, M = 4000, T = 4 us
    rng = np.random.default_rng(7)
    a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
    ch = Chain(16, 4000, 1e-9, 10, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    rows.append(("cfg4 10-atom chain (n=144, M=4000, 1 sample)", 4000, timed(ch, a, b, c)))
    # cfg 5 (one seed): 12-atom chain, n = 377, M = 256, T = 2 us, 512 noise samples
    rng = np.random.default_rng(100)
    a, b, c = rng.uniform(-1, 1, (16, 2)), rng.uniform(-1, 1, (16, 2)), rng.uniform(0, 0.5, (16, 2))
    ch5 = Chain(16, 256, 2e-6 / 256, 12, 2 * np.pi * 2e6, 2 * np.pi * 4e6)
    nrng = np.random.default_rng(1234)
    ch5.set_ensemble(del_total=nrng.normal(0, 2 * np.pi * 0.1e6, 512), rabi_scale=1 + nrng.normal(0, 0.01, 512))
    rows.append(("cfg5 one seed: 12-atom chain (n=377, M=256, 512 samples)", 256 * 512, timed(ch5, a, b, c)))
    print(f"{'config':62s} {'ms/eval':>9s} {'chain ms':>9s} {'units/s':>11s} {'GFLOP/s':>9s}  cost")
    for name, units, (ms, kms, flops, cost) in rows:
        print(f"{name:62s} {ms:9.3f} {kms:9.3f} {units / ms * 1e3:11.3e} {flops / kms * 1e-6:9.1f}  {cost:.6f}")


if __name__ == "__main__":
    main()
