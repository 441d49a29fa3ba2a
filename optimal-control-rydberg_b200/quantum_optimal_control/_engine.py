"""Shared host logic of the two PropagatorVL mirrors: argument packing, the qocvl plan, the
torch.autograd bridge (stands in for ``jax.value_and_grad``, TQ:592 / notebook 01 cell 9), the
noise-ensemble and multi-GPU sharding extensions.  All numbers come from libqocvl.so."""
from __future__ import annotations

import numpy as np
import torch

import qocvl


def halving_flags(count):
    """ST:77-87 / TQ:144-154 ``gen_contraction_array``: odd/even flag of each halving level."""
    flags = []
    while count > 1:
        flags.append(bool(count & 1))
        count >>= 1
    return flags


def block_diag2(m):
    z = np.zeros_like(m)
    return np.block([[m, z], [z, m]])


def block_upper(m):
    z = np.zeros_like(m)
    return np.block([[z, m], [z, z]])


class _TargetFn(torch.autograd.Function):
    """cost(a, b, c) with the engine's analytic gradient."""

    @staticmethod
    def forward(ctx, prop, a, b, c):
        out = prop._allreduce(prop._plan.cost_grad(prop._pack(a, b, c)))    # sharded ensembles: the one all-reduce
        n, ch = prop.input_dim, prop.num_ctrls
        ctx.grads = out[3:].view(3, n, ch)
        ctx.like = (a, b, c)
        return out[0].clone()

    @staticmethod
    def backward(ctx, gout):
        res = []
        for i, ref in enumerate(ctx.like):
            g = (ctx.grads[i] * gout).to(device=ref.device, dtype=ref.dtype) if ctx.needs_input_grad[i + 1] else None
            res.append(g)
        return (None, *res)


class _WaveFn(torch.autograd.Function):
    """return_physical_amplitudes with its reverse pass (notebook 05 cells 8-12 differentiate it)."""

    @staticmethod
    def forward(ctx, prop, a, b, c):
        abc = prop._pack(a, b, c)
        ctx.prop, ctx.abc, ctx.like = prop, abc, (a, b, c)
        return prop._plan.physical_amplitudes(abc)

    @staticmethod
    def backward(ctx, gwave):
        g = ctx.prop._plan.physical_amplitudes_vjp(ctx.abc, gwave)
        res = [g[i].to(device=r.device, dtype=r.dtype) if ctx.needs_input_grad[i + 1] else None
               for i, r in enumerate(ctx.like)]
        return (None, *res)


class EngineBase:
    """Everything the two models share.  Subclasses fill in the physics in ``_build``."""

    num_ctrls = 5

    # -- set by subclasses --------------------------------------------------------------------
    _model = None
    W_INFID = W_ADIAB = None

    def _make_plan(self, *, n, n_slices, n_basis_pts, pad, gens_u, gens_w, coef_const, starts, targets,
                   adiab_index, fid_norm, scale, taps, eps=1e-12, device=None, symmetry=None):
        self._plan = qocvl.Plan(
            model=self._model, n=n, n_gen=gens_u.shape[0], n_slices=n_slices, n_basis_pts=n_basis_pts,
            pad=pad, n_gauss=self.input_dim, n_chan=self.num_ctrls, n_vec=len(starts),
            adiab_index=adiab_index, dt=self.delta_t, duration=self.duration, w_infid=self.W_INFID,
            w_adiab=self.W_ADIAB, fid_norm=fid_norm, eps=eps, scale=scale, device=device)
        self._plan.set_generators(gens_u, gens_w, coef_const)
        self._plan.set_filter(taps)
        self._plan.set_cost(starts, targets)
        if symmetry is not None:
            self._plan.set_symmetry(symmetry)
        self.device = self._plan.device
        self._world = None
        self._pin_in = self._pin_out = None

    # -- argument handling ----------------------------------------------------------------------
    def _pack(self, a, b, c):
        """(a, b, c) each [input_dim, 5] (numpy / torch, any device) -> float64 [3,N,5] on the GPU."""
        parts = []
        for x in (a, b, c):
            t = x.detach() if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))
            if tuple(t.shape) != (self.input_dim, self.num_ctrls):
                raise ValueError(f"control arrays must have shape ({self.input_dim}, {self.num_ctrls}), "
                                 f"got {tuple(t.shape)}")
            parts.append(t.to(device=self.device, dtype=torch.float64))
        return torch.stack(parts, dim=0).contiguous()

    @staticmethod
    def _wants_grad(*xs):
        return torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad for x in xs)

    # -- the reference's method surface (ST:215-310, TQ:399-574) ------------------------------------
    def return_physical_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        if self._wants_grad(ctrl_a, ctrl_b, ctrl_c):
            return _WaveFn.apply(self, ctrl_a, ctrl_b, ctrl_c)
        return self._plan.physical_amplitudes(self._pack(ctrl_a, ctrl_b, ctrl_c))

    def return_auxiliary_amplitudes(self, ctrl_a, ctrl_b, ctrl_c):
        return self._plan.auxiliary_amplitudes(self._pack(ctrl_a, ctrl_b, ctrl_c))

    def exponentials(self, ctrl_a, ctrl_b, ctrl_c):
        return self._plan.exponentials(self._pack(ctrl_a, ctrl_b, ctrl_c))

    def propagate(self, ctrl_a, ctrl_b, ctrl_c):
        return self._plan.propagate(self._pack(ctrl_a, ctrl_b, ctrl_c))

    def metrics(self, ctrl_a, ctrl_b, ctrl_c):
        out = self._plan.cost(self._pack(ctrl_a, ctrl_b, ctrl_c))
        out = self._allreduce(out)
        return out[1], out[2]

    def target(self, ctrl_a, ctrl_b, ctrl_c):
        if self._wants_grad(ctrl_a, ctrl_b, ctrl_c):      # also when sharded: the all-reduce sits inside _TargetFn.forward
            return _TargetFn.apply(self, ctrl_a, ctrl_b, ctrl_c)
        return self._allreduce(self._plan.cost(self._pack(ctrl_a, ctrl_b, ctrl_c)))[0]

    # -- gradient callbacks -----------------------------------------------------------------------
    def cost_and_grad_packed(self, abc, out=None):
        """Device in / device out, no host sync: float64 [3+3NC] = cost, infid, adiab, ga, gb, gc."""
        return self._allreduce(self._plan.cost_grad(abc, out))

    def target_and_grad(self, ctrl_a, ctrl_b, ctrl_c):
        """Stand-in for ``jax.value_and_grad(target, argnums=(0,1,2))``: -> (cost, (ga, gb, gc)),
        all CUDA tensors, gradients shaped like the inputs (per-array masks apply, notebook 04 cell 5)."""
        out = self.cost_and_grad_packed(self._pack(ctrl_a, ctrl_b, ctrl_c))
        g = out[3:].view(3, self.input_dim, self.num_ctrls)
        return out[0], (g[0], g[1], g[2])

    def value_and_grad_host(self, ctrl_a, ctrl_b, ctrl_c):
        """Host optimiser callback (e.g. scipy.optimize.minimize(jac=True)): NumPy in, NumPy out.
        One pinned host->device copy of the 3*N*C parameters and one device->host copy of
        3+3*N*C results per call."""
        n = 3 * self.input_dim * self.num_ctrls
        if self._pin_in is None:
            self._pin_in = torch.empty((3, self.input_dim, self.num_ctrls), dtype=torch.float64).pin_memory()
            self._pin_out = torch.empty(3 + n, dtype=torch.float64).pin_memory()
            self._dev_in = torch.empty_like(self._pin_in, device=self.device)
        for i, x in enumerate((ctrl_a, ctrl_b, ctrl_c)):
            self._pin_in[i].copy_(torch.as_tensor(np.asarray(x, dtype=np.float64)))
        self._dev_in.copy_(self._pin_in, non_blocking=True)
        out = self.cost_and_grad_packed(self._dev_in)
        self._pin_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        res = self._pin_out.numpy()
        if not np.isfinite(res[0]):          # a slice left the Taylor range: the engine poisons the result and flags it
            self._plan.check()               # -> QocvlError(ERR_NORM); never hand NaNs to a host optimiser silently
        g = res[3:].reshape(3, self.input_dim, self.num_ctrls)
        return float(res[0]), (g[0].copy(), g[1].copy(), g[2].copy())

    # -- extensions: robustness ensemble and multi-GPU sharding ---------------------------------------
    def _noise_tables(self, del_total, rabi_scale):
        raise NotImplementedError

    def set_ensemble(self, del_total=None, rabi_scale=None, shard=False):
        """Robustness ensemble: sample s behaves as if the constructor had been called with
        ``del_total = del_total[s]`` and both Rabi frequencies multiplied by ``rabi_scale[s]``;
        ``target`` / ``metrics`` / gradients become ensemble means.  ``shard=True`` keeps only this
        rank's block of samples (torch.distributed must be initialised) and completes every
        evaluation with ONE all-reduce of 3+3*N*C doubles."""
        if del_total is None and rabi_scale is None:
            self._plan.set_noise(None, None)
            self._world = None
            return
        ref = del_total if del_total is not None else rabi_scale
        S = len(ref)
        dl = np.full(S, float(self.del_total)) if del_total is None else np.asarray(del_total, float)
        rs = np.ones(S) if rabi_scale is None else np.asarray(rabi_scale, float)
        mul, add = self._noise_tables(dl, rs)
        if shard:
            import torch.distributed as dist
            world, rank = dist.get_world_size(), dist.get_rank()
            if S < world:        # identical on every rank, so every rank raises (no rank is left waiting in a collective)
                raise ValueError(f"cannot shard {S} noise samples over {world} ranks: every rank needs at least one")
            lo, hi = shard_bounds(S, world, rank)
            mul, add = mul[lo:hi], add[lo:hi]
            self._world = world
        else:
            self._world = None
        self._plan.set_noise(mul, add, global_samples=S)

    def set_precision(self, dtype="complex128"):
        """north_star's optional complex64 mode: the ensemble kernel (reduced stacked system) carries state, Taylor
        terms and matrix entries in complex64; pulse map, coefficients and reductions stay float64.  Stated bounds
        against complex128: cost 5e-5 absolute, gradient 1e-2 of its largest entry.  The
        reference itself is complex128 only (TQ:1-12 enables x64)."""
        self._plan.set_precision(dtype)

    def _allreduce(self, out):
        if self._world is not None and self._world > 1:
            import torch.distributed as dist
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out

    # -- the optimiser loop of the notebooks, kept on the device ------------------------------------------
    def optimize(self, ctrl_a, ctrl_b, ctrl_c, num_iters, lr, masks=None, use_graph=True, method="gd",
                 betas=(0.9, 0.999), eps=1e-8, history=8):
        """``num_iters`` plain gradient-descent steps (notebook 01 cell 9, notebook 02 cell 13:
        ``ctrl -= lr * grad`` with the cost recorded BEFORE each update, TQ:580-600) without a host
        synchronisation per step.  ``masks`` = optional (mask_a, mask_b, mask_c) multiplied into the
        gradients (notebook 04 cells 5, 12: frozen columns).  One step is captured in a CUDA graph and
        replayed; the cost history stays on the device.  ``method="adam"`` replaces the update by Adam (Kingma & Ba;
        bias-corrected moments, the masks multiply the final step so frozen entries stay frozen) -- the reference only
        has plain descent; same loop, same graph capture.  ``method="lbfgs"``: limited-memory BFGS direction (two-loop
        recursion over the last ``history`` (s, y) pairs, initial scaling s.y / y.y, pairs with s.y <= 0 skipped, the
        first step is plain descent) with the fixed step ``lr`` instead of a line search, so that the whole step stays on
        the device and inside the same CUDA graph; the pair buffers are shifted, not indexed, and empty slots carry
        rho = 0, which makes every replay the same launch sequence.

        Returns (cost_history [num_iters] CUDA tensor, a, b, c) with a, b, c CUDA tensors [input_dim, C]."""
        if method not in ("gd", "adam", "lbfgs"):
            raise ValueError("method must be 'gd', 'adam' or 'lbfgs'")
        dev = self.device
        abc = self._pack(ctrl_a, ctrl_b, ctrl_c).clone()
        n = 3 * self.input_dim * self.num_ctrls
        out = torch.empty(3 + n, dtype=torch.float64, device=dev)
        hist = torch.zeros(max(1, num_iters), dtype=torch.float64, device=dev)
        step_idx = torch.zeros(1, dtype=torch.long, device=dev)
        if masks is None:
            scale = torch.full((3, self.input_dim, self.num_ctrls), float(lr), dtype=torch.float64, device=dev)
        else:
            scale = float(lr) * torch.stack([torch.as_tensor(np.asarray(m, dtype=np.float64)) for m in masks]).to(dev)
            scale = scale.expand(3, self.input_dim, self.num_ctrls).contiguous()

        m1, m2 = torch.zeros_like(abc), torch.zeros_like(abc)
        pow1 = torch.ones(1, dtype=torch.float64, device=dev)      # beta1^t, beta2^t kept on the device (graph replay)
        pow2 = torch.ones(1, dtype=torch.float64, device=dev)

        hm = max(1, int(history))
        nflat = abc.numel()
        s_buf = torch.zeros(hm, nflat, dtype=torch.float64, device=dev)     # newest pair in row 0
        y_buf = torch.zeros(hm, nflat, dtype=torch.float64, device=dev)
        rho = torch.zeros(hm, dtype=torch.float64, device=dev)
        x_prev, g_prev = torch.zeros(nflat, dtype=torch.float64, device=dev), torch.zeros(nflat, dtype=torch.float64, device=dev)
        have_prev = torch.zeros(1, dtype=torch.float64, device=dev)
        mask_flat = (scale != 0).to(torch.float64).reshape(-1)

        def lbfgs_direction(g):
            # new pair from the previous iterate (ignored on the first step and when the curvature condition fails)
            s_new, y_new = (abc.reshape(-1) - x_prev) * have_prev, (g - g_prev) * have_prev
            sy = (s_new * y_new).sum()     # (no cuBLAS inside the captured step: plain reductions)
            ok = (sy > 1e-300).to(torch.float64)
            s_buf.copy_(torch.roll(s_buf, 1, 0)); y_buf.copy_(torch.roll(y_buf, 1, 0)); rho.copy_(torch.roll(rho, 1, 0))
            s_buf[0].copy_(s_new * ok); y_buf[0].copy_(y_new * ok)
            rho[0:1].copy_((ok / torch.where(sy > 1e-300, sy, torch.ones_like(sy))).reshape(1))
            x_prev.copy_(abc.reshape(-1)); g_prev.copy_(g); have_prev.fill_(1.0)
            q = g.clone()
            alphas = []
            for i in range(hm):                           # newest to oldest
                a_i = rho[i] * (s_buf[i] * q).sum()
                q = q - a_i * y_buf[i]
                alphas.append(a_i)
            # initial Hessian scaling from the newest VALID pair (1 when there is none)
            yy = (y_buf * y_buf).sum(1)
            gam_all = torch.where(rho > 0, 1.0 / (rho * torch.where(yy > 0, yy, torch.ones_like(yy))), torch.zeros_like(rho))
            first = torch.cumsum((rho > 0).to(torch.float64), 0)
            pick = ((rho > 0) & (first == 1)).to(torch.float64)
            gamma = (gam_all * pick).sum() + (1.0 - pick.sum())
            r = gamma * q
            for i in reversed(range(hm)):                 # oldest to newest
                b_i = rho[i] * (y_buf[i] * r).sum()
                r = r + s_buf[i] * (alphas[i] - b_i)
            return r

        def one_step():
            self.cost_and_grad_packed(abc, out)
            hist.index_copy_(0, step_idx, out[0:1])
            g = out[3:].view_as(abc)
            if method == "gd":
                abc.sub_(scale * g)
            elif method == "lbfgs":
                d = lbfgs_direction(g.reshape(-1) * mask_flat)
                abc.sub_(scale * d.view_as(abc))
            else:
                m1.mul_(betas[0]).add_(g, alpha=1.0 - betas[0])
                m2.mul_(betas[1]).addcmul_(g, g, value=1.0 - betas[1])
                pow1.mul_(betas[0]); pow2.mul_(betas[1])
                abc.sub_(scale * (m1 / (1.0 - pow1)) / ((m2 / (1.0 - pow2)).sqrt() + eps))
            step_idx.add_(1)

        graph_ok = use_graph and self._world is None and num_iters > 2
        if graph_ok:
            self.cost_and_grad_packed(abc, out)             # configure the plan outside of the capture
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    one_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            for _ in range(num_iters):
                graph.replay()
        else:
            for _ in range(num_iters):
                one_step()
        self._plan.check()       # one sync after the loop: a slice that left the Taylor range raises instead of returning NaNs
        return hist[:num_iters], abc[0], abc[1], abc[2]


    # -- adiabatic seeding (notebook 05 cells 8-12): fit the pulses THROUGH the pulse map, on the device ----------
    def match_amplitudes(self, targets, ctrl_a, ctrl_b, ctrl_c, num_iters=1000, lr=0.03, use_graph=True):
        """Gradient-descent fit of selected physical-amplitude channels to target shapes, the notebook's
        ``cost_match_rabi`` / ``single_match_step``:  cost = mean over the chosen channels of
        mean_k (wave[k, ch] - target_ch[k])**2  (NB05 cell 10 uses channels 1 and 3, the two Rabi magnitudes),
        ``ctrl -= lr * grad`` with the cost recorded BEFORE each update.  ``targets``: {channel: array [M]}.
        The pulse map and its reverse pass are the CUDA kernels of the hot path's first stage
        (``qocvl_physical_amplitudes`` / ``qocvl_physical_amplitudes_vjp``); one step is captured in a CUDA
        graph and replayed, with no host synchronisation per step.

        Returns (cost_history [num_iters] CUDA tensor, a, b, c) with a, b, c CUDA tensors [input_dim, C]."""
        dev = self.device
        abc = self._pack(ctrl_a, ctrl_b, ctrl_c).clone()
        M = self._plan.desc.n_slices
        if not targets:
            raise ValueError("targets must name at least one channel")
        tgt = torch.zeros((M, self.num_ctrls), dtype=torch.float64, device=dev)
        wsel = torch.zeros((1, self.num_ctrls), dtype=torch.float64, device=dev)
        for ch, arr in targets.items():
            if not 0 <= int(ch) < self.num_ctrls:
                raise ValueError(f"channel {ch} out of range")
            t = torch.as_tensor(np.asarray(arr, dtype=np.float64))
            if tuple(t.shape) != (M,):
                raise ValueError(f"target of channel {ch} must have shape ({M},), got {tuple(t.shape)}")
            tgt[:, int(ch)] = t.to(dev)
            wsel[0, int(ch)] = 1.0 / (len(targets) * M)
        hist = torch.zeros(max(1, num_iters), dtype=torch.float64, device=dev)
        step_idx = torch.zeros(1, dtype=torch.long, device=dev)

        def one_step():
            diff = (self._plan.physical_amplitudes(abc) - tgt) * wsel          # (wave - target) / (n_ch * M)
            hist.index_copy_(0, step_idx, ((diff * diff).sum() / wsel.max()).view(1))
            abc.sub_(float(lr) * self._plan.physical_amplitudes_vjp(abc, 2.0 * diff).view_as(abc))
            step_idx.add_(1)

        if use_graph and num_iters > 2:
            self._plan.physical_amplitudes(abc)               # configure the plan outside of the capture
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    one_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            for _ in range(num_iters):
                graph.replay()
        else:
            for _ in range(num_iters):
                one_step()
        self._plan.check()
        return hist[:num_iters], abc[0], abc[1], abc[2]


def shard_bounds(n_items, world, rank):
    """Contiguous block partition of n_items over world ranks (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gradient_step(prop, ctrl_a, ctrl_b, ctrl_c, lr):
    """TQ:580-600 ``single_optimization_step``: plain gradient descent, returns (cost, a', b', c')."""
    cost, (ga, gb, gc) = prop.target_and_grad(ctrl_a, ctrl_b, ctrl_c)

    def upd(x, g):
        if isinstance(x, torch.Tensor):
            return x.detach() - lr * g.to(device=x.device, dtype=x.dtype)
        return torch.as_tensor(np.asarray(x, dtype=np.float64), device=g.device) - lr * g

    return cost, upd(ctrl_a, ga), upd(ctrl_b, gb), upd(ctrl_c, gc)
