"""Plan: a thin object around one ``qocvl_plan`` handle.  PyTorch supplies device memory and the
current CUDA stream; all arithmetic happens in libqocvl.so."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _host_c128(x):
    """contiguous complex128 numpy array viewed as interleaved doubles (keeps a reference alive)."""
    arr = np.ascontiguousarray(np.asarray(x, dtype=np.complex128))
    return arr, arr.ctypes.data_as(C.c_void_p)


def _host_f64(x):
    arr = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    return arr, arr.ctypes.data_as(C.c_void_p)


class options:
    """``with qocvl.options(store=0, tp_chunk=7): ...`` -- explicit switches (``qocvl_set_option``) applied to every Plan
    CREATED inside the block.  They exist so that each kernel path can be cross-checked against the others; the
    library itself reads no environment variables."""
    _active: dict = {}

    def __init__(self, **kv):
        self.kv = {str(k): int(v) for k, v in kv.items()}

    def __enter__(self):
        self._saved = dict(options._active)
        options._active = {**options._active, **self.kv}
        return self

    def __exit__(self, *exc):
        options._active = self._saved


class Plan:
    """Owns the device-side constants (generators, cost kets, filter taps, noise tables) and the
    workspaces of one problem on one GPU.  Not thread-safe; one Plan per device / rank."""

    def __init__(self, *, model, n, n_gen, n_slices, n_basis_pts, pad, n_gauss, n_chan, n_vec,
                 adiab_index, dt, duration, w_infid, w_adiab, fid_norm, eps=1e-12, scale=(0, 0, 0, 0),
                 device=None):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("qocvl.Plan needs a CUDA device (no CPU fallback exists)")
        index = None if device is None else torch.device(device).index
        self.device = torch.device("cuda", torch.cuda.current_device() if index is None else index)
        d = _lib.Desc()
        d.model, d.n, d.n_gen, d.n_slices, d.n_basis_pts, d.pad = model, n, n_gen, n_slices, n_basis_pts, pad
        d.n_gauss, d.n_chan, d.n_vec, d.adiab_index = n_gauss, n_chan, n_vec, adiab_index
        d.dt, d.duration, d.w_infid, d.w_adiab, d.fid_norm, d.eps = dt, duration, w_infid, w_adiab, fid_norm, eps
        for i, s in enumerate(scale):
            d.scale[i] = float(s)
        self.desc = d
        self._h = C.c_void_p()
        _lib.check(self._lib.qocvl_plan_create(C.byref(self._h), C.byref(d), self.device.index), "plan_create")
        self.n_samples = 1
        self.n_grad = 3 * n_gauss * n_chan
        for key, value in options._active.items():
            self.set_option(key, value)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.qocvl_plan_destroy(h)
            except Exception:
                pass

    # ---- one-time setup -------------------------------------------------------------------
    def set_generators(self, gens_u, gens_w, coef_const):
        gu, pu = _host_c128(gens_u)
        gw, pw = _host_c128(gens_w)
        cc, pc = _host_c128(coef_const)
        d = self.desc
        if gu.shape != (d.n_gen, d.n, d.n) or gw.shape != gu.shape or cc.shape != (d.n_gen,):
            raise ValueError("generator blocks must be [J,n,n] and coef_const [J]")
        _lib.check(self._lib.qocvl_set_generators(self._h, pu, pw, pc), "set_generators")

    def set_filter(self, taps):
        if taps is None:
            _lib.check(self._lib.qocvl_set_filter(self._h, None), "set_filter")
            return
        t, pt = _host_f64(taps)
        if t.shape != (self.desc.n_slices,):
            raise ValueError("filter taps must have n_slices entries")
        _lib.check(self._lib.qocvl_set_filter(self._h, pt), "set_filter")

    def set_cost(self, starts, targets):
        s, ps = _host_c128(starts)
        t, pt = _host_c128(targets)
        d = self.desc
        if s.shape != (d.n_vec, d.n) or t.shape != s.shape:
            raise ValueError("start / target kets must be [n_vec, n]")
        _lib.check(self._lib.qocvl_set_cost(self._h, ps, pt), "set_cost")

    def set_symmetry(self, perm):
        """Basis permutation that leaves every generator's diagonal block invariant (verified by the
        engine); start kets it maps onto each other are propagated once.  None clears the hint."""
        if perm is None:
            _lib.check(self._lib.qocvl_set_symmetry(self._h, None), "set_symmetry")
            return
        arr = np.ascontiguousarray(np.asarray(perm, dtype=np.int32))
        if arr.shape != (self.desc.n,):
            raise ValueError("perm must have n entries")
        _lib.check(self._lib.qocvl_set_symmetry(self._h, arr.ctypes.data_as(C.c_void_p)), "set_symmetry")

    def set_noise(self, mul=None, add=None, global_samples=None):
        """mul [S,J] real, add [S,J] complex (per-sample coefficient c*mul+add); None = nominal."""
        if mul is None and add is None:
            _lib.check(self._lib.qocvl_set_noise(self._h, 1, None, None), "set_noise")
            self.n_samples = 1
        else:
            J = self.desc.n_gen
            S = (np.asarray(mul) if mul is not None else np.asarray(add)).shape[0]
            m, pm = _host_f64(np.ones((S, J)) if mul is None else mul)
            a, pa = _host_c128(np.zeros((S, J)) if add is None else add)
            if m.shape != (S, J) or a.shape != (S, J):
                raise ValueError("noise tables must be [S, J]")
            _lib.check(self._lib.qocvl_set_noise(self._h, S, pm, pa), "set_noise")
            self.n_samples = S
        if global_samples is not None:
            _lib.check(self._lib.qocvl_set_shard(self._h, int(global_samples)), "set_shard")

    def set_option(self, name, value):
        """Explicit per-plan switch (``qocvl_set_option``; names in include/qocvl.h); a negative value = automatic."""
        _lib.check(self._lib.qocvl_set_option(self._h, str(name).encode(), int(value)), f"set_option({name})")

    def set_precision(self, dtype):
        """'complex128' (default) or 'complex64': arithmetic type of the ensemble kernel (qocvl_set_precision)."""
        code = {"complex128": 0, "c128": 0, "complex64": 1, "c64": 1}.get(str(dtype).replace("torch.", ""))
        if code is None:
            raise ValueError("dtype must be 'complex128' or 'complex64'")
        _lib.check(self._lib.qocvl_set_precision(self._h, code), "set_precision")

    @property
    def precision(self):
        return ("complex128", "complex64")[int(self._lib.qocvl_precision(self._h))]

    # ---- hot path -------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _abc(self, abc):
        if abc.device != self.device or abc.dtype != torch.float64 or not abc.is_contiguous():
            raise ValueError("abc must be a contiguous float64 tensor [3,N,C] on the plan's device")
        d = self.desc
        if tuple(abc.shape) != (3, d.n_gauss, d.n_chan):
            raise ValueError(f"abc must have shape (3, {d.n_gauss}, {d.n_chan}), got {tuple(abc.shape)}")
        return C.c_void_p(abc.data_ptr())

    def _new(self, shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def physical_amplitudes(self, abc):
        out = self._new((self.desc.n_slices, self.desc.n_chan))
        _lib.check(self._lib.qocvl_physical_amplitudes(self._h, self._abc(abc), C.c_void_p(out.data_ptr()),
                                                       self._stream()), "physical_amplitudes")
        return out

    def physical_amplitudes_vjp(self, abc, gwave):
        gwave = gwave.to(self.device, torch.float64).contiguous()
        out = self._new((3, self.desc.n_gauss, self.desc.n_chan))
        _lib.check(self._lib.qocvl_physical_amplitudes_vjp(
            self._h, self._abc(abc), C.c_void_p(gwave.data_ptr()), C.c_void_p(out.data_ptr()),
            self._stream()), "physical_amplitudes_vjp")
        return out

    def auxiliary_amplitudes(self, abc):
        out = self._new((self.desc.n_slices, self.desc.n_gen - 2), torch.complex128)
        _lib.check(self._lib.qocvl_auxiliary_amplitudes(self._h, self._abc(abc), C.c_void_p(out.data_ptr()),
                                                        self._stream()), "auxiliary_amplitudes")
        return out

    def exponentials(self, abc):
        D = 2 * self.desc.n
        out = self._new((self.desc.n_slices, D, D), torch.complex128)
        _lib.check(self._lib.qocvl_exponentials(self._h, self._abc(abc), C.c_void_p(out.data_ptr()),
                                                self._stream()), "exponentials")
        return out

    def propagate(self, abc):
        D = 2 * self.desc.n
        out = self._new((D, D), torch.complex128)
        _lib.check(self._lib.qocvl_propagate(self._h, self._abc(abc), C.c_void_p(out.data_ptr()),
                                             self._stream()), "propagate")
        return out

    def cost(self, abc, out=None):
        out = self._new((3,)) if out is None else out
        _lib.check(self._lib.qocvl_cost(self._h, self._abc(abc), C.c_void_p(out.data_ptr()), self._stream()),
                   "cost")
        return out

    def cost_grad(self, abc, out=None):
        """-> float64 [3 + 3*N*C] = cost, infid, adiab, ga, gb, gc (device tensor, stream-ordered)."""
        out = self._new((3 + self.n_grad,)) if out is None else out
        _lib.check(self._lib.qocvl_cost_grad(self._h, self._abc(abc), C.c_void_p(out.data_ptr()),
                                             self._stream()), "cost_grad")
        return out

    def cost_grad_from_wave(self, wave, want_grad=True):
        wave = wave.to(self.device, torch.float64).contiguous()
        d = self.desc
        if tuple(wave.shape) != (d.n_slices, d.n_chan):
            raise ValueError("wave must be [n_slices, n_chan]")
        out3 = self._new((3,))
        gw = self._new((d.n_slices, d.n_chan)) if want_grad else None
        _lib.check(self._lib.qocvl_cost_grad_from_wave(
            self._h, C.c_void_p(wave.data_ptr()), C.c_void_p(out3.data_ptr()),
            C.c_void_p(gw.data_ptr()) if want_grad else None, self._stream()), "cost_grad_from_wave")
        return out3, gw

    # ---- diagnostics ----------------------------------------------------------------------
    def check(self):
        """Synchronise and raise if a slice exceeded the Taylor propagator's norm range."""
        _lib.check(self._lib.qocvl_check(self._h), "check")

    @property
    def launches(self):
        return int(self._lib.qocvl_launch_count(self._h))

    @property
    def taylor_cap(self):
        return int(self._lib.qocvl_last_taylor_degree(self._h))

    @property
    def reduced_rows(self):
        """Rows per sample of the reduced stacked system (chain_aug.cu), 0 if the last configuration does not use it."""
        return int(self._lib.qocvl_reduced_rows(self._h))

    def hbm_bytes_per_call(self, with_grad=True):
        return float(self._lib.qocvl_hbm_bytes_per_call(self._h, 1 if with_grad else 0))

    def chain_info(self):
        """Launch geometry of the chain kernel (after the first evaluation)."""
        import ctypes
        buf = (ctypes.c_int * 8)()
        _lib.check(self._lib.qocvl_chain_info(self._h, buf), "chain_info")
        keys = ("path", "rows", "lanes", "slots", "classes", "nnz", "threads", "ctas")
        info = dict(zip(keys, (int(x) for x in buf)))
        info["path"] = ("lane_per_row", "reduced_stacked", "time_parallel", "rows_strided",
                        "reduced_stacked_streamed", "samples_minor_streamed", "samples_minor_recomputed",
                        "samples_minor_time_parallel", "reduced_rows_streamed", "reduced_rows",
                        "reduced_duo")[info["path"]]
        return info

    @property
    def vectors_propagated(self):
        return int(self._lib.qocvl_vectors_propagated(self._h))

    @property
    def workspace_bytes(self):
        return int(self._lib.qocvl_workspace_bytes(self._h))

    def set_timing(self, enabled=True):
        _lib.check(self._lib.qocvl_set_timing(self._h, 1 if enabled else 0), "set_timing")

    def last_chain_ms(self):
        return float(self._lib.qocvl_last_chain_ms(self._h))

    @property
    def slice_chunks(self):
        """Concurrent slice-axis chunks of the configured path (1 = sequential sweep per sample)."""
        return int(self._lib.qocvl_slice_chunks(self._h))

    def role_degrees(self):
        """Taylor degrees of the last evaluation: int array [roles, M] for the two-lane kernels (one row per role: every
        role runs the degree its own norm bound asks for), [1, M] for paths with one degree per slice."""
        import ctypes
        import numpy as np
        m = int(self.desc.n_slices)
        buf = (ctypes.c_int * (4 * m))()
        n = int(self._lib.qocvl_role_degrees(self._h, buf))
        return np.ctypeslib.as_array(buf)[: max(n, 1) * m].reshape(max(n, 1), m).copy()

    def role_degree_means(self):
        """{'per_role': mean degree of every role (two-lane kernels; [] otherwise), 'global': mean of the one degree per slice}."""
        import ctypes
        m = int(self.desc.n_slices)
        buf = (ctypes.c_int * (4 * m))()
        n = int(self._lib.qocvl_role_degrees(self._h, buf))
        q = [sum(buf[r * m:(r + 1) * m]) / m for r in range(max(n, 1))]
        return {"per_role": q if n else [], "global": None if n else q[0]}

    def last_stage_ms(self, stage):
        """Duration of one kernel of the last timed evaluation: 0 = forward sweep, 1 = reverse sweep (< 0: not timed)."""
        return float(self._lib.qocvl_last_stage_ms(self._h, int(stage)))

    def flops_per_call(self, with_grad=True):
        """with_grad: False / 0 cost only, True / 1 executed cost + gradient, 2 algorithmic (no recomputation), 3 the
        reverse sweep's share of 2 (include/qocvl.h)."""
        return float(self._lib.qocvl_flops_per_call(self._h, int(with_grad)))
