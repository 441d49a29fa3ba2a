/* qocvl.h -- C-ABI of the B200 (sm_100a) Van Loan cost + gradient engine.
 *
 * This is the drop-in boundary for ONE path of azhutov/optimal-control-rydberg: the
 * cost / cost-and-gradient evaluation of its two PropagatorVL classes
 *   ST = code/quantum_optimal_control/state_transfer/propagator_vl.py
 *   TQ = code/quantum_optimal_control/two_qubit/propagator_vl.py
 * The reference has no FFI of its own (it is pure Python on JAX); each entry point below
 * names the reference method(s) it replaces.  All hot-path calls are stream-ordered,
 * take caller-owned DEVICE pointers, never synchronise and never throw: they return a
 * status code (0 = ok, see qocvl_strerror).  A plan is bound to one device and is not
 * thread-safe; use one plan per GPU / rank.
 *
 * Complex numbers are interleaved (re, im) doubles.  Matrices are row-major.
 */
#ifndef QOCVL_H
#define QOCVL_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qocvl_plan qocvl_plan;

enum {
  QOCVL_OK = 0,
  QOCVL_ERR_ARG = 1,        /* bad argument / shape */
  QOCVL_ERR_CUDA = 2,       /* a CUDA runtime call failed; qocvl_last_cuda_error() has the text */
  QOCVL_ERR_STATE = 3,      /* generators / cost / noise not set yet */
  QOCVL_ERR_NORM = 4,       /* |X_k| bound too large for the Taylor propagator: refine delta_t */
  QOCVL_ERR_UNSUPPORTED = 5 /* e.g. complex64 on a path that has no complex64 kernel */
};

enum { QOCVL_MODEL_STATE_TRANSFER = 0, QOCVL_MODEL_TWO_QUBIT = 1, QOCVL_MODEL_CHAIN = 2 };

/* Problem description.  Mirrors the constructor arguments that reach the hot path:
 * ST:33-56 and TQ:49-102. */
typedef struct {
  int model;        /* QOCVL_MODEL_* : selects the pulse map and auxiliary-amplitude rules */
  int n;            /* Hilbert-space dimension (5, 16, or the blockade-subspace size) */
  int n_gen;        /* J: number of Van Loan generators (11 for ST/TQ) */
  int n_slices;     /* M: time slices = n_basis_pts + 2*pad */
  int n_basis_pts;  /* nt: points of the Gaussian basis (no_of_steps) */
  int pad;          /* zero padding on each side (TQ only; 0 otherwise) */
  int n_gauss;      /* N: Gaussians per channel (input_dim) */
  int n_chan;       /* C: control channels (5 for ST/TQ, 2 for CHAIN) */
  int n_vec;        /* number of start kets the cost reads (4 for TQ, 1 for ST) */
  int adiab_index;  /* which start ket is psi of the adiabaticity term */
  double dt;        /* delta_t */
  double duration;  /* ST:46 / TQ:88 */
  double w_infid, w_adiab; /* ST:310 (0.6/0.4), TQ:574 (0.2/0.8) */
  double fid_norm;  /* divisor of |sum_v <y_v|U|x_v>|^2: 16 for TQ (TQ:555-556), 1 for ST */
  double eps;       /* TQ:433 denominator clamp (1e-12); ignored by ST */
  double scale[4];  /* model scalars: ST {Rabi_1, Rabi_2}; TQ {Rabi_i, Rabi_r}; CHAIN {..} */
} qocvl_desc;

int qocvl_plan_create(qocvl_plan** out, const qocvl_desc* desc, int device);
int qocvl_plan_destroy(qocvl_plan* plan);

/* Generators, as built by ST:132-170 / TQ:219-314 but passed as their two distinct blocks:
 * gens_u[j] = top-left n x n block (== bottom-right), gens_w[j] = top-right n x n block.
 * HOST pointers, [J][n][n] complex.  coef_const[j] (complex, HOST) is the constant
 * coefficient of a drift generator (ST:254: 1 for j=0,1; TQ:470: -i for j=0,1) and is
 * ignored for pulse-driven generators (those get their coefficient from the auxiliary
 * amplitudes, ST:229-245 / TQ:406-455). */
int qocvl_set_generators(qocvl_plan* plan, const double* gens_u, const double* gens_w,
                         const double* coef_const);

/* Real impulse response h[M] of the frequency filter TQ:375-397 (HOST pointer):
 * wave[k] = sum_t h[(k - t - pad) mod M] * reg[t].  Pass NULL for "no filter" (ST). */
int qocvl_set_filter(qocvl_plan* plan, const double* taps);

/* What the cost reads (ST:291-302, TQ:549-567): start kets x_v and target kets y_v,
 * HOST [n_vec][n] complex.  infid = 1 - |sum_v y_v^dag U x_v|^2 / fid_norm,
 * adiab = Re(1 - (U psi)^dag (W psi) / duration), psi = x_{adiab_index}. */
int qocvl_set_cost(qocvl_plan* plan, const double* starts, const double* targets);

/* Optional symmetry hint: perm[n] (HOST) is a permutation of the basis states under which every
 * generator's diagonal block is invariant, G^u_j[perm[a]][perm[b]] == G^u_j[a][b] (e.g. exchange of
 * the two identical atoms of the TQ model, TQ:201-207).  The engine VERIFIES the invariance
 * (QOCVL_ERR_ARG if it does not hold) and then propagates only one of any two start kets that the
 * permutation maps onto each other (with their target kets), counting its overlap twice.  Results
 * are unchanged; NULL clears the hint. */
int qocvl_set_symmetry(qocvl_plan* plan, const int* perm);

/* Robustness ensemble: sample s uses coefficients  c_kj * mul[s][j] + add[s][j].
 * HOST pointers: mul [S][J] real, add [S][J] complex (NULL = zeros).  n_samples = 1 with
 * mul = NULL restores the single nominal sample.  The cost and gradient are ensemble means. */
int qocvl_set_noise(qocvl_plan* plan, int n_samples, const double* mul, const double* add);

/* Device bytes held by the plan after the last set_* / evaluation call. */
size_t qocvl_workspace_bytes(const qocvl_plan* plan);

/* ---- hot path (DEVICE pointers, stream-ordered; stream is a cudaStream_t) ------------ */

/* ST:215-227 / TQ:399-404 return_physical_amplitudes: abc = [3][N][C] (a, b, c) -> wave [M][C]. */
int qocvl_physical_amplitudes(qocvl_plan* plan, const double* abc, double* wave, void* stream);

/* Reverse of the above (what jax.grad builds for notebook 05 cells 8-12, which differentiates
 * THROUGH return_physical_amplitudes): gwave [M][C] -> grads [3][N][C] = d/d(a,b,c) of
 * sum_kc gwave[k][c] * wave[k][c].  Recomputes the forward pulse map internally. */
int qocvl_physical_amplitudes_vjp(qocvl_plan* plan, const double* abc, const double* gwave,
                                  double* grads, void* stream);

/* ST:229-245 / TQ:406-455 return_auxiliary_amplitudes: -> aux [M][J-2] complex. */
int qocvl_auxiliary_amplitudes(qocvl_plan* plan, const double* abc, double* aux, void* stream);

/* ST:247-256 / TQ:460-484 exponentials: -> [M][2n][2n] complex (dense, nominal sample).  2n <= 32: one CTA per slice in
 * shared memory; larger n: block-triangular (u, w) pairs multiplied by a batched FP64 tensor-core (DMMA) complex GEMM --
 * that path synchronises the stream (it sizes the scaling on the host) and is not graph-capturable. */
int qocvl_exponentials(qocvl_plan* plan, const double* abc, double* out, void* stream);

/* ST:258-283 / TQ:489-535 propagate: -> [2n][2n] complex; same product order as the reference. */
int qocvl_propagate(qocvl_plan* plan, const double* abc, double* out, void* stream);

/* ST:285-310 / TQ:540-574 metrics + target: out3 = {cost, infidelity, adiabaticity}
 * (ensemble means when a noise ensemble is set). */
int qocvl_cost(qocvl_plan* plan, const double* abc, double* out3, void* stream);

/* jax.value_and_grad(target, argnums=(0,1,2)) (TQ:592, notebook 01 cell 9):
 * out = {cost, infid, adiab, ga[N][C], gb[N][C], gc[N][C]}  (3 + 3*N*C doubles).
 * With sample sharding (qocvl_set_shard) the values are this rank's partial SUMS already
 * divided by the GLOBAL sample count, so one all-reduce(SUM) over ranks completes them. */
int qocvl_cost_grad(qocvl_plan* plan, const double* abc, double* out, void* stream);

/* Same two evaluations starting from given physical amplitudes wave [M][C] (device): the
 * pulse map is skipped; gwave (may be NULL) receives d cost / d wave [M][C].  Used for
 * stage-wise parity tests and by callers that own their own pulse parametrisation. */
int qocvl_cost_grad_from_wave(qocvl_plan* plan, const double* wave, double* out3,
                              double* gwave, void* stream);

/* Multi-GPU: tell the plan that its samples are a shard of a larger ensemble, so means
 * are normalised by global_samples. */
int qocvl_set_shard(qocvl_plan* plan, int global_samples);

/* Explicit per-plan switches; the library reads NO environment variables.  value < 0 restores the automatic choice.
 * Names: "aug" (0 = do not use the reduced stacked system), "duo" (0 = lane-group kernel of chain_aug.cu instead
 * of the two-lanes-per-sector kernels of chain_duo.cu), "duo_wpc" (warps per CTA of those kernels), "store" (lane-group
 * kernel: 0 = recompute the Taylor terms instead of streaming them through HBM), "spb" (samples per CTA), "nofold"
 * (1 = no orbit folding), "force_big" (1 = rows-strided kernel for any n), "qcap" (Taylor-degree cap), "tp" (0 = no
 * time-parallel single-sample path), "tp_chunk" (slices per chunk), "tlocal", "wide" (0 = no samples-minor kernel),
 * "wide_store", "wide_spc", "wide_threads", "strong_chunks" (slice-axis chunks of the time-parallel evaluation of an
 * ensemble on the two-lane kernels; automatic: 1 above ~1500 samples, up to 16 below), "duo_nomirror" (1 = sweep every
 * basis column of a chunk propagator), "duo_cols" (1 = three basis columns per work item of the chunk-propagator sweep instead of one),
 * "debug_sync" (1 = synchronise after every launch of the two-lane path and name the kernel that failed),
 * "duo_nolean" (1 = qocvl_cost_grad accumulates the gradient of every matrix entry,
 * also of those only a pulse channel that the model's pulse map discards would consume),
 * "duo_roleq" (Taylor-degree rule of the two-lane kernels: 0 = one degree per slice from the whole system's 1-norm
 * bound, 1 = per role from the role's 1-norm bound, 2 = min with the role's 2-norm bound, 3 [default] = the same for
 * the generator shifted by the role's phase, see qocvl_role_degrees), "wide_norms" / "wide_shift" (0 = the samples-minor
 * kernel without the 2-norm degree kernel / without the phase shift), "taylor_tol_exp" (e: Taylor remainder
 * tolerance 2^-e instead of 2^-55 for complex128 / 2^-26 for complex64), "graph" (0 = qocvl_cost / qocvl_cost_grad
 * always launch kernel by kernel; default: a caller that repeats the call with the same buffers gets the launch
 * sequence replayed from a cached CUDA graph, bit-identical results).
 * They exist so that every kernel path can be cross-checked against the others on the same plan API.
 * QOCVL_ERR_ARG for an unknown name. */
int qocvl_set_option(qocvl_plan* plan, const char* name, int value);

/* Arithmetic type of the ensemble kernel (north_star: "batched complex128 (optional complex64)"; the reference
 * itself runs complex128 only, jax_enable_x64 at TQ:1-12).  QOCVL_C64 runs the reduced stacked system
 * (qocvl_reduced_rows) with complex64 state, Taylor terms and matrix entries, Taylor remainder 2^-26; the pulse
 * map, the coefficients, the per-sample cost and the reduction of the gradient stay in fp64.  Stated bounds
 * against the complex128 result: cost 5e-5 absolute, gradient 1e-2 of its largest entry.  Large Hilbert spaces
 * (blockade chains) run the complex64 instantiation of the samples-minor kernel (vectors, Taylor terms, HBM stream
 * in fp32).  Only the cross-check fallback kernels have no complex64 form: an evaluation forced onto them returns
 * QOCVL_ERR_UNSUPPORTED in this mode (no silent fp64 run). */
enum { QOCVL_C128 = 0, QOCVL_C64 = 1 };
int qocvl_set_precision(qocvl_plan* plan, int dtype);
int qocvl_precision(const qocvl_plan* plan);

/* Diagnostics */
/* Synchronises; returns QOCVL_ERR_NORM if any evaluation since the last check saw a slice whose
 * norm bound exceeded the Taylor propagator's range (or NaN pulses), QOCVL_OK otherwise. */
int qocvl_check(qocvl_plan* plan);
int qocvl_last_taylor_degree(const qocvl_plan* plan); /* cap on the per-slice Taylor degree for this plan */
/* Vectors the chain kernel propagates per sample after dropping inert start kets and merging
 * symmetric ones (moving start kets + the W psi vector); 0 before the first evaluation. */
int qocvl_vectors_propagated(const qocvl_plan* plan);
/* Rows per sample of the reduced stacked system the ensemble kernel runs on (every propagated vector restricted
 * to the basis states it can reach, symmetric sectors folded onto orbit representatives; CZ: 12 instead of
 * 3 x 16), or 0 when the evaluation does not use it (single-sample time-parallel path, large Hilbert spaces,
 * QOCVL_AUG=0).  Results are identical either way: only structural zeros are skipped. */
int qocvl_reduced_rows(const qocvl_plan* plan);
long long qocvl_launch_count(const qocvl_plan* plan); /* kernels launched so far by this plan */
/* FP64 flop model of the last evaluation (useful complex multiply-adds on the non-zeros of the propagated system, 8 flops
 * each).  with_grad: 0 = cost only; 1 = cost + gradient as EXECUTED (including the Taylor powers a checkpointing kernel
 * recomputes in its reverse sweep); 2 = cost + gradient ALGORITHMIC (forward q + adjoint q + Xbar q per slice: what
 * SURVEY 8(d) prices, no recomputation); 3 = the reverse sweep's share of 2 (adjoint q + Xbar q); 4 = 2 with Xbar counted
 * on every entry (modes 1-3 do not count the Xbar updates that qocvl_cost_grad skips because only a pulse channel the
 * model's pulse map discards would consume them). */
double qocvl_flops_per_call(const qocvl_plan* plan, int with_grad);
/* Algorithmic HBM bytes of one chain evaluation: checkpoints written then read once, plus the per-warp partial
 * gradient data written by the reverse sweep and read once by the reduction (0 before the first evaluation). */
double qocvl_hbm_bytes_per_call(const qocvl_plan* plan, int with_grad);
/* Launch geometry after configuration, out[8] = {path: 0 lane-per-row / 1 reduced stacked system / 2 time-parallel
 * single sample / 3 rows-strided big-n / 4 reduced stacked system with the Taylor terms streamed through HBM /
 * 5 samples-minor big-n ensemble kernel with streamed Taylor terms / 6 the same with recomputed terms / 7 its
 * time-parallel single-sample form / 8, 9 retired / 10 reduced stacked system on the two-lanes-per-sector kernels
 * (chain_duo.cu; lanes = lanes per sample, slots = widest entry list of a lane);
 * rows per sample; lanes per sample (paths 5-6: samples per CTA); off-diagonal slots per row (5-6: generator blocks
 * with off-diagonal entries); multiplier classes per entry (5-6: generator blocks with diagonal entries);
 * non-zeros of the propagated system; threads per CTA; CTAs}. */
int qocvl_chain_info(const qocvl_plan* plan, int* out8);
/* Kernel timing for the roofline report: when enabled, CUDA events bracket the chain kernel on
 * the launch stream; qocvl_last_chain_ms synchronises on the end event and returns the duration
 * of the most recent chain-kernel launch in milliseconds (< 0 if none was timed). */
int qocvl_set_timing(qocvl_plan* plan, int enabled);
double qocvl_last_chain_ms(qocvl_plan* plan);
/* Number of concurrent slice-axis chunks of the configured path (1 = every sample is swept sequentially; > 1 = the
 * time-parallel evaluation: chunk propagators -> combine -> per-chunk sweeps).  0 before the first evaluation. */
int qocvl_slice_chunks(const qocvl_plan* plan);
/* Taylor degrees of the last evaluation, per role (independent diagonal block of the propagated system) of the two-lane
 * ensemble kernels, which pick the degree per role and slice from the role's own norm bound (option "duo_roleq": 2 =
 * min(1-norm, 2-norm) bound [default], 1 = 1-norm bound, 0 = one degree per slice for the whole system):
 * out[r * n_slices + k] for r < return value (the number of roles, <= 4).  Returns 0 when the configured path has one
 * degree per slice for the whole system (then out[k] holds it).  `out` has room for 4 * n_slices ints.  Synchronises. */
int qocvl_role_degrees(const qocvl_plan* plan, int* out);
/* The same for one stage of the most recent timed evaluation: stage 0 = forward-sweep kernel, 1 = reverse-sweep kernel
 * (ensemble kernels of chain_duo.cu; < 0 when the path that ran has no such stage or nothing was timed). */
double qocvl_last_stage_ms(qocvl_plan* plan, int stage);
/* Pure-DFMA microbenchmark (no memory traffic): sustained FP64 FMA throughput of `device` in
 * TFLOP/s, best of `repeats` launches.  This is the FP64 roofline denominator (the driver's
 * MEASURED_PEAKS.json only has HBM and bf16).  Returns < 0 on error. */
double qocvl_measure_fp64_peak(int device, int repeats);
/* Same loop with three REGISTER operands per FMA (a = fma(a, b, c)), which is what real kernels issue:
 * the register file of the B200 SM sustains ~25 TFLOP/s of these, 2/3 of the issue peak above. */
double qocvl_measure_fp64_peak_reg(int device, int repeats);
/* ---- master-equation validation (the step after the hot path) -------------------------------------------
 * Replaces QuTiP's mesolve as the reference uses it in two_qubit/qutip_simulation.py:57-103 and
 * two_qubit/qutip_five_level.py:88-142: propagates vectorised density matrices under
 *   d rho/dt = (sum_j coef[traj][step][j] * L_j) rho,  L_j given as complex values vals[J][nnz] on one union
 * CSR pattern (rowptr[n+1], col[nnz]), with the exponential midpoint rule rho <- exp(L dt) rho per step (Taylor
 * series with remainder < 2^-55, as the chain kernels).  One CTA per trajectory.  expect_out receives
 * sum_i obs[o][i] * rho[i] at every step boundary, [n_traj][n_steps+1][n_obs] complex; rho_out the final vectors.
 * ALL pointers are HOST pointers; the call allocates, runs and synchronises (one-shot validation, not stream-ordered).
 * QOCVL_ERR_NORM: a step has |L dt|_1 > 4, sub-step finer.  QOCVL_ERR_UNSUPPORTED: pattern does not fit shared memory. */
int qocvl_lindblad_propagate(int device, int n, int nnz, int n_gen, int n_steps, int n_traj, int n_obs,
                             const int* rowptr, const int* col, const double* vals, const double* coef,
                             const double* rho0, const double* obs, double dt, double* expect_out, double* rho_out);
const char* qocvl_strerror(int status);
const char* qocvl_last_cuda_error(void);
int qocvl_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QOCVL_H */
