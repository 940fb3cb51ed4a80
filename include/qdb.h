/*
 * qdb.h -- C-ABI of libqdb.so, the B200 (sm_100a) time-evolution hot path.
 *
 * The reference (qiskit-dynamics 0.6.0) is pure Python and has no FFI of its own; the seams
 * this library sits behind are its two string-selected factories and the model call protocol
 * (SURVEY.md section 8(b)).  Every entry point below names the reference code it replaces
 * (paths relative to /root/reference/qiskit_dynamics/).  The Python binding a reference
 * maintainer would add is shown in INTEGRATION.md (ctypes, no torch types cross this boundary).
 *
 * Conventions
 *   - All pointers are DEVICE pointers owned by the caller unless a parameter says "host".
 *   - complex128 = interleaved (re, im) doubles ("qdb_c128"); matrices are row-major.
 *   - A state batch y is (n, B): row i = basis component, B columns contiguous, leading
 *     dimension ldy >= B (in complex elements).  This is the reference's own layout
 *     (models/operator_collections.py:124-134: `G @ y`, batch = columns).
 *   - Frame phases: mu[a] (real, length n) are the frame frequencies of row a.  The kernels use
 *     p_a(t) = exp(-i mu_a t); RHS = conj(p) .* (G (p .* y)); generator = G .* outer(conj p, p)
 *     (models/rotating_frame.py:255,350-353; vectorised states: :568-577, mu_{i+k n} = lam_i - lam_k).
 *     mu == NULL means "no rotating frame".
 *   - Functions enqueue work on `stream` (a cudaStream_t passed as void*), never synchronise
 *     the device, never allocate persistent memory and never throw.  Scratch memory is passed
 *     in by the caller (size from qdb_workspace_bytes).
 *   - Return value: 0 = OK, negative = invalid argument (QDB_E_*), positive = cudaError_t.
 *     qdb_last_error_string() gives the text of the last failure on the calling thread.
 */
#ifndef QDB_H
#define QDB_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } qdb_c128;

#define QDB_OK 0
#define QDB_E_ARG (-1)        /* bad dimension / NULL pointer / unsupported flag */
#define QDB_E_WORKSPACE (-2)  /* workspace too small */
#define QDB_E_UNSUPPORTED (-3)/* shape outside what this build supports */

/* layouts of an (n x n) operator in device memory */
#define QDB_LAYOUT_ROWMAJOR 0 /* [n][n] */
#define QDB_LAYOUT_PACKED 1   /* DMMA A-fragment order, zero padded to npad = 8*ceil(n/8) rows and
                                 kpad = 16*ceil(n/16) columns:
                                 element (r,c) at ((r/8)*(kpad/4) + c/4)*32 + (r%8)*4 + c%4 */

#define QDB_LAYOUT_PACKED3M 2 /* generator tables only: a PACKED complex plane (npad*kpad elements) followed by a
                                 plane of npad*kpad doubles holding re + im of each element -- 24 bytes per
                                 element.  The on-chip RK4 kernel multiplies complex tiles with three real
                                 tensor-core products (re*re, im*im, (re+im)*(re+im)) and reads the sums from
                                 this plane instead of forming them in its inner loop. */

/* workspace kinds for qdb_workspace_bytes */
#define QDB_WS_RHS 0
#define QDB_WS_RK4 1
#define QDB_WS_EXPM 2
#define QDB_WS_MAGNUS 3
#define QDB_WS_PROP 4

const char* qdb_last_error_string(void);
int qdb_version(void);

/* n rounded up to the DMMA tile (8) and the element count of one packed operator (npad*kpad). */
int qdb_npad(int n);
size_t qdb_packed_elems(int n);

/* Bytes of scratch the steppers need.  S = number of steps handled per call (the RK4 stepper
 * accepts any workspace >= the S=1 size and chunks the step loop to fit). */
size_t qdb_workspace_bytes(int kind, int n, int K, int B, int S);

/* Re-order `count` row-major (n x n) operators into QDB_LAYOUT_PACKED (one-time, at model
 * construction).  Replaces nothing in the reference; it is the HBM layout of
 * OperatorCollection._operators / _static_operator (models/operator_collections.py:73-81). */
int qdb_pack_operators(int n, int count, const qdb_c128* src, qdb_c128* dst, void* stream);

/* a1 + a5: generator table.  For each of the T times
 *     out[t] = scale * (stat + sum_j coeff[t][j] * ops[j]) .* outer(conj p(t), p(t))
 * ops/stat/out all in `layout` (QDB_LAYOUT_PACKED3M: ops/stat PACKED, out PACKED3M).
 * coeff is [T][K] real, or [T][K] complex when coeff_complex != 0.
 * stat or ops may be NULL (K = 0), not both.  times may be NULL when mu is NULL.
 * Replaces OperatorCollection.evaluate (models/operator_collections.py:101-122 ->
 * arraylias/register_functions/linear_combo.py:30-32) and RotatingFrame.operator_into_frame in
 * the frame basis (models/rotating_frame.py:350-353), i.e. GeneratorModel.evaluate
 * (models/generator_model.py:256-279) and LindbladModel.evaluate (models/lindblad_model.py:436-475). */
int qdb_generator_c128(int n, int K, int T, int layout,
                       const qdb_c128* ops, const qdb_c128* stat,
                       const double* coeff, int coeff_complex,
                       const double* mu, const double* times, double scale,
                       qdb_c128* out, void* stream);

/* a3 on its own:  y_out[a][b] = q_a * y_in[a][b],  q = exp(-i mu t) (conj_phase = 0: "out of frame")
 * or exp(+i mu t) (conj_phase = 1: "into frame"), in the frame basis.  May run in place.
 * Replaces RotatingFrame.state_into_frame / state_out_of_frame (models/rotating_frame.py:225-284). */
int qdb_frame_apply_c128(int n, int B, const double* mu, double t, int conj_phase,
                         const qdb_c128* y_in, qdb_c128* y_out, int ldy, void* stream);

/* General complex GEMM with the fused pro/epilogues the path needs:
 *     C[m][b] = beta * C[m][b] + alpha * colscale[b] * post[m] * sum_k A[m][k] * pre[k] * Bm[k][b]
 * colscale (real, length N), pre (length Kd) and post (length M) may each be NULL (= 1).
 * Replaces np.matmul in _matmul (models/operator_collections.py:31-32,134), the frame-basis
 * changes U^dag y0 / U y (solvers/solver_functions.py:396-403,436-448) and the products inside
 * scipy.linalg.expm (solvers/fixed_step_solvers.py:22,104).
 * Products that fill half the SMs with 128 x 32 tiles and have Kd >= 384 run on the int8 tensor cores (tcgen05.mma
 * kind::i8: six signed byte slices per operand against a power-of-two scale per row of A and per column of Bm, exact int32
 * slice products, normwise error 2^-48 per operand -- measured 9e-14 against cuBLAS ZGEMM; 1.95x its speed at
 * 729 x 4096 x 729, Kd <= 4096); smaller ones on the fp64 DMMA kernels.  Environment: QDB_ZGEMM_INT8=0 keeps every product on the
 * DMMA kernels, QDB_ZGEMM_SLICES=5 trades accuracy (2^-40) for ~10 % speed.  The emulated path takes its scratch
 * from the device's stream-ordered memory pool (cudaMallocAsync on `stream`). */
int qdb_zgemm_c128(int M, int N, int Kd,
                   const qdb_c128* A, int lda, const qdb_c128* Bm, int ldb,
                   qdb_c128* C, int ldc, qdb_c128 alpha, qdb_c128 beta,
                   const double* colscale, const qdb_c128* pre, const qdb_c128* post,
                   void* stream);

/* a1 + a2 + a3 fused: one RHS evaluation  y_out = conj(p) .* ((stat + sum_j c_j ops_j)(p .* y_in)).
 * coeff_per_col == 0: coeff is [K] (shared by all columns);
 * coeff_per_col == 1: coeff is [K][ldc] real, column b uses coeff[j][b] ("sweep mode" -- the
 *                     reference reaches this only through its sequential list-of-simulations loop,
 *                     solvers/solver_classes.py:556-590).
 * ops/stat row-major.  Replaces GeneratorModel.evaluate_rhs (models/generator_model.py:281-316),
 * OperatorCollection.evaluate_rhs (models/operator_collections.py:124-134) and
 * RotatingFrame.state_into_frame/state_out_of_frame (models/rotating_frame.py:225-284). */
int qdb_rhs_c128(int n, int K, int B,
                 const qdb_c128* ops, const qdb_c128* stat,
                 const double* coeff, int coeff_per_col, int ldc,
                 const double* mu, double t,
                 const qdb_c128* y_in, qdb_c128* y_out, int ldy,
                 void* workspace, size_t ws_bytes, void* stream);

/* a7 + a8 inner loop: S fixed RK4 steps of size h, state resident on chip.
 *   times  : HOST pointer, [2S+1] stage times t_0, t_0+h/2, t_1, ... t_S built with the
 *            reference's accumulation (solvers/fixed_step_solvers.py:448-454, :64-66)
 *   coeff  : device, signal table.  sig_mode 0: [2S+1][K] shared; sig_mode 1: [2S+1][K][ldc]
 *            per column (sweep).
 *   ops_packed / stat_packed : QDB_LAYOUT_PACKED (sig_mode 1 and the on-chip path), plus the
 *            row-major copies ops_rm / stat_rm used when n is too large for the on-chip path.
 *   workspace : qdb_workspace_bytes(QDB_WS_RK4, n, K, B, S) covers either mode.  sig_mode 0 chunks the
 *            step loop when the generator table does not fit a smaller workspace.  sig_mode 1 with
 *            3 <= K <= 16 forms the per-column generator on the tensor pipe (rk4_sweepf_kernel) when the
 *            workspace has room for its operator copy behind the stage times, and runs the
 *            operator-pass kernels otherwise (and for other K).
 * Replaces RK4_solver.take_step + the fixed_step_solver_template loop
 * (solvers/fixed_step_solvers.py:43-77,441-454) applied to GeneratorModel.evaluate_rhs. */
int qdb_rk4_steps_c128(int n, int K, int B, int S,
                       const qdb_c128* ops_rm, const qdb_c128* stat_rm,
                       const qdb_c128* ops_packed, const qdb_c128* stat_packed,
                       const double* coeff, int sig_mode, int ldc,
                       const double* mu, const double* times_host, double h,
                       qdb_c128* y, int ldy,
                       void* workspace, size_t ws_bytes, void* stream);

/* The on-chip RK4 kernel alone: S steps from a prebuilt generator table of 2S+1 entries G_frame(t) at the
 * stage times, as produced by qdb_generator_c128 in `table_layout` (QDB_LAYOUT_PACKED: 4-product kernels;
 * QDB_LAYOUT_PACKED3M: the 3-product kernel, available when qdb_rk4_table_layout says so).  n <= 256.
 * This is the dominant launch of qdb_rk4_steps_c128 in shared-signal mode; exported so that it can be
 * timed/profiled alone. */
int qdb_rk4_table_steps_c128(int n, int B, int S, const qdb_c128* gen_table, int table_layout, double h,
                             qdb_c128* y, int ldy, void* stream);

/* The same S steps with the fp64 contraction emulated on the int8 tensor cores (tcgen05.mma kind::i8, accumulators and
 * generator slices in TMEM): every real operand is split error-free into 5 signed byte slices against a per-row /
 * per-column power-of-two scale, slice products are exact in int32, the 15 slice pairs above the operands' truncation
 * error are kept (normwise error 2^-40 per operand and RHS evaluation; measured 3e-12 against the DMMA kernel after
 * 1000 steps of the headline problem).  n = 65..128 (rows and the contraction index are padded to 128).
 * qdb_rk4_int8_preferred: 1 when qdb_rk4_steps_c128 takes this kernel for a shared-signal solve of this shape (where it
 *   measured faster than the DMMA kernels: n >= 121 from B >= 960, n >= 96 from B >= 1536, n >= 65 above 16 columns per
 *   SM; the environment variable QDB_RK4_INT8=0 keeps every batch on the fp64 DMMA kernels).
 * qdb_rk4_ozaki_slice_c128: T table entries (QDB_LAYOUT_ROWMAJOR or QDB_LAYOUT_PACKED) -> int8 slice planes + row
 *   exponents in `workspace` (qdb_rk4_ozaki_workspace_bytes(S) for T = 2S+1 entries).
 * qdb_rk4_ozaki_steps_c128: gen_table_rowmajor = 2S+1 entries in QDB_LAYOUT_ROWMAJOR, sliced into `workspace` first; or
 *   NULL: `workspace` already holds the planes of these 2S+1 entries (the stepper alone, for timing / profiling). */
size_t qdb_rk4_ozaki_workspace_bytes(int S);
int qdb_rk4_int8_preferred(int n, int B);
int qdb_rk4_ozaki_slice_c128(int n, int T, const qdb_c128* gen_table, int table_layout, void* workspace, size_t ws_bytes,
                             void* stream);
int qdb_rk4_ozaki_steps_c128(int n, int B, int S, const qdb_c128* gen_table_rowmajor, double h,
                             qdb_c128* y, int ldy, void* workspace, size_t ws_bytes, void* stream);

/* Table layout qdb_rk4_steps_c128 uses for this shape (QDB_LAYOUT_PACKED or QDB_LAYOUT_PACKED3M), and the
 * bytes of one table entry in a layout. */
int qdb_rk4_table_layout(int n, int B);
size_t qdb_table_entry_bytes(int n, int layout);

/* Tiling the on-chip RK4 kernels pick for a shape (diagnostic: reported by bench.py, asserted by the
 * tests).  sweep_K = 0 asks about the shared-signal solve (the better of the 3- and 4-product kernels),
 * sweep_K < 0 about the 4-product shared-signal kernel (what a QDB_LAYOUT_PACKED table runs on), > 0 about
 * the per-column kernel.
 * out[0..7] = {warps along rows, warps along columns, row tiles per warp, own column tiles per warp,
 * split (1 = 2-CTA clusters sharing one column octet through DSMEM), CTAs, threads per CTA,
 * dynamic shared memory bytes, 3M (1 = three-product complex tiles; 2 = formed-generator sweep kernel)}.  Returns QDB_E_UNSUPPORTED when n is
 * outside the on-chip path. */
int qdb_rk4_tiling(int n, int B, int sweep_K, int* out /* host, 9 ints */);

/* fp64 tensor-pipe (DMMA m8n8k4) issue-rate probe: launches register-resident DMMA chains on every
 * SM; *flops_out (host) receives the flop count of the launch.  Timed by the caller with CUDA
 * events, it gives the live roofline denominator for the fp64 kernels. */
int qdb_dmma_probe(double* sink, int iters, double* flops_out, void* stream);

/* a9: S exponential (Magnus order 1) steps  y <- expm(h * G_frame(t_s + h/2)) y  with a
 * scaling-and-squaring Taylor propagator built from qdb_zgemm_c128 products.
 *   times_mid_host : HOST [S] midpoints t_s + h/2;  coeff : device [S][K] at those midpoints
 *   squarings_host : HOST [S] number of squarings per step (chosen by the caller from a norm
 *                    bound so that no device->host sync is needed inside the loop)
 *   workspace      : qdb_workspace_bytes(QDB_WS_EXPM, n, K, B, S).  With the S = 1 size the steps run one at a
 *                    time; with more, the propagators of a chunk of steps are built side by side (one generator
 *                    launch + one batched Taylor exponential, all steps of the chunk sharing the largest number
 *                    of squarings) and then applied in order -- so the result depends on the workspace size at the
 *                    rounding level (extra squarings on the weaker steps of a chunk).
 * Replaces get_exponential_take_step(magnus_order=1) + scipy.linalg.expm
 * (solvers/fixed_step_solvers.py:343-346,400-401,104). */
int qdb_expm_steps_c128(int n, int K, int B, int S,
                        const qdb_c128* ops_rm, const qdb_c128* stat_rm,
                        const double* coeff, const double* mu,
                        const double* times_mid_host, const int* squarings_host, double h,
                        qdb_c128* y, int ldy,
                        void* workspace, size_t ws_bytes, void* stream);

/* a9 at Magnus orders 1, 2 and 3: S steps  y <- expm(Omega_s) y, where Omega_s is the truncated Magnus
 * exponent of the step built from the generator at the order's Gauss-Legendre nodes:
 *   order 1: Omega = h G(t + h/2)
 *   order 2: Omega = h (g1 + g2)/2 + (sqrt(3)/12) h^2 [g2, g1],            nodes t + (1/2 -+ sqrt(3)/6) h
 *   order 3: a1 = h g2, a2 = (sqrt(15)/3) h (g3 - g1), a3 = (10/3) h (g3 - 2 g2 + g1),
 *            Omega = a1 + a3/12 + [-20 a1 - a3 + [a1,a2], a2 + [2 a3 + [a1,a2], a1]/60]/240,
 *                                                                          nodes t + (1/2 -+ sqrt(15)/10) h, t + h/2
 * (generator_kernel at the nodes, commutators as pairs of DMMA GEMMs, then the same Taylor
 * scaling-and-squaring propagator and apply GEMM as qdb_expm_steps_c128).
 *   times_host     : HOST [S][magnus_order] node times;  coeff : device [S][magnus_order][K] signal values there
 *   squarings_host : HOST [S], from a norm bound on Omega_s
 *   workspace      : qdb_workspace_bytes(QDB_WS_MAGNUS, n, K, B, S)
 * Replaces get_exponential_take_step(magnus_order) + scipy.linalg.expm inside scipy_expm_solver
 * (solvers/fixed_step_solvers.py:80-108, 327-401). */
int qdb_magnus_steps_c128(int n, int K, int B, int S, int magnus_order,
                          const qdb_c128* ops_rm, const qdb_c128* stat_rm,
                          const double* coeff, const double* mu,
                          const double* times_host, const int* squarings_host, double h,
                          qdb_c128* y, int ldy,
                          void* workspace, size_t ws_bytes, void* stream);

/* The Magnus exponent alone: out = scale * Omega(h) from g = [magnus_order][n][n] row-major generators at the
 * nodes (for generators that are arbitrary host callables: the callable protocol of scipy_expm_solver,
 * solvers/fixed_step_solvers.py:80-108).  workspace: 0 / n^2 / 7 n^2 complex numbers for order 1 / 2 / 3. */
int qdb_magnus_terms_c128(int n, int magnus_order, const qdb_c128* g, double h, double scale, qdb_c128* out,
                          void* workspace, size_t ws_bytes, void* stream);

/* Time-parallel LMDE stepping: the product  P_total = P_{S-1} ... P_1 P_0  of the S one-step propagators of an
 * interval, built side by side (the step index is the batch dimension of the DMMA GEMM launches) and multiplied
 * pairwise; the caller applies it to the state batch with one qdb_zgemm_c128.
 *   kind 0      : RK4 propagator  1 + (h/6)(k1 + 2 k2 + 2 k3 + k4)  (generator at t, t + h/2, t + h:
 *                 times_host [S][3], coeff [S][3][K])
 *   kind 1,2,3  : expm of the Magnus exponent of that order (times_host [S][kind], coeff [S][kind][K],
 *                 squarings_host [S]; see qdb_magnus_steps_c128)
 *   workspace   : qdb_workspace_bytes(QDB_WS_PROP, n, K, 0, S) holds all S steps at once; with less (at least the
 *                 S = 1 size) the steps are processed in chunks.
 * Replaces jax_RK4_parallel_solver / jax_expm_parallel_solver and their vmap + associative_scan template
 * (solvers/fixed_step_solvers.py:206-244, 279-311, 524-613), which the reference offers on JAX only. */
int qdb_step_propagators_c128(int n, int K, int S, int kind,
                              const qdb_c128* ops_rm, const qdb_c128* stat_rm,
                              const double* coeff, const double* mu,
                              const double* times_host, const int* squarings_host, double h,
                              qdb_c128* P_total,
                              void* workspace, size_t ws_bytes, void* stream);

/* Matrix exponential of one (n x n) row-major matrix with `squarings` halvings (building block
 * of qdb_expm_steps_c128, exported for parity tests against scipy.linalg.expm). */
int qdb_expm_c128(int n, const qdb_c128* A, int squarings, qdb_c128* out,
                  void* workspace, size_t ws_bytes, void* stream);

/* f3 (and a6 on the device): coefficient table of K signal channels at T times.
 *     out[t][j]    (B == 0: shared signals)   or   out[t][j][b]  (B > 0: one column per simulation)
 *         = sum over terms i with chan[i] == j of  Re[ scale[i][b] * f_i(t) * exp(i (2 pi freq[i] t + phase[i])) ]
 * f_i is piecewise constant: samples[samp_off[i] + b * samp_col_stride + idx], idx = (t - t0[i]) // dt[i] with
 * NumPy's float floor division, zero outside [0, samp_len[i]) -- or constant (samp_len[i] == -1: the single
 * sample at samp_off[i]).  samp_col_stride = 0 shares the samples between columns; scale may be NULL (= 1).
 * chan, samp_off, samp_len have nterms entries; dt, t0, freq, phase have nterms entries, or [nterms][B] when
 * params_per_col != 0 (frequency / phase / timing sweeps).  Everything lives on the device; times == NULL with
 * T == 1 evaluates the single time t_scalar (one RHS call: no host-to-device copy at all).
 * Replaces SignalList.__call__ -> SignalSum.complex_value -> DiscreteSignal.envelope
 * (signals/signals.py:801-803, 574-577, 296-311, 148-155) evaluated on the stage-time grid. */
int qdb_signal_table_f64(int T, int K, int B, int nterms,
                         const int* chan, const long long* samp_off, const int* samp_len,
                         const double* dt, const double* t0, const double* freq, const double* phase,
                         int params_per_col,
                         const qdb_c128* samples, long long samp_col_stride, const qdb_c128* scale,
                         const double* times, double t_scalar, double* out, void* stream);

/* f4: memory-slot outcome probabilities of a batch of final states (already in the measurement basis):
 *     out[o][b] = sum_{i : outcome_of[i] == o} |y[i][b]|^2,  divided by sum_i |y[i][b]|^2 when normalize != 0.
 * outcome_of (device, n ints in [0, n_out)) maps every basis state to its outcome bin; out is [n_out][B].
 * Replaces Statevector.probabilities_dict + _get_memory_slot_probabilities + the normalisation of
 * _get_experiment_result (backend/dynamics_backend.py:846-866, backend/backend_utils.py:106-147). */
int qdb_outcome_probabilities_f64(int n, int B, int n_out, const qdb_c128* y, int ldy, const int* outcome_of,
                                  int normalize, double* out, void* stream);

/* f2: the non-vectorised Lindblad equation on a batch of density matrices rho (B, n, n) row-major -- batch = LEADING axis,
 * as in the reference -- evaluated as
 *     rhs(X) = M1 X + X M2 + sum_j g_j L_j X L_j^dag,   M1 = A + B,  M2 = A - B,  B = -i H(t),  A = -1/2 sum_j g_j L_j^dag L_j
 * (g_j = 1 for static dissipators), with the frame phases around it: X = rho .* (p_i conj p_k), out = rhs(X) .* (conj p_i p_k),
 * p = exp(-i mu t); mu == NULL: no frame.  O(n^3) per density matrix (the vectorised form is O(n^4)).
 *   m1_packed / m2t_packed : M1(t) and the TRANSPOSE of M2(t) in QDB_LAYOUT_PACKED -- entries of generator tables the caller
 *                  builds with qdb_generator_c128 from [-i H_j ; -1/2 L_j^dag L_j] (M1) and their transposes with +i H_j (M2^T)
 *   diss_packed  : [J] dissipators L_j (static and time dependent ones alike), QDB_LAYOUT_PACKED
 *   gamma        : device [J] coefficients g_j(t), or NULL (all 1)
 * One CTA per density matrix, operands in shared memory, operators streamed from L2; n <= 32 (qdb_lindblad_supported).
 * Replaces LindbladCollection.evaluate_rhs (models/operator_collections.py:451-567, batch broadcast :506-510) inside
 * LindbladModel.evaluate_rhs (models/lindblad_model.py:477-538) with RotatingFrame.operator_out_of_frame / operator_into_frame
 * (models/rotating_frame.py:286-370). */
int qdb_lindblad_supported(int n);
int qdb_lindblad_rhs_c128(int n, int J, int B,
                          const qdb_c128* m1_packed, const qdb_c128* m2t_packed, const qdb_c128* diss_packed,
                          const double* gamma, const double* mu, double t,
                          const qdb_c128* rho_in, qdb_c128* rho_out, void* stream);

/* f2 + a7 + a8: S fixed RK4 steps of the same equation with every density matrix resident on chip for the whole launch.
 *   m1_table / m2t_table : [2S+1] entries at the stage times t_0, t_0 + h/2, t_1, ... (QDB_LAYOUT_PACKED)
 *   gamma_table  : device [2S+1][J], or NULL;   times_dev : device [2S+1] stage times (needed when mu != NULL)
 *   rho          : (B, n, n), in/out
 * Replaces RK4_solver.take_step + the fixed_step_solver_template loop (solvers/fixed_step_solvers.py:43-77, 441-454) applied
 * to LindbladModel.evaluate_rhs with vectorized=False (the route solve_lmde(method="RK4") takes, solvers/solver_functions.py:315-327). */
int qdb_lindblad_rk4_steps_c128(int n, int J, int B, int S,
                                const qdb_c128* m1_table, const qdb_c128* m2t_table, const qdb_c128* diss_packed,
                                const double* gamma_table, const double* mu, const double* times_dev, double h,
                                qdb_c128* rho, void* stream);

/* Number of kernels this library has launched on the calling process since load (bench.py's
 * "gpu_launches" evidence). */
unsigned long long qdb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QDB_H */
