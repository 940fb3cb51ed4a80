// Non-vectorised Lindblad equation on a batch of density matrices:
//
//     rhs(rho) = M1 rho + rho M2 + sum_j g_j L_j rho L_j^dag ,
//     M1 = A + B,  M2 = A - B,  B = -i H(t),  A = -1/2 sum_j g_j L_j^dag L_j        (g_j = 1 for static dissipators)
//
// which is LindbladCollection.evaluate_rhs of the reference (models/operator_collections.py:451-567; batch = LEADING
// axis, :506-510) -- O(n^3) per density matrix instead of the O(n^4) of the vectorised form.  In a rotating frame the
// model conjugates with the frame phases around the collection call (models/lindblad_model.py:520-538,
// models/rotating_frame.py:350-353): X = rho .* (p_i conj p_k), out = rhs(X) .* (conj p_i p_k), p = exp(-i mu t).
//
// One CTA owns ONE density matrix (n <= 32) for the whole launch; the fixed-step RK4 loop runs on chip:
//   * rho, the RK4 k-sum and the product accumulators live in registers in DMMA C-fragment order;
//   * the stage input X lives in shared memory twice -- row-major (A operand of the right products) and in B-fragment
//     order (B operand of the left products) -- and the intermediate W_j = g_j X L_j^dag in B-fragment order, double
//     buffered, so that a dissipator costs one __syncthreads;
//   * every operator streams from L2 in DMMA A-fragment order (QDB_LAYOUT_PACKED): M1(t) and M2(t)^T are entries of two
//     generator tables built by generator_kernel for all stage times of the interval, L_j is used as stored -- as the A
//     operand of L_j W_j, and, conjugated, as the B operand of X L_j^dag (B[k][c] = conj(L[c][k]) is what the lane that
//     holds A[c][k] loads).  All CTAs walk the same operators: one HBM read, then L2.
// 4-product complex tiles (cr += ar br - ai bi, ci += ar bi + ai br).  Roofline: fp64 tensor pipe,
// (2 + 2 J) 8 n^3 flops per density matrix and RHS evaluation.
#include "qdb_common.cuh"
#include "rk4_device.cuh"

namespace qdb {

namespace {

constexpr int LD = 36;  // row pitch (in double2) of the row-major stage copy: A-fragment loads are bank-conflict free

struct Tile {
    double cr[2], ci[2];
    __device__ __forceinline__ void zero() { cr[0] = cr[1] = ci[0] = ci[1] = 0.0; }
};

// acc += a * b  (complex 8x4 by 4x8 fragment product), conj_b: b replaced by its conjugate
template <bool CONJ_B>
__device__ __forceinline__ void cdmma(Tile& t, double2 a, double2 b) {
    const double bi = CONJ_B ? negate(b.y) : b.y;
    dmma(t.cr[0], t.cr[1], a.x, b.x);
    dmma(t.ci[0], t.ci[1], a.x, bi);
    dmma(t.cr[0], t.cr[1], negate(a.y), bi);
    dmma(t.ci[0], t.ci[1], a.y, b.x);
}

// NT = row tiles (npad / 8 <= 4); TPW = output tiles per warp (NT^2 <= 8 TPW); RK4 = step loop (else one RHS evaluation)
template <int NT, int TPW, bool RK4>
__global__ void __launch_bounds__(256, 2)
lindblad_kernel(int n, int J, int B, int S, const double2* __restrict__ m1, const double2* __restrict__ m2t,
                const double2* __restrict__ diss, const double* __restrict__ gam, const double* __restrict__ mu,
                const double* __restrict__ times, double t_scalar, double h, const double2* __restrict__ rho_in,
                double2* __restrict__ rho_out) {
    constexpr int KT = NT <= 2 ? 4 : 8;  // k-tiles of the packed layout: kpad = 16 or 32
    extern __shared__ __align__(16) double2 lsm[];
    double2* xa = lsm;                                       // X row-major [NT * 8][LD]
    double2* xb = xa + NT * 8 * LD;                          // X in B-fragment order [kt][ct][lane]
    double2* wb0 = xb + KT * NT * 32;                        // W_j in B-fragment order, double buffered
    double2* wb1 = wb0 + KT * NT * 32;
    double2* ph = wb1 + KT * NT * 32;                        // p_i = exp(-i mu_i t) of the stage [NT * 8]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int swl = frag_swizzle(lane);
    const int b = blockIdx.x;
    const size_t entry_elems = (size_t)NT * 8 * KT * 4;  // one packed operator
    const bool framed = mu != nullptr;

    int trt[TPW], tct[TPW];
    bool tv[TPW];
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
        const int t = warp + 8 * i;
        tv[i] = t < NT * NT;
        trt[i] = tv[i] ? t / NT : 0;
        tct[i] = tv[i] ? t % NT : 0;
    }
    // zero the operand buffers once: rows / columns beyond n stay zero for the whole launch
    for (int i = tid; i < NT * 8 * LD; i += 256) xa[i] = make_double2(0.0, 0.0);
    for (int i = tid; i < KT * NT * 32; i += 256) xb[i] = wb0[i] = wb1[i] = make_double2(0.0, 0.0);

    double2 yv[TPW][2];   // rho elements (row 8 rt + g, columns 8 ct + 2q, 2q + 1)
    double2 ks[TPW][2];   // RK4 k-sum
    const double2* src = rho_in + (size_t)b * n * n;
#pragma unroll
    for (int i = 0; i < TPW; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int r = 8 * trt[i] + g, c = 8 * tct[i] + 2 * q + e;
            yv[i][e] = (tv[i] && r < n && c < n) ? src[(size_t)r * n + c] : make_double2(0.0, 0.0);
            ks[i][e] = make_double2(0.0, 0.0);
        }
    __syncthreads();

    // writes the stage input X = v .* (p_r conj p_c) of this thread's elements into both shared copies
    auto put_x = [&](int i, int e, double2 v) {
        const int r = 8 * trt[i] + g, c = 8 * tct[i] + 2 * q + e;
        if (!tv[i]) return;
        if (framed) v = cmul(v, cmul_conj_a(ph[c], ph[r]));  // p_r * conj(p_c)
        xa[r * LD + c] = v;
        // B-fragment order: k-tile = r / 4, lane = (r % 4) + 4 (c % 8), column tile = c / 8
        xb[((r >> 2) * NT + (c >> 3)) * 32 + frag_swizzle((r & 3) + 4 * (c & 7))] = v;
    };
    auto stage_phases = [&](double t) {
        if (framed && tid < NT * 8) ph[tid] = tid < n ? frame_phase(mu[tid], t) : make_double2(1.0, 0.0);
    };

    const int total = RK4 ? 4 * S : 1;
    stage_phases(RK4 ? (framed ? times[0] : 0.0) : t_scalar);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TPW; ++i)
#pragma unroll
        for (int e = 0; e < 2; ++e) put_x(i, e, yv[i][e]);
    __syncthreads();

#pragma unroll 1
    for (int sidx = 0; sidx < total; ++sidx) {
        const int step = sidx >> 2, stage = sidx & 3;
        const int entry = RK4 ? 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1)) : 0;
        const double2* M1 = m1 + (size_t)entry * entry_elems;
        const double2* M2T = m2t + (size_t)entry * entry_elems;
        Tile acc[TPW];
#pragma unroll
        for (int i = 0; i < TPW; ++i) acc[i].zero();

        // ---- P = M1 X + X M2 ----
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            if (!tv[i]) continue;
            const int rt = trt[i], ct = tct[i];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const double2 a1 = __ldg(M1 + ((size_t)rt * KT + kt) * 32 + lane);
                const double2 b1 = xb[(kt * NT + ct) * 32 + swl];
                cdmma<false>(acc[i], a1, b1);
                const double2 a2 = xa[(8 * rt + g) * LD + 4 * kt + q];
                const double2 b2 = __ldg(M2T + ((size_t)ct * KT + kt) * 32 + lane);
                cdmma<false>(acc[i], a2, b2);
            }
        }
        // ---- + sum_j g_j L_j (X L_j^dag) ----
#pragma unroll 1
        for (int j = 0; j < J; ++j) {
            const double2* L = diss + (size_t)j * entry_elems;
            const double gj = gam ? gam[(size_t)entry * J + j] : 1.0;
            double2* wj = (j & 1) ? wb1 : wb0;
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                if (!tv[i]) continue;
                const int rt = trt[i], ct = tct[i];
                Tile w;
                w.zero();
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    const double2 a = xa[(8 * rt + g) * LD + 4 * kt + q];
                    const double2 bl = __ldg(L + ((size_t)ct * KT + kt) * 32 + lane);  // conj -> L^dag[k][c]
                    cdmma<true>(w, a, bl);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r = 8 * rt + g, c = 8 * ct + 2 * q + e;
                    wj[((r >> 2) * NT + (c >> 3)) * 32 + frag_swizzle((r & 3) + 4 * (c & 7))] = make_double2(gj * w.cr[e], gj * w.ci[e]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                if (!tv[i]) continue;
                const int rt = trt[i], ct = tct[i];
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    const double2 a = __ldg(L + ((size_t)rt * KT + kt) * 32 + lane);
                    const double2 bw = wj[(kt * NT + ct) * 32 + swl];
                    cdmma<false>(acc[i], a, bw);
                }
            }
        }

        // ---- epilogue: out of the frame phases, RK4 combine, next stage input ----
        double2 k[TPW][2];
#pragma unroll
        for (int i = 0; i < TPW; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                k[i][e] = make_double2(acc[i].cr[e], acc[i].ci[e]);
                if (framed && tv[i]) {
                    const int r = 8 * trt[i] + g, c = 8 * tct[i] + 2 * q + e;
                    k[i][e] = cmul(k[i][e], cmul_conj_a(ph[r], ph[c]));  // conj(p_r) p_c
                }
            }
        if (!RK4) {
            double2* dst = rho_out + (size_t)b * n * n;
#pragma unroll
            for (int i = 0; i < TPW; ++i)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int r = 8 * trt[i] + g, c = 8 * tct[i] + 2 * q + e;
                    if (tv[i] && r < n && c < n) dst[(size_t)r * n + c] = k[i][e];
                }
            return;
        }
        __syncthreads();  // every warp is done with X, W and the phases of this stage
        const StageCoef sc(stage, h);
        if (sidx + 1 < total) {
            const int nstage = (stage + 1) & 3, nstep = step + (stage == 3 ? 1 : 0);
            const int nentry = 2 * nstep + (nstage == 0 ? 0 : (nstage == 3 ? 2 : 1));
            if (nentry != entry) stage_phases(framed ? times[nentry] : 0.0);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < TPW; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                ks[i][e].x = sc.keep * ks[i][e].x + sc.wk * k[i][e].x;
                ks[i][e].y = sc.keep * ks[i][e].y + sc.wk * k[i][e].y;
                const double2 v = sc.last ? ks[i][e] : k[i][e];
                const double2 nxt = make_double2(yv[i][e].x + sc.astep * v.x, yv[i][e].y + sc.astep * v.y);
                if (sc.last) yv[i][e] = nxt;
                put_x(i, e, nxt);
            }
        __syncthreads();
    }
    if (RK4) {
        double2* dst = rho_out + (size_t)b * n * n;
#pragma unroll
        for (int i = 0; i < TPW; ++i)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = 8 * trt[i] + g, c = 8 * tct[i] + 2 * q + e;
                if (tv[i] && r < n && c < n) dst[(size_t)r * n + c] = yv[i][e];
            }
    }
}

template <bool RK4>
int launch_t(int n, int J, int B, int S, const double2* m1, const double2* m2t, const double2* diss, const double* gam,
             const double* mu, const double* times, double t_scalar, double h, const double2* rho_in, double2* rho_out,
             cudaStream_t st) {
    const int NT = round_up8(n) / 8;
#define QDB_LB(NTv, TPWv)                                                                                                     \
    if (NT == NTv) {                                                                                                          \
        constexpr int KTv = NTv <= 2 ? 4 : 8;                                                                                 \
        constexpr size_t smem = (size_t)(NTv * 8 * LD + 3 * KTv * NTv * 32 + NTv * 8) * sizeof(double2);                      \
        QDB_CUDA(cudaFuncSetAttribute(lindblad_kernel<NTv, TPWv, RK4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        lindblad_kernel<NTv, TPWv, RK4><<<B, 256, smem, st>>>(n, J, B, S, m1, m2t, diss, gam, mu, times, t_scalar, h, rho_in, rho_out); \
        QDB_LAUNCH_CHECK("lindblad_kernel");                                                                                  \
        return QDB_OK;                                                                                                        \
    }
    QDB_LB(1, 1)
    QDB_LB(2, 1)
    QDB_LB(3, 2)
    QDB_LB(4, 2)
#undef QDB_LB
    set_error("lindblad kernels: on-chip path needs n <= 32 (got %d)", n);
    return QDB_E_UNSUPPORTED;
}

}  // namespace

bool lindblad_fused_supported(int n) { return n >= 1 && n <= 32; }

int launch_lindblad_rhs(int n, int J, int B, const double2* m1, const double2* m2t, const double2* diss, const double* gam,
                        const double* mu, double t, const double2* rho_in, double2* rho_out, cudaStream_t st) {
    return launch_t<false>(n, J, B, 0, m1, m2t, diss, gam, mu, nullptr, t, 0.0, rho_in, rho_out, st);
}

int launch_lindblad_rk4(int n, int J, int B, int S, const double2* m1_table, const double2* m2t_table, const double2* diss,
                        const double* gam_table, const double* mu, const double* times_dev, double h, double2* rho,
                        cudaStream_t st) {
    return launch_t<true>(n, J, B, S, m1_table, m2t_table, diss, gam_table, mu, times_dev, 0.0, h, rho, rho, st);
}

}  // namespace qdb
