// Time-parallel fixed-step solvers for LMDEs (the reference's jax_RK4_parallel_solver / jax_expm_parallel_solver,
// solvers/fixed_step_solvers.py:206-244, 279-311, and their template :524-613): every step has a propagator that does
// not depend on the state,
//     RK4 :  P = 1 + (h/6)(k1 + 2 k2 + 2 k3 + k4),  k1 = G(t), k2 = G(t+h/2)(1 + h/2 k1), k3 = G(t+h/2)(1 + h/2 k2),
//            k4 = G(t+h)(1 + h k3)
//     expm:  P = expm(Omega)  (Magnus order 1, 2 or 3, see expm.cu)
// so all S propagators of an interval are built side by side and multiplied together pairwise,
//     P_total = P_{S-1} ... P_1 P_0,
// and the state batch is touched once per interval.  (The exponentials -- and at Magnus orders 2 and 3 the exponents with
// their commutators -- are batched over the steps as well.)  The reference does this with jax.vmap + associative_scan; here the
// step dimension is the batch dimension (grid.z) of one DMMA GEMM launch per RK4 stage and per level of the product
// tree -- S products of 128^3 fill the chip where a single one occupies four CTAs.  For n x n generators and B columns
// a step costs 4 (8 n^3) flops instead of 4 (8 n^2 B): the shared-signal shortcut of SURVEY 8(d) (32x fewer flops at
// n = 128, B = 4096), offered under the reference's own method names; it does not exist for per-column signals.
#include "qdb_common.cuh"

namespace qdb {

namespace {

// P[s] = 1 + (1/6) h (k1 + 2 k2 + 2 k3 + k4), k1 = G[3 s]  (evaluation order of fixed_step_solvers.py:237)
__global__ void rk4_prop_combine_kernel(int n, size_t nn, size_t total, double h, const double2* __restrict__ G,
                                        const double2* __restrict__ K2, const double2* __restrict__ K3,
                                        const double2* __restrict__ K4, double2* __restrict__ P) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t s = idx / nn, e = idx - s * nn;
    const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
    const double2 k1 = G[3 * s * nn + e], k2 = K2[idx], k3 = K3[idx], k4 = K4[idx];
    const double f = (1.0 / 6) * h;
    double2 v = make_double2(f * (k1.x + 2 * k2.x + 2 * k3.x + k4.x), f * (k1.y + 2 * k2.y + 2 * k3.y + k4.y));
    if (r == c) v.x += 1.0;
    P[idx] = v;
}

inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// gather every third matrix of G (offset `which`) into a dense [S][nn] array
int gather_third(size_t nn, int S, const double2* G, int which, double2* dst, cudaStream_t st) {
    QDB_CUDA(cudaMemcpy2DAsync(dst, nn * sizeof(double2), G + (size_t)which * nn, 3 * nn * sizeof(double2), nn * sizeof(double2),
                               (size_t)S, cudaMemcpyDeviceToDevice, st));
    return QDB_OK;
}

// RK4 step propagators of S steps from G = [S][3][n][n] (generator at t, t + h/2, t + h): four launches over all steps
int rk4_step_propagators(int n, int S, const double2* G, double h, double2* K2, double2* K3, double2* K4, double2* P,
                         cudaStream_t st) {
    const size_t nn = (size_t)n * n;
    const double2 one = make_double2(1.0, 0.0);
    int rc;
    // k2 = Gm + (h/2) Gm k1
    if ((rc = gather_third(nn, S, G, 1, K2, st)) != QDB_OK) return rc;
    if ((rc = launch_zgemm_batched(n, n, n, G + nn, n, 3 * (long long)nn, G, n, 3 * (long long)nn, K2, n, (long long)nn,
                                   make_double2(0.5 * h, 0.0), one, S, st)) != QDB_OK) return rc;
    // k3 = Gm + (h/2) Gm k2
    if ((rc = gather_third(nn, S, G, 1, K3, st)) != QDB_OK) return rc;
    if ((rc = launch_zgemm_batched(n, n, n, G + nn, n, 3 * (long long)nn, K2, n, (long long)nn, K3, n, (long long)nn,
                                   make_double2(0.5 * h, 0.0), one, S, st)) != QDB_OK) return rc;
    // k4 = G1 + h G1 k3
    if ((rc = gather_third(nn, S, G, 2, K4, st)) != QDB_OK) return rc;
    if ((rc = launch_zgemm_batched(n, n, n, G + 2 * nn, n, 3 * (long long)nn, K3, n, (long long)nn, K4, n, (long long)nn,
                                   make_double2(h, 0.0), one, S, st)) != QDB_OK) return rc;
    const size_t total = (size_t)S * nn;
    rk4_prop_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(n, nn, total, h, G, K2, K3, K4, P);
    QDB_LAUNCH_CHECK("rk4_prop_combine_kernel");
    return QDB_OK;
}

// out = P[count-1] ... P[1] P[0]; P is overwritten, tmp holds ceil(count / 2) matrices.  One batched GEMM per level.
int product_tree(int n, int count, double2* P, double2* tmp, double2* out, cudaStream_t st) {
    const size_t nn = (size_t)n * n;
    double2 *cur = P, *other = tmp;
    int m = count, rc;
    while (m > 1) {
        const int pairs = m / 2;
        if ((rc = launch_zgemm_batched(n, n, n, cur + nn, n, 2 * (long long)nn, cur, n, 2 * (long long)nn, other, n, (long long)nn,
                                       make_double2(1.0, 0.0), make_double2(0.0, 0.0), pairs, st)) != QDB_OK) return rc;
        if (m & 1)
            QDB_CUDA(cudaMemcpyAsync(other + (size_t)pairs * nn, cur + (size_t)(m - 1) * nn, nn * sizeof(double2),
                                     cudaMemcpyDeviceToDevice, st));
        double2* t = cur;
        cur = other;
        other = t;
        m = pairs + (m & 1);
    }
    QDB_CUDA(cudaMemcpyAsync(out, cur, nn * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    return QDB_OK;
}

}  // namespace

// out[z] = c0 1 + c1 A1[z] + c2 A2[z] + c3 A3[z] + c4 A4[z] for a batch of n x n matrices (null pointers are skipped);
// input i advances by s_i elements per matrix of the batch (the node generators of a step sit side by side), out is dense
static __global__ void poly_batched_kernel(int n, size_t nn, size_t total, double c0, double c1, const double2* __restrict__ A1,
                                           long long s1, double c2, const double2* __restrict__ A2, long long s2, double c3,
                                           const double2* __restrict__ A3, long long s3, double c4,
                                           const double2* __restrict__ A4, long long s4, double2* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t z = idx / nn, e = idx - z * nn;
    const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
    double2 v = make_double2(r == c ? c0 : 0.0, 0.0);
    if (A1) { const double2 a = A1[z * (size_t)s1 + e]; v.x = fma(c1, a.x, v.x); v.y = fma(c1, a.y, v.y); }
    if (A2) { const double2 a = A2[z * (size_t)s2 + e]; v.x = fma(c2, a.x, v.x); v.y = fma(c2, a.y, v.y); }
    if (A3) { const double2 a = A3[z * (size_t)s3 + e]; v.x = fma(c3, a.x, v.x); v.y = fma(c3, a.y, v.y); }
    if (A4) { const double2 a = A4[z * (size_t)s4 + e]; v.x = fma(c4, a.x, v.x); v.y = fma(c4, a.y, v.y); }
    out[idx] = v;
}

// expm of `count` matrices at once: the degree-16 Taylor / Paterson-Stockmeyer scheme of expm_core (expm.cu) with every
// product a batched GEMM.  As [count][nn] is already scaled by 2^-squarings (one common count: the largest any step
// needs); ws holds 5 count matrices; out [count][nn].
int expm_core_batched(int n, int count, const double2* As, int squarings, double2* out, double2* ws, cudaStream_t st) {
    const size_t nn = (size_t)n * n, tot = (size_t)count * nn;
    const long long sn = (long long)nn;
    double2 *A2 = ws, *A3 = ws + tot, *A4 = ws + 2 * tot, *T1 = ws + 3 * tot, *T2 = ws + 4 * tot;
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
    double f[17];
    f[0] = 1.0;
    for (int k = 1; k <= 16; ++k) f[k] = f[k - 1] / k;
    const unsigned blocks = (unsigned)((tot + 255) / 256);
    int rc;
#define BGEMM(Cp, Ap, Bp, beta) \
    if ((rc = launch_zgemm_batched(n, n, n, Ap, n, sn, Bp, n, sn, Cp, n, sn, one, beta, count, st)) != QDB_OK) return rc
#define BPOLY(c0, c1, c2, c3, c4, A4p, dst)                                                                  \
    poly_batched_kernel<<<blocks, 256, 0, st>>>(n, nn, tot, c0, c1, As, sn, c2, A2, sn, c3, A3, sn, c4, A4p, sn, dst); \
    QDB_LAUNCH_CHECK("poly_batched_kernel")
    BGEMM(A2, As, As, zero);
    BGEMM(A3, A2, As, zero);
    BGEMM(A4, A2, A2, zero);
    BPOLY(f[12], f[13], f[14], f[15], f[16], A4, T1);                  // B3
    BPOLY(f[8], f[9], f[10], f[11], 0.0, (const double2*)nullptr, T2);  // B2 = P2 + A4 B3
    BGEMM(T2, A4, T1, one);
    BPOLY(f[4], f[5], f[6], f[7], 0.0, (const double2*)nullptr, T1);    // B1 = P1 + A4 B2
    BGEMM(T1, A4, T2, one);
    double2* dst = squarings == 0 ? out : T2;
    BPOLY(f[0], f[1], f[2], f[3], 0.0, (const double2*)nullptr, dst);   // B0 = P0 + A4 B1
    BGEMM(dst, A4, T1, one);
    double2* cur = dst;
    for (int s = 0; s < squarings; ++s) {
        double2* nxt = (s == squarings - 1) ? out : (cur == T2 ? T1 : T2);
        BGEMM(nxt, cur, cur, zero);
        cur = nxt;
    }
#undef BGEMM
#undef BPOLY
    return QDB_OK;
}

// Magnus exponents of `count` steps at once (orders 2 and 3; see magnus_terms in expm.cu for the formulas): g holds the
// generator at the nodes, [count][order][n][n]; out [count][n][n] = scale * Omega; ws: count (order 2) / 7 count (order 3)
// matrices.  Commutators are pairs of batched GEMMs, linear combinations one strided poly launch each.
int magnus_terms_batched(int n, int order, int count, const double2* g, double h, double scale, double2* out, double2* ws,
                         cudaStream_t st) {
    const size_t nn = (size_t)n * n, tot = (size_t)count * nn;
    const long long sn = (long long)nn, sg = (long long)order * sn;
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0), minus = make_double2(-1.0, 0.0);
    const unsigned blocks = (unsigned)((tot + 255) / 256);
    const double2* none = nullptr;
    int rc;
#define BCOMM(Cp, Ap, sA, Bp, sB)                                                                                             \
    if ((rc = launch_zgemm_batched(n, n, n, Ap, n, sA, Bp, n, sB, Cp, n, sn, one, zero, count, st)) != QDB_OK) return rc;   \
    if ((rc = launch_zgemm_batched(n, n, n, Bp, n, sB, Ap, n, sA, Cp, n, sn, minus, one, count, st)) != QDB_OK) return rc
#define BLIN(c1, A1, s1, c2, A2, s2, c3, A3, s3, dst)                                                                       \
    poly_batched_kernel<<<blocks, 256, 0, st>>>(n, nn, tot, 0.0, c1, A1, s1, c2, A2, s2, c3, A3, s3, 0.0, none, 0, dst);      \
    QDB_LAUNCH_CHECK("poly_batched_kernel")
    if (order == 2) {
        const double2 *g1 = g, *g2 = g + nn;
        double2* C = ws;
        BCOMM(C, g2, sg, g1, sg);
        const double p2 = sqrt(3.0) / 12.0;
        BLIN(scale * h * 0.5, g1, sg, scale * h * 0.5, g2, sg, scale * p2 * h * h, C, sn, out);
        return QDB_OK;
    }
    if (order == 3) {
        const double2 *g1 = g, *g2 = g + nn, *g3 = g + 2 * nn;
        double2 *a1 = ws, *a2 = ws + tot, *a3 = ws + 2 * tot, *c1 = ws + 3 * tot, *X = ws + 4 * tot, *Q = ws + 5 * tot, *R = ws + 6 * tot;
        const double k0 = sqrt(15.0) / 3.0 * h, k1 = 10.0 / 3.0 * h;
        BLIN(h, g2, sg, 0.0, none, 0, 0.0, none, 0, a1);
        BLIN(k0, g3, sg, -k0, g1, sg, 0.0, none, 0, a2);
        BLIN(k1, g3, sg, -2.0 * k1, g2, sg, k1, g1, sg, a3);
        BCOMM(c1, a1, sn, a2, sn);                                   // comm1 = [a1, a2]
        BLIN(2.0, a3, sn, 1.0, c1, sn, 0.0, none, 0, X);
        BCOMM(Q, X, sn, a1, sn);                                     // 60 comm2 = [2 a3 + comm1, a1]
        double2* L = X;
        BLIN(-20.0, a1, sn, -1.0, a3, sn, 1.0, c1, sn, L);
        BLIN(1.0, a2, sn, 1.0 / 60.0, Q, sn, 0.0, none, 0, R);
        BCOMM(Q, L, sn, R, sn);                                      // [-20 a1 - a3 + comm1, a2 + comm2]
        BLIN(scale, a1, sn, scale / 12.0, a3, sn, scale / 240.0, Q, sn, out);
        return QDB_OK;
    }
#undef BCOMM
#undef BLIN
    set_error("magnus_terms_batched: order %d not in {2, 3}", order);
    return QDB_E_ARG;
}

size_t propagator_workspace_bytes(int n, int S) {
    const size_t nn = align256((size_t)n * n * sizeof(double2));
    if (S < 1) S = 1;
    // per step: P + {RK4: G (3) + k2..k4 (3); exponential: A + Taylor (5)}; the product tree ceil(S / 2); 20 fixed matrices
    // (two accumulators + scratch for the step-by-step Magnus route); the node times.  Magnus orders 2 and 3 batch fewer
    // steps into the same room (17 matrices per step).
    return (size_t)S * 7 * nn + (size_t)((S + 1) / 2) * nn + 20 * nn + align256((size_t)3 * S * sizeof(double));
}

// kind 0: RK4 (times/coeff [S][3]); kind 1..3: exponential of the Magnus exponent of that order (times/coeff [S][kind])
int step_propagator_product(int n, int K, int S, int kind, const double2* ops_rm, const double2* stat_rm, const double* coeff,
                            const double* mu, const double* times_host, const int* squarings_host, double h, double2* P_total,
                            void* workspace, size_t ws_bytes, cudaStream_t st) {
    const size_t nn = (size_t)n * n, nnb = align256(nn * sizeof(double2));
    const int Q = kind == 0 ? 3 : kind;  // generator evaluations per step
    // matrices per step behind the 20 fixed ones: P + tree/2 + {RK4: G (3) + k2..k4 (3); order 1: A + Taylor (5);
    // orders 2, 3: node generators (Q) + Magnus temporaries (1 / 7) + A + Taylor (5)}
    const size_t big_per_step = kind <= 1 ? 6 : (size_t)Q + (kind == 2 ? 1 : 7) + 6;
    auto need = [&](int steps) {
        return (size_t)steps * (1 + big_per_step) * nnb + (size_t)((steps + 1) / 2) * nnb + 20 * nnb +
               align256((size_t)3 * steps * sizeof(double));
    };
    // largest chunk of steps that fits the workspace
    int Sc = S;
    while (Sc > 1 && need(Sc) > ws_bytes) Sc = (Sc + 1) / 2;
    if (need(Sc) > ws_bytes) {
        if (kind >= 2 && propagator_workspace_bytes(n, 1) <= ws_bytes) {
            Sc = 0;  // room for the step-by-step route only (its temporaries live in the 18 scratch matrices)
        } else {
            set_error("qdb_step_propagators_c128: workspace too small (%zu < %zu)", ws_bytes, propagator_workspace_bytes(n, 1));
            return QDB_E_WORKSPACE;
        }
    }
    const bool stepwise = (Sc == 0);
    if (stepwise) {
        Sc = S;
        while (Sc > 1 && propagator_workspace_bytes(n, Sc) > ws_bytes) Sc = (Sc + 1) / 2;
    }
    char* ws = (char*)workspace;
    double2* acc_a = (double2*)ws;                      // running product
    double2* acc_b = (double2*)(ws + nnb);              // chunk product / swap
    double2* scratch = (double2*)(ws + 2 * nnb);        // 18 matrices: expm As, core (5), node generators (3), Magnus (7), spare
    double2* P = (double2*)(ws + 20 * nnb);             // [Sc]
    double2* tree = P + (size_t)Sc * nn;                // [ceil(Sc/2)]
    double2* big = tree + (size_t)((Sc + 1) / 2) * nn;  // RK4: G [3 Sc], K2, K3, K4 [Sc each]
    double* times_dev = (double*)(ws + (stepwise ? propagator_workspace_bytes(n, Sc) : need(Sc)) - align256((size_t)3 * Sc * sizeof(double)));
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
    int rc;
    bool first = true;
    for (int s0 = 0; s0 < S; s0 += Sc) {
        const int Sn = S - s0 < Sc ? S - s0 : Sc;
        const double* cs = coeff ? coeff + (size_t)s0 * Q * K : nullptr;
        if (mu) QDB_CUDA(cudaMemcpyAsync(times_dev, times_host + (size_t)s0 * Q, (size_t)Sn * Q * sizeof(double), cudaMemcpyHostToDevice, st));
        if (kind == 0) {
            double2 *G = big, *K2 = big + (size_t)3 * Sn * nn, *K3 = K2 + (size_t)Sn * nn, *K4 = K3 + (size_t)Sn * nn;
            rc = launch_generator(n, K, 3 * Sn, QDB_LAYOUT_ROWMAJOR, ops_rm, stat_rm, cs, 0, mu, mu ? times_dev : nullptr, 0.0, 1.0, G, st);
            if (rc != QDB_OK) return rc;
            if ((rc = rk4_step_propagators(n, Sn, G, h, K2, K3, K4, P, st)) != QDB_OK) return rc;
        } else if (Q == 1) {
            // Magnus order 1: one generator launch writes (h / 2^sq) G(t_s + h/2) for every step, one batched exponential
            int sq = 0;
            for (int s = 0; s < Sn; ++s) {
                const int v = squarings_host[s0 + s];
                if (v < 0 || v >= 64) {
                    set_error("qdb_step_propagators_c128: bad squarings[%d]=%d", s0 + s, v);
                    return QDB_E_ARG;
                }
                sq = v > sq ? v : sq;
            }
            double2 *As = big, *bws = big + (size_t)Sn * nn;  // As [Sn] + 5 [Sn] of the 6 Sc matrices of `big`
            rc = launch_generator(n, K, Sn, QDB_LAYOUT_ROWMAJOR, ops_rm, stat_rm, cs, 0, mu, mu ? times_dev : nullptr, 0.0,
                                  ldexp(h, -sq), As, st);
            if (rc != QDB_OK) return rc;
            if ((rc = expm_core_batched(n, Sn, As, sq, P, bws, st)) != QDB_OK) return rc;
        } else if (!stepwise) {
            // Magnus orders 2, 3: node generators of every step in one launch, exponents and exponentials batched
            int sq = 0;
            for (int s = 0; s < Sn; ++s) {
                const int v = squarings_host[s0 + s];
                if (v < 0 || v >= 64) {
                    set_error("qdb_step_propagators_c128: bad squarings[%d]=%d", s0 + s, v);
                    return QDB_E_ARG;
                }
                sq = v > sq ? v : sq;
            }
            double2* gn = big;                                             // [Sn][Q]
            double2* mws = gn + (size_t)Sn * Q * nn;                       // [Sn] or [7 Sn]
            double2* As = mws + (size_t)Sn * (kind == 2 ? 1 : 7) * nn;     // [Sn]
            double2* bws = As + (size_t)Sn * nn;                           // [5 Sn]
            rc = launch_generator(n, K, Sn * Q, QDB_LAYOUT_ROWMAJOR, ops_rm, stat_rm, cs, 0, mu, mu ? times_dev : nullptr, 0.0, 1.0, gn, st);
            if (rc != QDB_OK) return rc;
            if ((rc = magnus_terms_batched(n, Q, Sn, gn, h, ldexp(1.0, -sq), As, mws, st)) != QDB_OK) return rc;
            if ((rc = expm_core_batched(n, Sn, As, sq, P, bws, st)) != QDB_OK) return rc;
        } else {
            double2 *As = scratch, *core_ws = scratch + nn, *gnodes = scratch + 6 * nn, *mag_ws = scratch + 9 * nn;
            for (int s = 0; s < Sn; ++s) {
                const int sq = squarings_host[s0 + s];
                if (sq < 0 || sq >= 64) {
                    set_error("qdb_step_propagators_c128: bad squarings[%d]=%d", s0 + s, sq);
                    return QDB_E_ARG;
                }
                const double sc = ldexp(1.0, -sq);
                const double* c1 = cs ? cs + (size_t)s * Q * K : nullptr;
                if (Q == 1) {
                    rc = launch_generator(n, K, 1, QDB_LAYOUT_ROWMAJOR, ops_rm, stat_rm, c1, 0, mu, mu ? times_dev + s : nullptr, 0.0, sc * h, As, st);
                    if (rc != QDB_OK) return rc;
                } else {
                    rc = launch_generator(n, K, Q, QDB_LAYOUT_ROWMAJOR, ops_rm, stat_rm, c1, 0, mu, mu ? times_dev + (size_t)s * Q : nullptr, 0.0,
                                          1.0, gnodes, st);
                    if (rc != QDB_OK) return rc;
                    if ((rc = magnus_terms(n, Q, gnodes, h, sc, As, mag_ws, st)) != QDB_OK) return rc;
                }
                if ((rc = expm_core(n, As, sq, P + (size_t)s * nn, core_ws, st)) != QDB_OK) return rc;
            }
        }
        double2* dst = first ? acc_a : acc_b;
        if ((rc = product_tree(n, Sn, P, tree, dst, st)) != QDB_OK) return rc;
        if (!first) {  // total <- chunk * total
            if ((rc = launch_zgemm(n, n, n, acc_b, n, acc_a, n, scratch, n, one, zero, nullptr, nullptr, nullptr, st)) != QDB_OK) return rc;
            QDB_CUDA(cudaMemcpyAsync(acc_a, scratch, nn * sizeof(double2), cudaMemcpyDeviceToDevice, st));
        }
        first = false;
    }
    QDB_CUDA(cudaMemcpyAsync(P_total, acc_a, nn * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    return QDB_OK;
}

}  // namespace qdb
