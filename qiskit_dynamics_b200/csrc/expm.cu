// Matrix exponential by scaling and squaring with a degree-16 Taylor polynomial evaluated in
// Paterson-Stockmeyer form (6 products) -- products only, so the whole propagator is built from the
// DMMA GEMM (replaces scipy.linalg.expm, solvers/fixed_step_solvers.py:22,104).
//
//   A2 = A A, A3 = A2 A, A4 = A2 A2
//   exp(A) ~ P0 + A4 (P1 + A4 (P2 + A4 P3)),  P_i = sum_{k<4} A^k / (4i+k)!   (P3 also has A4/16!)
// The caller scales A by 2^-s beforehand such that ||A||_1 <= 0.7 (truncation error 0.7^17/17! < 1e-17)
// and passes s; the result is squared s times.
#include "qdb_common.cuh"

namespace qdb {

int expm_core(int n, const double2* As, int squarings, double2* out, double2* ws, cudaStream_t st) {
    const size_t e = (size_t)n * n;
    double2* A2 = ws;
    double2* A3 = ws + e;
    double2* A4 = ws + 2 * e;
    double2* T1 = ws + 3 * e;
    double2* T2 = ws + 4 * e;
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
    double f[17];
    f[0] = 1.0;
    for (int k = 1; k <= 16; ++k) f[k] = f[k - 1] / k;  // 1/k!
    int rc;
#define GEMM(Cp, Ap, Bp, beta)                                                                      \
    if ((rc = launch_zgemm(n, n, n, Ap, n, Bp, n, Cp, n, one, beta, nullptr, nullptr, nullptr, st)) \
        != QDB_OK) return rc
    GEMM(A2, As, As, zero);
    GEMM(A3, A2, As, zero);
    GEMM(A4, A2, A2, zero);
    // B3
    if ((rc = launch_poly(n, f[12], f[13], As, f[14], A2, f[15], A3, f[16], A4, T1, st)) != QDB_OK) return rc;
    // B2 = P2 + A4 B3
    if ((rc = launch_poly(n, f[8], f[9], As, f[10], A2, f[11], A3, 0.0, nullptr, T2, st)) != QDB_OK) return rc;
    GEMM(T2, A4, T1, one);
    // B1 = P1 + A4 B2
    if ((rc = launch_poly(n, f[4], f[5], As, f[6], A2, f[7], A3, 0.0, nullptr, T1, st)) != QDB_OK) return rc;
    GEMM(T1, A4, T2, one);
    // B0 = P0 + A4 B1  -> T2, or straight into `out` when no squaring follows
    double2* dst = squarings == 0 ? out : T2;
    if ((rc = launch_poly(n, f[0], f[1], As, f[2], A2, f[3], A3, 0.0, nullptr, dst, st)) != QDB_OK) return rc;
    GEMM(dst, A4, T1, one);
    double2* cur = dst;
    for (int s = 0; s < squarings; ++s) {
        double2* nxt = (s == squarings - 1) ? out : (cur == T2 ? T1 : T2);
        GEMM(nxt, cur, cur, zero);
        cur = nxt;
    }
#undef GEMM
    return QDB_OK;
}

}  // namespace qdb

// ------------------------------------------------------------------------------------------------
// Magnus exponents of orders 2 and 3 (SURVEY.md 8(a) row a9 beyond first order; replaces the
// magnus_order == 2 / 3 branches of get_exponential_take_step, solvers/fixed_step_solvers.py:348-395).
// g holds the generator at the Gauss-Legendre nodes of the step (contiguous, stride n*n, row-major);
// out = scale * Omega(h).  Commutators are pairs of DMMA GEMMs (C = A B, then C -= B A), the linear
// combinations run through poly_kernel.  ws: n^2 (order 2) / 7 n^2 (order 3) complex numbers.
// ------------------------------------------------------------------------------------------------
namespace qdb {

int magnus_terms(int n, int order, const double2* g, double h, double scale, double2* out, double2* ws, cudaStream_t st) {
    const size_t e = (size_t)n * n;
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0), minus = make_double2(-1.0, 0.0);
    int rc;
#define COMMUTATOR(Cp, Ap, Bp)                                                                                          \
    if ((rc = launch_zgemm(n, n, n, Ap, n, Bp, n, Cp, n, one, zero, nullptr, nullptr, nullptr, st)) != QDB_OK) return rc; \
    if ((rc = launch_zgemm(n, n, n, Bp, n, Ap, n, Cp, n, minus, one, nullptr, nullptr, nullptr, st)) != QDB_OK) return rc
    if (order == 1) return launch_poly(n, 0.0, scale * h, g, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr, out, st);
    if (order == 2) {
        // Omega = h (g1 + g2) / 2 + (sqrt(3) / 12) h^2 [g2, g1]
        const double2 *g1 = g, *g2 = g + e;
        double2* C = ws;
        COMMUTATOR(C, g2, g1);
        const double p2 = sqrt(3.0) / 12.0;
        return launch_poly(n, 0.0, scale * h * 0.5, g1, scale * h * 0.5, g2, scale * p2 * h * h, C, 0.0, nullptr, out, st);
    }
    if (order == 3) {
        const double2 *g1 = g, *g2 = g + e, *g3 = g + 2 * e;
        double2 *a1 = ws, *a2 = ws + e, *a3 = ws + 2 * e, *c1 = ws + 3 * e, *X = ws + 4 * e, *Q = ws + 5 * e, *R = ws + 6 * e;
        const double k0 = sqrt(15.0) / 3.0 * h, k1 = 10.0 / 3.0 * h;
        // a1 = h g2, a2 = (sqrt(15)/3) h (g3 - g1), a3 = (10/3) h (g3 - 2 g2 + g1)
        if ((rc = launch_poly(n, 0.0, h, g2, 0.0, nullptr, 0.0, nullptr, 0.0, nullptr, a1, st)) != QDB_OK) return rc;
        if ((rc = launch_poly(n, 0.0, k0, g3, -k0, g1, 0.0, nullptr, 0.0, nullptr, a2, st)) != QDB_OK) return rc;
        if ((rc = launch_poly(n, 0.0, k1, g3, -2.0 * k1, g2, k1, g1, 0.0, nullptr, a3, st)) != QDB_OK) return rc;
        COMMUTATOR(c1, a1, a2);                                                             // comm1 = [a1, a2]
        if ((rc = launch_poly(n, 0.0, 2.0, a3, 1.0, c1, 0.0, nullptr, 0.0, nullptr, X, st)) != QDB_OK) return rc;
        COMMUTATOR(Q, X, a1);                                                               // 60 comm2 = [2 a3 + comm1, a1]
        double2* L = X;                                                                     // X is free again
        if ((rc = launch_poly(n, 0.0, -20.0, a1, -1.0, a3, 1.0, c1, 0.0, nullptr, L, st)) != QDB_OK) return rc;
        if ((rc = launch_poly(n, 0.0, 1.0, a2, 1.0 / 60.0, Q, 0.0, nullptr, 0.0, nullptr, R, st)) != QDB_OK) return rc;
        COMMUTATOR(Q, L, R);                                                                // [-20 a1 - a3 + comm1, a2 + comm2]
        return launch_poly(n, 0.0, scale, a1, scale / 12.0, a3, scale / 240.0, Q, 0.0, nullptr, out, st);
    }
#undef COMMUTATOR
    set_error("magnus_terms: order %d not in {1, 2, 3}", order);
    return QDB_E_ARG;
}

}  // namespace qdb

// ------------------------------------------------------------------------------------------------
// fp64 tensor-pipe peak probe: every warp issues independent DMMA m8n8k4 chains from registers.
// Used by bench.py as the live roofline denominator (MEASURED_PEAKS.json has no fp64 entry).
// ------------------------------------------------------------------------------------------------
namespace qdb {

__global__ void __launch_bounds__(256) dmma_probe_kernel(double* sink, int iters, double a0, double b0) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

int launch_dmma_probe(double* sink, int iters, int* grid_out, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 2;
    dmma_probe_kernel<<<grid, 256, 0, st>>>(sink, iters, 1.0, 1.0);
    QDB_LAUNCH_CHECK("dmma_probe_kernel");
    if (grid_out) *grid_out = grid;
    return QDB_OK;
}

}  // namespace qdb
