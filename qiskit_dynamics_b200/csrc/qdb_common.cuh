// Shared device helpers for libqdb (sm_100a).  fp64 tensor work on Blackwell is the warp-level
// DMMA (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4); tcgen05 has no f64 kind.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qdb.h"

namespace qdb {

// ---- error plumbing (host) --------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
int sm_count();  // SMs of the current device (148 on B200 when the query fails), per-device cached

#define QDB_CUDA(call)                                        \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return qdb::cuda_fail(_e, #call); \
    } while (0)

#define QDB_LAUNCH_CHECK(what)                                \
    do {                                                      \
        cudaError_t _e = cudaGetLastError();                  \
        if (_e != cudaSuccess) return qdb::cuda_fail(_e, what); \
        qdb::count_launch();                                  \
    } while (0)

// ---- packed (DMMA A-fragment) operator layout -------------------------------------------------
// rows padded to npad = 8*ceil(n/8), columns (the GEMM k dimension) to kpad = 16*ceil(n/16) so that the
// k-tile count kpad/4 is a multiple of the 4-deep fragment ring of the fused steppers;
// element (r, c) lives at ((r/8)*(kpad/4) + c/4)*32 + (r%8)*4 + c%4
__host__ __device__ inline int round_up8(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int round_up16(int n) { return (n + 15) & ~15; }
constexpr int kMaxGridY = 65535;  // CUDA limit of gridDim.y / gridDim.z: launchers slice longer axes
__host__ __device__ inline size_t packed_index(int kpad, int r, int c) {
    return ((size_t)(r >> 3) * (kpad >> 2) + (c >> 2)) * 32 + ((r & 7) << 2) + (c & 3);
}

// ---- DMMA -------------------------------------------------------------------------------------
// D(8x8) += A(8x4) * B(4x8); lane = 4*g + q holds A[g][q], B[q][g], C[g][2q], C[g][2q+1].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// sign flip on the integer pipe (keeps the fp64 pipe free for DMMA)
__device__ __forceinline__ double negate(double x) {
    return __longlong_as_double(__double_as_longlong(x) ^ (long long)0x8000000000000000ULL);
}

// complex helpers on double2
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmul_conj_a(double2 a, double2 b) {  // conj(a) * b
    return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// p = exp(-i * mu * t)
__device__ __forceinline__ double2 frame_phase(double mu, double t) {
    double s, c;
    sincos(mu * t, &s, &c);
    return make_double2(c, -s);
}

// 16-byte streaming global load (read once per CTA per stage: do not allocate in L1)
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// cp.async 16 B with zero fill when !pred
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int bytes = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

}  // namespace qdb

// ---- internal launchers shared between translation units --------------------------------------
namespace qdb {
int launch_pack(int n, int count, const double2* src, double2* dst, cudaStream_t st);
int launch_generator(int n, int K, int T, int layout, const double2* ops, const double2* stat,
                     const double* coeff, int coeff_complex, const double* mu, const double* times,
                     double t_scalar, double scale, double2* out, cudaStream_t st);
int launch_frame_apply(int n, int B, const double* mu, double t, int conj_phase, const double2* y_in,
                       double2* y_out, int ldy, cudaStream_t st);
int launch_phase_vectors(int n, const double* mu, double t, double2* pre, double2* post, cudaStream_t st);
int launch_poly(int n, double c0, double c1, const double2* A1, double c2, const double2* A2, double c3,
                const double2* A3, double c4, const double2* A4, double2* out, cudaStream_t st);
int expm_core(int n, const double2* As, int squarings, double2* out, double2* ws /*5 n^2*/, cudaStream_t st);
int magnus_terms(int n, int order, const double2* g /*[order][n][n]*/, double h, double scale, double2* out,
                 double2* ws /*order 2: n^2, order 3: 7 n^2*/, cudaStream_t st);
int launch_zgemm(int M, int N, int Kd, const double2* A, int lda, const double2* B, int ldb,
                 double2* C, int ldc, double2 alpha, double2 beta, const double* colscale,
                 const double2* pre, const double2* post, cudaStream_t st);
int launch_zgemm_batched(int M, int N, int Kd, const double2* A, int lda, long long sA, const double2* B, int ldb,
                         long long sB, double2* C, int ldc, long long sC, double2 alpha, double2 beta, int count,
                         cudaStream_t st);
// two-output RK4 stage epilogue variant (generic large-n path):
//   k = G * yin ;  yout = ybase + a_next * k ;  acc = (first ? 0 : acc) + w * k
int launch_zgemm_rk4stage(int n, int B, const double2* G, const double2* yin, int ldy,
                          const double2* ybase, double2* yout, double2* acc, double a_next, double w,
                          int first, cudaStream_t st);
int launch_dmma_probe(double* sink, int iters, int* grid_out, cudaStream_t st);
// time-parallel solvers (propagator.cu)
size_t propagator_workspace_bytes(int n, int S);
int magnus_terms_batched(int n, int order, int count, const double2* g /*[count][order][n^2]*/, double h, double scale,
                         double2* out /*[count][n^2]*/, double2* ws /*order 2: count n^2, order 3: 7 count n^2*/, cudaStream_t st);
int expm_core_batched(int n, int count, const double2* As, int squarings, double2* out, double2* ws /*5 count n^2*/, cudaStream_t st);
int step_propagator_product(int n, int K, int S, int kind, const double2* ops_rm, const double2* stat_rm, const double* coeff,
                            const double* mu, const double* times_host, const int* squarings_host, double h, double2* P_total,
                            void* workspace, size_t ws_bytes, cudaStream_t st);
bool rk4_fused_supported(int n);
bool rk4_fused_tiling(int n, int B, int sweep_K, int* out);
int launch_rk4_fused_shared(int n, int B, int S, const double2* gen_table, int table_layout, double h,
                            double2* y, int ldy, cudaStream_t st);
int rk4_fused_table_layout(int n, int B);
// complex GEMM emulated on the int8 tensor cores (zgemm_ozaki.cu)
bool zgemm_int8_preferred(int M, int N, int Kd);
int launch_zgemm_int8(int M, int N, int Kd, const double2* A, int lda, const double2* B, int ldb, double2* C, int ldc, double2 alpha,
                      double2 beta, const double* colscale, const double2* pre, const double2* post, cudaStream_t st);
int launch_zgemm_int8_rk4stage(int n, int B, const double2* G, const double2* yin, int ldy, const double2* ybase, double2* yout, double2* acc,
                               double a_next, double w, int first, cudaStream_t st);
int launch_zgemm_int8_batched(int M, int N, int Kd, const double2* A, int lda, long long sA, const double2* B, int ldb, long long sB,
                              double2* C, int ldc, long long sC, double2 alpha, double2 beta, int count, cudaStream_t st);
// fp64 emulation on the int8 tensor cores (rk4_ozaki.cu), n = 121..128
bool rk4_ozaki_supported(int n);
size_t rk4_ozaki_table_bytes(int T);
void rk4_ozaki_debug(long long* host64);
bool rk4_ozaki_preferred(int n, int B);
int launch_ozaki_slice(int n, int T, const double2* gen, int gen_layout, void* ws, cudaStream_t st);
int launch_rk4_ozaki(int n, int B, int S, const double2* gen, int gen_layout, double h, double2* y, int ldy, void* ws, cudaStream_t st);
int launch_rk4_rowsplit3m(int n, int B, int S, const double2* gen_table, double h, double2* y, int ldy, cudaStream_t st);
int launch_rk4_fused_sweep(int n, int K, int B, int S, const double2* stat_packed /*or null*/,
                           const double2* ops_packed /*[K]*/, const double* coeff, int ldc,
                           const double* mu, const double* times_dev,
                           double h, double2* y, int ldy, void* ws /*rk4_sweepf_workspace_bytes*/, cudaStream_t st);
// formed-generator sweep kernel (rk4_sweepf.cu): 1 <= K <= 16; chosen automatically for K >= 3, and for K <= 2 on large batches
bool rk4_sweepf_supported(int n, int K);
bool rk4_sweepf_selected(int n, int K, int B, bool small_kernel_available);
size_t rk4_sweepf_workspace_bytes(int n, int K);
bool rk4_sweepf_tiling(int n, int B, int K, int* out);
int launch_rk4_sweepf(int n, int K, int B, int S, const double2* stat_packed, const double2* ops_packed, const double* coeff,
                      int ldc, const double* mu, const double* times_dev, double h, double2* y, int ldy, void* ws, cudaStream_t st);
int launch_signal_table(int T, int K, int B, int nterms, const int* chan, const long long* samp_off,
                        const int* samp_len, const double* dt, const double* t0, const double* freq,
                        const double* phase, int params_per_col, const double2* samples, long long col_stride,
                        const double2* scale, const double* times, double t_scalar, double* out, cudaStream_t st);
int launch_outcome_probabilities(int n, int B, int n_out, const double2* y, int ldy, const int* outcome_of,
                                 int normalize, double* out, cudaStream_t st);
bool rk4_sweep_small_supported(int n, int K, bool has_static);
int launch_rk4_sweep_small(int n, int K, int B, int S, const double2* stat, const double2* ops, const double* coeff,
                           int ldc, const double* mu, const double* times, double h, double2* y, int ldy,
                           cudaStream_t st);
// non-vectorised Lindblad on (B, n, n) batches (lindblad.cu), n <= 32
bool lindblad_fused_supported(int n);
int launch_lindblad_rhs(int n, int J, int B, const double2* m1, const double2* m2t, const double2* diss, const double* gam,
                        const double* mu, double t, const double2* rho_in, double2* rho_out, cudaStream_t st);
int launch_lindblad_rk4(int n, int J, int B, int S, const double2* m1_table, const double2* m2t_table, const double2* diss,
                        const double* gam_table, const double* mu, const double* times_dev, double h, double2* rho,
                        cudaStream_t st);
int launch_rk4_combine(size_t count, const double2* ybase, const double2* k, double2* yout, double2* acc, double a_next,
                       double w, int first, cudaStream_t st);
int launch_axpby(size_t count, double2* dst, const double2* x, double a, const double2* y, double b,
                 cudaStream_t st);
}  // namespace qdb
