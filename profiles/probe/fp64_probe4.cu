// Probe 4: issue patterns of the formed-generator sweep kernel (rk4_sweepf.cu).
//   VAR 0: 24 in-place DMMAs (reference)                     VAR 1: 12 DMMAs D = A B + C (C kept) followed by 12 in-place on the results
//   VAR 2: VAR 1 + 48 DFMAs consuming the DMMA results (the kernel's column body, MR=3 NCW=2 KS=2), blocks in source order
//   VAR 3: VAR 2 with the DFMAs of tile t issued after the DMMAs of tile t+1 (software pipelined by one row tile)
//   VAR 4: VAR 2 with results consumed one whole iteration later (DFMAs read the previous iteration's G)
//   VAR 5: all 12 formation DMMAs, then the 12 accumulating ones, then all 48 DFMAs
// (VAR 1 has no consumer of its results and is partly eliminated by ptxas: ignore its number.)
// Result (profiles/r01_p_fp64_probe4.jsonl): the column body reaches 32.4 of 37.1 TFLOP/s with two warps per
// sub-partition whatever the source order (ptxas reschedules it): 87 % is the ceiling of this instruction mix at the
// occupancy the kernel's register budget allows.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_from(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
template <int VAR>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters, double s) {
    double g[12][2], gp[12][2], acc[12][4], a[12], b[4], cs[6][2], y[4][2];
#pragma unroll
    for (int i = 0; i < 12; ++i) { g[i][0] = g[i][1] = gp[i][0] = gp[i][1] = 0.0; a[i] = s + i + threadIdx.x * 1e-9; acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { b[i] = s * (i + 1); y[i][0] = s + i; y[i][1] = s - i; }
#pragma unroll
    for (int i = 0; i < 6; ++i) { cs[i][0] = cs[i][1] = s * i; }
    for (int it = 0; it < iters; ++it) {
        // operands change every iteration (one integer op each) so that nothing is loop invariant
#pragma unroll
        for (int i = 0; i < 12; ++i) a[i] = __hiloint2double(__double2hiint(a[i]) ^ (it & 1), __double2loint(a[i]));
        if (VAR == 0) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 12; ++i) dmma(g[i][0], g[i][1], a[i], b[(i + r) & 3]);
        } else if (VAR == 1 || VAR == 2 || VAR == 4) {
#pragma unroll
            for (int m = 0; m < 3; ++m) {
#pragma unroll
                for (int t = 0; t < 4; ++t) dmma_from(g[4 * m + t][0], g[4 * m + t][1], a[4 * m + t], b[t & 1], cs[2 * m + (t >> 1)][0], cs[2 * m + (t >> 1)][1]);
#pragma unroll
                for (int t = 0; t < 4; ++t) dmma(g[4 * m + t][0], g[4 * m + t][1], a[(4 * m + t + 5) % 12], b[2 + (t & 1)]);
                if (VAR == 2 || VAR == 4) {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const double gv = VAR == 2 ? g[4 * m + t][i] : gp[4 * m + t][i];
                            acc[4 * m + t][2 * i] = fma(gv, y[t][0], acc[4 * m + t][2 * i]);
                            acc[4 * m + t][2 * i + 1] = fma(gv, y[t][1], acc[4 * m + t][2 * i + 1]);
                        }
                }
            }
            if (VAR == 4) {
#pragma unroll
                for (int i = 0; i < 12; ++i) { gp[i][0] = g[i][0]; gp[i][1] = g[i][1]; }
            }
        } else if (VAR == 5) {
#pragma unroll
            for (int i = 0; i < 12; ++i) dmma_from(g[i][0], g[i][1], a[i], b[i & 1], cs[i >> 1][0], cs[i >> 1][1]);
#pragma unroll
            for (int i = 0; i < 12; ++i) dmma(g[i][0], g[i][1], a[(i + 5) % 12], b[2 + (i & 1)]);
#pragma unroll
            for (int i = 0; i < 12; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    acc[i][2 * j] = fma(g[i][j], y[i & 3][0], acc[i][2 * j]);
                    acc[i][2 * j + 1] = fma(g[i][j], y[i & 3][1], acc[i][2 * j + 1]);
                }
        } else if (VAR == 3) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                if (m < 3) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) dmma_from(g[4 * m + t][0], g[4 * m + t][1], a[4 * m + t], b[t & 1], cs[2 * m + (t >> 1)][0], cs[2 * m + (t >> 1)][1]);
#pragma unroll
                    for (int t = 0; t < 4; ++t) dmma(g[4 * m + t][0], g[4 * m + t][1], a[(4 * m + t + 5) % 12], b[2 + (t & 1)]);
                }
                if (m > 0) {
                    const int mm = m - 1;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            acc[4 * mm + t][2 * i] = fma(g[4 * mm + t][i], y[t][0], acc[4 * mm + t][2 * i]);
                            acc[4 * mm + t][2 * i + 1] = fma(g[4 * mm + t][i], y[t][1], acc[4 * mm + t][2 * i + 1]);
                        }
                }
            }
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) r += g[i][0] + g[i][1] + acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3] + gp[i][0];
    if (r == 123.456) out[0] = r;
}
template <int VAR>
void run(const char* name, int warps_per_sm, int iters, double* d, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 32 * warps_per_sm;
    k<VAR><<<sms, threads>>>(d, 10, 1.0); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<VAR><<<sms, threads>>>(d, iters, 1.0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double warps = (double)sms * warps_per_sm;
    const double dmma_fma = warps * iters * 24.0 * 256.0;
    const double dfma_fma = (VAR >= 2) ? warps * iters * 48.0 * 32.0 : 0.0;
    printf("{\"probe\": \"%s\", \"var\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"total_tflops\": %.2f}\n",
           name, VAR, warps_per_sm, best, 2 * dmma_fma / best * 1e-9, 2 * dfma_fma / best * 1e-9, 2 * (dmma_fma + dfma_fma) / best * 1e-9);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* d; CK(cudaMalloc(&d, 8));
    const int iters = 20000, sms = p.multiProcessorCount;
    for (int w : {4, 8}) {
        run<5>("column_body_all_dmma_then_all_dfma", w, iters, d, sms);
        run<0>("inplace24", w, iters, d, sms);
        run<1>("from12_inplace12", w, iters, d, sms);
        run<2>("column_body_source_order", w, iters, d, sms);
        run<3>("column_body_pipelined_by_tile", w, iters, d, sms);
        run<4>("column_body_consume_next_iteration", w, iters, d, sms);
    }
    return 0;
}
