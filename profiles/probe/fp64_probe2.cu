// Probe 2: (a) DMMA with distinct operand registers (realistic 2x4 tile pattern), (b) DMMA and DFMA
// issued concurrently from different warps / the same warp: are the tensor-DMMA and fp64-FMA pipes independent?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: all warps DMMA (2x4 complex tile pattern, distinct operands)
// mode 1: even warps DMMA, odd warps DFMA
// mode 2: all warps DFMA
// mode 3: every warp interleaves 32 DMMA with NF DFMA
template <int NF>
__global__ void __launch_bounds__(256) mix_kernel(double* out, int iters, int mode, double s) {
    double cr[2][4][2], ci[2][4][2];
    double ar[2], ai[2], br[4], bi[4];
    double f[16];
#pragma unroll
    for (int m = 0; m < 2; ++m) { ar[m] = s + m; ai[m] = s - m;
#pragma unroll
        for (int c = 0; c < 4; ++c) cr[m][c][0] = cr[m][c][1] = ci[m][c][0] = ci[m][c][1] = 0.0; }
#pragma unroll
    for (int c = 0; c < 4; ++c) { br[c] = s * c; bi[c] = s + 2 * c; }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
    const int warp = threadIdx.x >> 5;
    const bool do_mma = mode == 0 || mode == 3 || (mode == 1 && (warp & 1) == 0);
    const bool do_fma = mode == 2 || mode == 3 || (mode == 1 && (warp & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (do_mma) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < 4; ++c) { dmma(cr[m][c][0], cr[m][c][1], ar[m], br[c]); dmma(ci[m][c][0], ci[m][c][1], ar[m], bi[c]); }
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < 4; ++c) { dmma(cr[m][c][0], cr[m][c][1], -ai[m], bi[c]); dmma(ci[m][c][0], ci[m][c][1], ai[m], br[c]); }
        }
        if (do_fma) {
            const int reps = (mode == 3) ? 1 : 16;
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int i = 0; i < NF; ++i) f[i % 16] = fma(f[i % 16], 1.0000001, 1e-9);
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc += cr[m][c][0] + cr[m][c][1] + ci[m][c][0] + ci[m][c][1];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += f[i];
    if (acc == 123.456) out[0] = acc;
}

template <int NF>
void run(const char* name, int mode, int grid, int iters, double* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix_kernel<NF><<<grid, 256>>>(d, 10, mode, 1.0); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); mix_kernel<NF><<<grid, 256>>>(d, iters, mode, 1.0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warps = grid * 8.0;
    double mma_w = (mode == 0 || mode == 3) ? warps : (mode == 1 ? warps / 2 : 0);
    double fma_w = (mode == 2 || mode == 3) ? warps : (mode == 1 ? warps / 2 : 0);
    double mma_fl = mma_w * iters * 32.0 * 512.0;
    double fma_fl = fma_w * iters * (mode == 3 ? NF : 16.0 * NF) * 64.0;
    printf("{\"probe\": \"%s\", \"mode\": %d, \"nf\": %d, \"ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"total_tflops\": %.2f}\n",
           name, mode, NF, best, mma_fl / best * 1e-9, fma_fl / best * 1e-9, (mma_fl + fma_fl) / best * 1e-9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* d; CK(cudaMalloc(&d, 8));
    int grid = p.multiProcessorCount * 2, iters = 4000;
    run<16>("dmma_distinct_operands", 0, grid, iters, d);
    run<16>("dfma_only", 2, grid, iters / 4, d);
    run<16>("split_warps_dmma_dfma", 1, grid, iters, d);
    run<8>("interleaved_32dmma_8dfma", 3, grid, iters, d);
    run<16>("interleaved_32dmma_16dfma", 3, grid, iters, d);
    run<32>("interleaved_32dmma_32dfma", 3, grid, iters, d);
    run<64>("interleaved_32dmma_64dfma", 3, grid, iters, d);
    run<128>("interleaved_32dmma_128dfma", 3, grid, iters, d);
    run<256>("interleaved_32dmma_256dfma", 3, grid, iters, d);
    return 0;
}
