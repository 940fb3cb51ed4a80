// fp64 pipe probe for B200 (sm_100a): DMMA m8n8k4 vs DFMA issue throughput per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x * 1e-3 + i;
    double a = a0, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
    double* d; CK(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = p.multiProcessorCount;
    int iters = 20000;
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int threads = 256; int blocks_per_sm = warps_per_sm * 32 / threads; if (blocks_per_sm < 1) { blocks_per_sm = 1; threads = warps_per_sm * 32; }
        int grid = sms * blocks_per_sm;
        // DMMA
        dmma_kernel<16><<<grid, threads>>>(d, 100, 1.0, 1.0); CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0); dmma_kernel<16><<<grid, threads>>>(d, iters, 1.0, 1.0); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double nwarps = (double)grid * threads / 32;
        double flops = nwarps * iters * 16.0 * 512.0;  // 8*8*4*2 flops per DMMA
        printf("{\"probe\": \"dmma_m8n8k4\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps_per_sm, best, flops / best * 1e-9);
        // DFMA
        dfma_kernel<16><<<grid, threads>>>(d, 100, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
        best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0); dfma_kernel<16><<<grid, threads>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        flops = nwarps * iters * 16.0 * 64.0;  // 32 lanes * 2 flops
        printf("{\"probe\": \"dfma\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", warps_per_sm, best, flops / best * 1e-9);
    }
    // sustained DMMA for ~2 s to see power-capped rate
    {
        int grid = sms * 2, threads = 256;
        cudaEventRecord(e0);
        for (int r = 0; r < 40; ++r) dmma_kernel<16><<<grid, threads>>>(d, iters * 4, 1.0, 1.0);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 40.0 * grid * threads / 32 * iters * 4 * 16.0 * 512.0;
        printf("{\"probe\": \"dmma_sustained\", \"ms\": %.2f, \"tflops\": %.3f}\n", ms, flops / ms * 1e-9);
    }
    return 0;
}
