// DMMA m8n8k4 dependent-issue latency on B200 (sm_100a): one CTA per SM, W warps, NACC independent accumulator chains
// per warp; cycles per DMMA per warp from clock64.  NACC = 1 gives the latency of the accumulate dependency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_latency dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void lat_kernel(double* out, long long* cyc, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NACC>
void run(int warps, double* d, long long* dc) {
    const int iters = 4096;
    lat_kernel<NACC><<<148, warps * 32>>>(d, dc, iters, 1.0, 1.0);
    cudaDeviceSynchronize();
    lat_kernel<NACC><<<148, warps * 32>>>(d, dc, iters, 1.0, 1.0);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("{\"probe\": \"dmma_chain\", \"warps_per_sm\": %d, \"chains_per_warp\": %d, \"cycles_per_dmma_per_warp\": %.2f, "
           "\"cycles_per_chain_step\": %.2f}\n", warps, NACC, (double)c / iters / NACC, (double)c / iters);
}

int main() {
    double* d; long long* dc;
    cudaMalloc(&d, 8); cudaMalloc(&dc, 8);
    for (int warps : {1, 4, 8, 16}) {
        run<1>(warps, d, dc); run<2>(warps, d, dc); run<3>(warps, d, dc); run<4>(warps, d, dc);
        run<6>(warps, d, dc); run<8>(warps, d, dc); run<12>(warps, d, dc);
    }
    return 0;
}
