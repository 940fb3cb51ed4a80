// tcgen05.mma kind::i8 probe for the Ozaki-split fp64 emulation: one CTA per SM multiplies int8
// tiles A (128 x 128, K-major) by B (N x 128, K-major) into int32 accumulators in TMEM, (1) checks the result against
// the host, (2) times the instruction stream of one emulated RK4 stage (21 slice pairs x 4 real products x 4 k-steps)
// to get the sustained int8 MAC rate per SM at N = 32 / 64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_probe umma_i8_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int M = 128, K = 128;

// K-major, no swizzle: 8 x 16 B core matrices; element (r, k) of a ROWS x 128 int8 tile
__host__ __device__ inline int tile_off(int rows, int r, int k) { return ((k >> 4) * (rows >> 3) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 15); }

__device__ inline uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // layout_type 0 = no swizzle, base_offset 0
}
__host__ __device__ inline uint32_t idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);  // S32 acc, int8 x int8, K-major
}
// executed by a whole warp in uniform control flow; one elected lane issues
__device__ inline void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc),
        "r"(accumulate), "r"(0u));
}
// A operand from TMEM (lane = row of A, 32-bit column c = elements 4c .. 4c+3 of the row)
__device__ inline void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc),
        "r"(accumulate), "r"(0u));
}
__device__ inline void mbar_init(uint64_t* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(c)); }
__device__ inline void mbar_wait(uint64_t* b, unsigned parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(b);
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
__device__ inline void umma_commit(uint64_t* b) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) probe_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                                       int swap_lbo_sbo, int reps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sa = sm;                 // 128 x 128 int8 = 16 KB
    uint8_t* sb = sm + M * K;         // N x 128 int8
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(A)[i];
    for (int i = tid; i < N * K / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(B)[i];
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t a_lbo = (M / 8) * 128, a_sbo = 128, b_lbo = (N / 8) * 128, b_sbo = 128;
    const uint32_t idesc = idesc_i8(M, N);
    const uint32_t sa_addr = (uint32_t)__cvta_generic_to_shared(sa), sb_addr = (uint32_t)__cvta_generic_to_shared(sb);
    // leading byte offset = distance of the core matrices along K, stride byte offset = along M / N (the swapped reading faults)
    auto adesc = [&](int kstep) { return smem_desc(sa_addr + kstep * 2 * a_lbo, a_lbo, a_sbo); };
    auto bdesc = [&](int kstep) { return smem_desc(sb_addr + kstep * 2 * b_lbo, b_lbo, b_sbo); };

    // ---- (1) one product D = A B^T into TMEM columns [0, N) ----
    if (warp == 0) {
        for (int ks = 0; ks < K / 32; ++ks) mma_i8(tmem, adesc(ks), bdesc(ks), idesc, ks > 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        uint32_t v[32];
        for (int c0 = 0; c0 < N; c0 += 32) {
            const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                           "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                           "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (blockIdx.x == 0)
                for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- (1b) the same product with A in TMEM columns [256, 288): written by tcgen05.st, one row per thread ----
    {
        const uint32_t ta = tmem + 256;
        for (int c0 = 0; c0 < 32; c0 += 8) {
            uint32_t w[8];
            for (int j = 0; j < 8; ++j) {
                uint32_t x = 0;
                for (int b = 0; b < 4; ++b) x |= (uint32_t)(uint8_t)sa[tile_off(M, tid, 4 * (c0 + j) + b)] << (8 * b);
                w[j] = x;
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta + ((uint32_t)(32 * warp) << 16) + c0),
                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (warp == 0) {
            for (int ks = 0; ks < K / 32; ++ks) mma_i8_ts(tmem + 128, ta + 8 * ks, bdesc(ks), idesc, ks > 0);
            umma_commit(&bar);
        }
        mbar_wait(&bar, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[32];
        int bad = 0;
        for (int c0 = 0; c0 < N && c0 < 128; c0 += 32) {
            const uint32_t taddr = tmem + 128 + ((uint32_t)(32 * warp) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                           "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                           "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (blockIdx.x == 0)
                for (int j = 0; j < 32; ++j) bad += ((int32_t)v[j] != D[(size_t)tid * N + c0 + j]);
        }
        if (blockIdx.x == 0 && bad) atomicAdd((unsigned long long*)&cycles[1], (unsigned long long)bad);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }

    // ---- (2) the instruction stream of one emulated stage, reps times: 21 slice pairs x 4 real products x 4 k-steps ----
    long long t0 = clock64();
    unsigned parity = 0;
    for (int r = 0; r < reps; ++r) {
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (swap_lbo_sbo == 0) {  // order 0: the 4 k-steps of a product back to back (same accumulator)
                for (int pair = 0; pair < 21; ++pair)
                    for (int part = 0; part < 4; ++part) {
                        const uint32_t col = (uint32_t)((((pair % 6) * 3 + (part < 2 ? part : 2)) % (512 / N)) * N);
                        for (int ks = 0; ks < K / 32; ++ks) mma_i8(tmem + col, adesc(ks), bdesc(ks), idesc, 1);
                    }
            } else {                  // order 1: A from TMEM (TS mode)
                for (int pair = 0; pair < 21; ++pair)
                    for (int part = 0; part < 4; ++part) {
                        const uint32_t col = (uint32_t)((((pair % 2) * 2 + (part & 1)) % (128 / (N < 128 ? N : 128))) * (N < 128 ? N : 128));
                        for (int ks = 0; ks < K / 32; ++ks) mma_i8_ts(tmem + col, tmem + 256 + 8 * ks, bdesc(ks), idesc, 1);
                    }
            }
            umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
    }
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// B in MN-major (N contiguous) no-swizzle layout: core matrix = 8 k-rows x 16 bytes of n; element (n, k) at
// ((k / 8) * (N / 16) + n / 16) * 128 + (k % 8) * 16 + n % 16.  Tries both readings of the descriptor's two strides.
template <int N>
__global__ void __launch_bounds__(128, 1) mn_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ Bmn, int32_t* __restrict__ D, int variant) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sa = sm;
    uint8_t* sb = sm + M * K;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * K / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(A)[i];
    for (int i = tid; i < N * K / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(Bmn)[i];
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = idesc_i8(M, N) | (1u << 16);  // B MN-major
    const uint32_t sa_addr = (uint32_t)__cvta_generic_to_shared(sa), sb_addr = (uint32_t)__cvta_generic_to_shared(sb);
    const uint32_t a_lbo = (M / 8) * 128, a_sbo = 128;
    const uint32_t n_stride = 128, k_stride = (N / 16) * 128;  // between core matrices along n / along k (8 k-rows each)
    if (warp == 0) {
        for (int ks = 0; ks < K / 32; ++ks) {
            const uint64_t db = variant == 0 ? smem_desc(sb_addr + ks * 4 * k_stride, n_stride, k_stride) : smem_desc(sb_addr + ks * 4 * k_stride, k_stride, n_stride);
            mma_i8(tmem, smem_desc(sa_addr + ks * 2 * a_lbo, a_lbo, a_sbo), db, idesc, ks > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    for (int c0 = 0; c0 < N; c0 += 32) {
        const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                       "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                       "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int N>
void run_mn(int variant) {
    std::vector<int8_t> A(M * K), B(N * K), At(M * K), Bt(N * K);
    srand(11 + N);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) { A[r * K + k] = (int8_t)(rand() % 255 - 127); At[tile_off(M, r, k)] = A[r * K + k]; }
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) { B[r * K + k] = (int8_t)(rand() % 255 - 127); Bt[((k >> 3) * (N / 16) + (r >> 4)) * 128 + (k & 7) * 16 + (r & 15)] = B[r * K + k]; }
    int8_t *dA, *dB; int32_t* dD;
    CK(cudaMalloc(&dA, M * K)); CK(cudaMalloc(&dB, N * K)); CK(cudaMalloc(&dD, M * N * 4));
    CK(cudaMemcpy(dA, At.data(), M * K, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bt.data(), N * K, cudaMemcpyHostToDevice));
    const int smem = M * K + N * K;
    CK(cudaFuncSetAttribute(mn_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mn_kernel<N><<<1, 128, smem>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"probe\": \"umma_i8_b_mn_major\", \"N\": %d, \"variant\": %d, \"error\": \"%s\"}\n", N, variant, cudaGetErrorString(e)); exit(0); }
    std::vector<int32_t> D(M * N);
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int r = 0; r < M; ++r) for (int c = 0; c < N; ++c) {
        long long s = 0; for (int k = 0; k < K; ++k) s += (int)A[r * K + k] * (int)B[c * K + k];
        if (s != D[r * N + c]) ++bad;
    }
    printf("{\"probe\": \"umma_i8_b_mn_major\", \"N\": %d, \"variant\": %d, \"mismatches\": %lld}\n", N, variant, bad);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

template <int N>
void run(int swap) {
    std::vector<int8_t> A(M * K), B(N * K), At(M * K), Bt(N * K);
    srand(7 + N);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) { A[r * K + k] = (int8_t)(rand() % 255 - 127); At[tile_off(M, r, k)] = A[r * K + k]; }
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) { B[r * K + k] = (int8_t)(rand() % 255 - 127); Bt[tile_off(N, r, k)] = B[r * K + k]; }
    int8_t *dA, *dB; int32_t* dD; long long* dc;
    CK(cudaMalloc(&dA, M * K)); CK(cudaMalloc(&dB, N * K)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dc, 16)); CK(cudaMemset(dc, 0, 16));
    CK(cudaMemcpy(dA, At.data(), M * K, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bt.data(), N * K, cudaMemcpyHostToDevice));
    const int smem = M * K + N * K;
    CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int reps = 200;
    probe_kernel<N><<<148, 128, smem>>>(dA, dB, dD, swap, reps, dc);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> D(M * N); long long cyc, ts_bad;
    CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&ts_bad, dc + 1, 8, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int r = 0; r < M; ++r) for (int c = 0; c < N; ++c) {
        long long s = 0; for (int k = 0; k < K; ++k) s += (int)A[r * K + k] * (int)B[c * K + k];
        if (s != D[r * N + c]) ++bad;
    }
    const double macs = (double)reps * 21 * 4 * M * N * K;
    printf("{\"probe\": \"umma_i8\", \"N\": %d, \"issue_order\": %d, \"mismatches\": %lld, \"ts_mismatches\": %lld, \"cycles_per_stage\": %.1f, \"mac_per_clk_per_sm\": %.1f}\n",
           N, swap, bad, ts_bad, (double)cyc / reps, macs / (double)cyc);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc);
}

int main(int argc, char** argv) {
    if (argc > 1) { run_mn<32>(atoi(argv[1])); return 0; }
    for (int order = 0; order < 2; ++order) { run<32>(order); run<64>(order); run<128>(order); }
    return 0;
}
