// Device helpers shared by the fused RK4 kernels (rk4_fused.cu, rk4_sweep_small.cu).
#pragma once
#include "qdb_common.cuh"

namespace qdb {

// B-fragment-ordered stage buffer: k-tile = 2 rt + g/4, fragment lane L = g%4 + 4 cin (bits: k0 k1 c0 c1 c2).
// A 16 B shared access is served per quarter warp (8 lanes -> 8 distinct 16 B bank groups = slot mod 8):
//   fragment load : the 8 lanes vary (k0, k1, c0);  epilogue store (C-fragment order, fixed i): (k0, c1, c2).
// slot = L ^ ((L >> 2) & 6) sends (k0, k1^c1, c0^c2) to the bank bits: distinct in both cases.
__device__ __forceinline__ int frag_swizzle(int L) { return L ^ ((L >> 2) & 6); }

// RK4 stage combine shared by both kernels.  k-sum weights 1,2,2,1; next-input step h/2, h/2, h;
// final update y + ((1/6) h) * ksum  (reference: fixed_step_solvers.py:60-73).
struct StageCoef {
    bool last;
    double keep, wk, astep;
    __device__ __forceinline__ StageCoef(int stage, double h) {
        last = (stage == 3);
        keep = stage == 0 ? 0.0 : 1.0;
        wk = (stage == 1 || stage == 2) ? 2.0 : 1.0;
        astep = stage < 2 ? 0.5 * h : (stage == 2 ? h : (1.0 / 6) * h);
    }
};

// mbarrier plumbing for the per-stage DSMEM exchange: the producer's st.async carries its own completion
// (complete_tx on the CONSUMER's mbarrier), so no cluster-wide barrier or memory fence sits in the stage loop.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(a), "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared (cp.async.bulk, SASS UBLKCP), completion counted in bytes on an mbarrier.
// dst and src 16 B aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

}  // namespace qdb
