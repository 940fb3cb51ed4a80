// Device helpers of the int8 tensor-core emulation (an Ozaki-style error-free split of fp64 operands into signed byte slices):
// slice exponents and digit extraction, the tcgen05 / TMEM plumbing, operand-plane layouts.  Shared by rk4_ozaki.cu (fused
// RK4, generator in TMEM) and zgemm_ozaki.cu (complex GEMM).
#pragma once
#include <cstdint>

#include "qdb_common.cuh"
#include "rk4_device.cuh"

namespace qdb {
namespace {

constexpr int NS = 5;        // byte slices per operand: 2^-40, 15 slice pairs
constexpr int OZ_MIN_N = 65; // rows and k are padded to 128 (k chunks of 32 past n are skipped): below this the DMMA kernels win
constexpr int KD = 128;      // padded dimension (MMA M and K)
constexpr int LOADERS = 4;   // generator loader warps (one per TMEM lane quarter)
constexpr int NACC = 3;           // accumulator buffers per set (re | im: 32 columns each): a stage's five groups never wait for a drain
constexpr uint32_t TMEM_A = 192;  // TMEM columns [192, 512): generator slice planes; [0, 192): accumulators (the allocation is the whole TMEM: base 0)
static_assert(TMEM_A + 2 * NS * 32 <= 512, "TMEM");
__host__ __device__ constexpr int acc_of_group(int g) { return (NS + 1 - g) % NACC; }

// Stage-vector operand planes (the MMA's B operand, 128 k x 2 CS int8) in the MN-major no-swizzle layout: core matrix = 8 k-rows
// of 16 consecutive columns.  One N = 2 CS MMA computes both accumulators of a column set: (re | im) += A_re x (B_re | B_im)
// and += A_im x (-B_im | B_re) -- half the MMA count, and at CS = 32 the peak rate of the TMEM-operand path.  The two planes
// share B_re: per slice and k group (8 k-rows) the image holds 3 CS / 16 cores [-im | re | im]; (re | im) starts CS / 16 cores
// in, (-im | re) at the start, both with SBO = 128 B between cores and LBO = the k-group size.  3/4 of the bytes and of the
// epilogue's stores of two separate planes.  A thread (one k, eight consecutive columns of one part) owns 8 contiguous bytes.
template <int CS>
struct BImage {
    static constexpr int NC = CS / 16;             // cores per part
    static constexpr int KG = 3 * NC * 128;        // bytes of a k group = LBO
    static constexpr int SLICE = (KD / 8) * KG;    // bytes of a slice
    static constexpr int RE_IM = NC * 128;         // start of the (re | im) plane; (-im | re) starts at 0
    __device__ static __forceinline__ int off8(int oc, int k, int which /* 0: -im, 1: re, 2: im */) {
        return (k >> 3) * KG + (which * NC + (oc >> 1)) * 128 + (k & 7) * 16 + (oc & 1) * 8;
    }
};

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
// s32 += s8 x s8, A K-major (TMEM), B MN-major, M = 128, N = ncols
__host__ __device__ constexpr uint32_t idesc_for(int ncols) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(ncols >> 3) << 17) | ((uint32_t)(KD >> 4) << 24);
}

// executed by a whole warp in uniform control flow; one elected lane issues
// (the shared-memory descriptor arrives as two words: only the low one -- the address field -- varies between the MMAs)
template <uint32_t IDESC>
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t db_lo, uint32_t db_hi, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %6};\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(db_lo), "r"(IDESC),
        "r"(accumulate), "r"(0u), "r"(db_hi)
        : "memory");
}
// the same issued by ONE thread (the caller has elected it: the whole issue loop runs in that thread)
template <uint32_t IDESC>
__device__ __forceinline__ void mma_ts1(uint32_t tmem_d, uint32_t tmem_a, uint32_t db_lo, uint32_t db_hi, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %6};\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(db_lo), "r"(IDESC),
        "r"(accumulate), "r"(0u), "r"(db_hi)
        : "memory");
}
__device__ __forceinline__ void umma_commit1(uint64_t* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t x;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(x));
    return x != 0u;
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 2^e as a double
__device__ __forceinline__ double pow2(int e) {
    e = max(-1022, min(1023, e));  // products of scales beyond the double range saturate instead of wrapping the exponent field
    return __longlong_as_double((long long)(1023 + e) << 52);
}

// Slice exponent from the HIGH WORD of max |x| (sign cleared): e with |x| 2^-e < 1/2 - 2^-8 for every |x| <= that maximum,
// so that the leading digit of X + bias (below) fits a signed byte; zero / denormal -> 0
__device__ __forceinline__ int slice_exponent_hi(unsigned hi) {
    if (hi < 0x00100000u) return 0;
    int e = (int)(hi >> 20) - 1021;  // 2^(e-2) <= m < 2^(e-1)
    if ((hi & 0xFFFFFu) >= 0xFC000u) ++e;
    return e;
}
__device__ __forceinline__ unsigned abs_hi(double x) { return (unsigned)(__double_as_longlong(x) >> 32) & 0x7FFFFFFFu; }

// x -> NS signed byte slices against 2^e: x 2^-e = sum_p q_p 2^(-8p) + O(2^(-8 NS - 1)).  With X = rint(x 2^(8 NS - e)),
// |X| < 2^(8 NS - 1) - 2^(8 NS - 8), the balanced base-256 digits of X are the bytes of X + bias (bias = 0x80 in each of the
// NS - 1 low bytes) minus 128 each, i.e. with the top bit flipped: byte j of (X + bias) ^ bias = q_(NS - j) as a signed
// byte, j = 0 .. NS - 1.  X comes from t = fma(x, scale, 1.5 2^52): the integer sits in the low mantissa bits of t
// (|X| < 2^51), so rounding, conversion and bias are one fp64 FMA and one 64-bit integer add -- no F2I (a quarter-rate
// instruction).
template <int NSL>
__host__ __device__ constexpr long long bias_of() {  // 0x80 in the NSL - 1 low bytes
    long long b = 0;
    for (int j = 0; j < NSL - 1; ++j) b |= 0x80LL << (8 * j);
    return b;
}
constexpr double kMagic = 6755399441055744.0;
constexpr long long kMagicBits = 0x4338000000000000LL;
template <int NSL = NS>
__device__ __forceinline__ long long digits_of(double x, double scale) {
    constexpr long long kBias = bias_of<NSL>();
    return (__double_as_longlong(fma(x, scale, kMagic)) + (kBias - kMagicBits)) ^ kBias;
}
template <int NSL = NS>
__device__ __forceinline__ long long digits_of_negated(double x, double scale) {
    constexpr long long kBias = bias_of<NSL>();
    return ((kMagicBits + kBias) - __double_as_longlong(fma(x, scale, kMagic))) ^ kBias;
}

// the same slice of four columns in one word: w[j] = {a.byte j, b.byte j, c.byte j, d.byte j} (a = lowest address)
__device__ __forceinline__ void transpose4(unsigned a, unsigned b, unsigned c, unsigned d, unsigned (&w)[4]) {
    const unsigned t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(a, b, 0x7362);
    const unsigned t2 = __byte_perm(c, d, 0x5140), t3 = __byte_perm(c, d, 0x7362);
    w[0] = __byte_perm(t0, t2, 0x5410);
    w[1] = __byte_perm(t0, t2, 0x7632);
    w[2] = __byte_perm(t1, t3, 0x5410);
    w[3] = __byte_perm(t1, t3, 0x7632);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

}  // namespace
}  // namespace qdb
