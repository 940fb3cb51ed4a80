// Fused shared-signal RK4 with the fp64 contraction EMULATED on the int8 tensor cores (tcgen05.mma kind::i8, accumulators
// and the generator operand in TMEM) -- an Ozaki-style error-free split.  n = 121..128.
//
// Why.  tcgen05 has no fp64 kind; DMMA is the fp64 tensor pipe of sm_100a and rk4_shared3m_kernel already runs it at 85 %.
// The only way past that roof is to leave the fp64 pipe: every real operand is cut into NS = 6 signed 7-bit slices against a
// per-row (generator) / per-column (stage vector) power-of-two scale,
//     x = 2^e (q_1 2^-7 + q_2 2^-14 + ... + q_6 2^-42) + O(2^(e-43)),     q_p integers, |q_1| <= 127, |q_p| <= 64,
// so that the products of slices are EXACT in int32 (K = 128: |sum| < 2^25) and G y = sum over slice pairs.  Pairs of equal
// weight 2^(-7 (p + q)) share an accumulator ("group" g = p + q); pairs with g > NS + 1 lie below the truncation error of the
// operands and are dropped: 21 pairs x 4 real products (re = Ar Br + Ai (-Bi), im = Ar Bi + Ai Br).  Error per RHS
// evaluation: normwise 2^-42 per operand (measured against the DMMA kernel in tests/test_ozaki_gpu.py).
//
// Hardware mapping (measured first: profiles/probe/umma_i8_probe.cu -> profiles/r02_m_umma_i8_probe.jsonl).  With both
// operands in shared memory an M128 x N x K32 int8 MMA costs (4096 + 32 N) / 128 cycles -- the operand READ, 41 cycles at
// N = 32 -- so the generator slices live in TMEM (A operand from TMEM: 21 cycles at N = 32, 33 at N = 64 = the peak of
// 8192 MAC/clk/SM).  A CTA owns 32 whole columns for the launch (4096 columns = 128 CTAs):
//   TMEM (512 columns): [0, 128) four int32 accumulators (re, im) x two groups in flight; [128, 512) the 12 generator
//     slice planes (2 parts x 6 slices x 32 columns of packed int8), reloaded when the stage time changes (every other stage);
//   shared memory: the 18 stage-vector slice planes (re, im, -im) in the no-swizzle K-major core-matrix layout the MMA
//     reads (72 KB), y and the RK4 k-sum as fp64 (128 KB);
//   warps 0-15: epilogue -- drain a group (tcgen05.ld), combine the groups in int64, one conversion to fp64, RK4 combine, column
//     scales (warp REDUX + one named barrier), re-slice the next stage vector (integer digits) into shared memory; warp 16:
//     issues the MMAs of a stage group by group; warps 17-24: load the next generator entry's slices from L2 and tcgen05.st
//     them into TMEM while the epilogue runs.
//   Pipelines: full / empty mbarriers per accumulator buffer (MMA <-> epilogue), b_ready (stage vector sliced), a_ready /
//     a_free (generator slices in TMEM).
#include <cstdint>

#include "qdb_common.cuh"
#include "rk4_device.cuh"

namespace qdb {

__device__ long long g_oz_dbg[64];

namespace {

#define OZ_DBG(i) do { if (blockIdx.x == 0 && sidx == 8 && (threadIdx.x & 31) == 0) g_oz_dbg[i] = clock64(); } while (0)

constexpr int NS = 6;        // slices per operand
constexpr int NCOL = 32;     // columns per CTA (MMA N)
constexpr int KD = 128;      // padded dimension (MMA M and K)
constexpr int EPI_WARPS = 16, MMA_WARP = 16, LOADERS = 4, NWARPS = 20;  // warp 16: MMA issuer + loader of lane quarter 0; 17-19: loaders
constexpr uint32_t TMEM_A = 128;  // TMEM columns [128, 512): generator slice planes; [0, 128): accumulators (the allocation is the whole TMEM: base 0)
constexpr int BPLANE = NCOL * KD;  // bytes of one stage-vector slice plane

// shared-memory carve-up (bytes)
constexpr int SM_B = 0;                                   // [NS][3][BPLANE] int8
constexpr int SM_Y = SM_B + NS * 3 * BPLANE;              // [NCOL][KD] double2
constexpr int SM_K = SM_Y + NCOL * KD * 16;               // [NCOL][KD] double2
constexpr int SM_RED = SM_K + NCOL * KD * 16;             // [4][NCOL] unsigned (high words of the column maxima)
constexpr int SM_EA = SM_RED + 4 * NCOL * 8;              // [2][KD] int
constexpr int SM_BAR = SM_EA + 2 * KD * 4;                // 8 mbarriers
constexpr int SM_TMEM = SM_BAR + 8 * 8;
constexpr int SM_TOTAL = SM_TMEM + 16;

// Stage-vector slice planes (the MMA's B operand, 32 columns x 128 k int8) in the MN-major no-swizzle layout: core matrix =
// 8 k-rows of 16 consecutive columns.  A thread (one k, eight consecutive columns) owns 8 contiguous bytes of every plane:
// one 64-bit store instead of eight byte stores.
__device__ __forceinline__ int bplane_off8(int oc, int k) { return ((k >> 3) * (NCOL >> 4) + (oc >> 1)) * 128 + (k & 7) * 16 + (oc & 1) * 8; }

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46);
}
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(KD >> 4) << 24);  // s32 += s8 x s8, A K-major (TMEM), B MN-major, M128 N32

// executed by a whole warp in uniform control flow; one elected lane issues
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(kIdesc),
        "r"(accumulate), "r"(0u)
        : "memory");
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

// 2^e as a double (|e| < 1000)
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// Slice exponent from the HIGH WORD of max |x| (sign cleared): e with |x| 2^(7 - e) < 127 for every x <= that maximum, so
// that the leading slice fits a signed byte (the top six mantissa bits set: one more bit of head room); zero / denormal -> 0
__device__ __forceinline__ int slice_exponent_hi(unsigned hi) {
    if (hi < 0x00100000u) return 0;
    int e = (int)(hi >> 20) - 1022;  // 2^(e-1) <= m < 2^e
    if ((hi & 0xFFFFFu) >= 0xFC000u) ++e;
    return e;
}
__device__ __forceinline__ unsigned abs_hi(double x) { return (unsigned)(__double_as_longlong(x) >> 32) & 0x7FFFFFFFu; }

// x -> NS signed 7-bit slices against 2^e: x 2^-e = sum_p q_p 2^(-7p) + O(2^(-7 NS - 1)).  One fp64 multiply and one
// conversion (X = rint(x 2^(7 NS - e)), |X| < 2^42), then balanced base-128 digits on the INTEGER pipe -- the fp64 pipe is the
// scarce one (16 lanes per clock and sub-partition): q_NS .. q_2 in [-64, 63], q_1 in [-127, 127].
__device__ __forceinline__ void slice7(double x, double scale /* 2^(7 NS - e) */, int (&q)[NS]) {
    static_assert(NS == 6, "digit extraction is written for six slices");
    // balanced digits of X = plain base-128 digits of X + 64 (1 + 128 + ... + 128^4), minus 64 each; the leading one is the rest
    const long long Y = __double2ll_rn(x * scale) + 17315143744LL;
    const unsigned lo = (unsigned)Y, hi = (unsigned)((unsigned long long)Y >> 32);
    q[5] = (int)(lo & 127u) - 64;
    q[4] = (int)((lo >> 7) & 127u) - 64;
    q[3] = (int)((lo >> 14) & 127u) - 64;
    q[2] = (int)((lo >> 21) & 127u) - 64;
    q[1] = (int)(__funnelshift_r(lo, hi, 28) & 127u) - 64;
    q[0] = (int)hi >> 3;
}

// ---- generator table (row-major fp64, as generator_kernel writes it) -> int8 slice planes + row exponents ----
// planes[t][part][p][row][k] (128 B rows = the TMEM image of the row), expo[t][row]; one block per (row, t)
__global__ void __launch_bounds__(128) ozaki_gslice_kernel(int n, const double2* __restrict__ gen, int8_t* __restrict__ planes,
                                                           int* __restrict__ expo) {
    const int row = blockIdx.x, t = blockIdx.y, k = threadIdx.x;
    __shared__ unsigned wmax[4];
    double2 v = make_double2(0.0, 0.0);
    if (row < n && k < n) v = gen[((size_t)t * n + row) * n + k];
    unsigned m = max(abs_hi(v.x), abs_hi(v.y));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((k & 31) == 0) wmax[k >> 5] = m;
    __syncthreads();
    m = max(max(wmax[0], wmax[1]), max(wmax[2], wmax[3]));
    const int e = slice_exponent_hi(m);
    if (k == 0) expo[(size_t)t * KD + row] = e;
    const double scale = pow2(7 * NS - e);
    int qr[NS], qi[NS];
    slice7(v.x, scale, qr);
    slice7(v.y, scale, qi);
    int8_t* base = planes + (size_t)t * 2 * NS * KD * KD + (size_t)row * KD + k;
#pragma unroll
    for (int p = 0; p < NS; ++p) {
        base[(size_t)(0 * NS + p) * KD * KD] = (int8_t)qr[p];
        base[(size_t)(1 * NS + p) * KD * KD] = (int8_t)qi[p];
    }
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

__device__ __forceinline__ int stage_entry(int sidx) {
    const int step = sidx >> 2, stage = sidx & 3;
    return 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
}

__global__ void __launch_bounds__(NWARPS * 32, 1)
rk4_ozaki_kernel(int n, int B, int S, const int8_t* __restrict__ planes, const int* __restrict__ expo, double h, double2* __restrict__ y,
                 int ldy) {
    extern __shared__ __align__(1024) uint8_t sm[];
    int8_t* bsl = reinterpret_cast<int8_t*>(sm + SM_B);
    double2* ysm = reinterpret_cast<double2*>(sm + SM_Y);
    double2* ksm = reinterpret_cast<double2*>(sm + SM_K);
    unsigned* red = reinterpret_cast<unsigned*>(sm + SM_RED);
    int* ea_s = reinterpret_cast<int*>(sm + SM_EA);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BAR);
    uint64_t *full = bars, *empty = bars + 2, *b_ready = bars + 4, *a_ready = bars + 5, *a_free = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + SM_TMEM);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = blockIdx.x * NCOL;
    const int total = 4 * S;

    if (tid == 0) {
        mbar_init(full, 1);
        mbar_init(full + 1, 1);
        mbar_init(empty, EPI_WARPS);
        mbar_init(empty + 1, EPI_WARPS);
        mbar_init(b_ready, EPI_WARPS);
        mbar_init(a_ready, LOADERS);
        mbar_init(a_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0u) __trap();  // all 512 columns of the only resident CTA: the allocation starts at column 0, lane 0
    constexpr uint32_t tmem = 0u;

    if (warp < EPI_WARPS) {
        // =========================== epilogue warps: thread = (row, 8 columns) ===========================
        const int qd = warp & 3, oc = warp >> 2;
        const int row = 32 * qd + lane;
        const uint32_t lane_base = ((uint32_t)(32 * qd) << 16);
        int eb[8];  // column exponents of the current stage vector (this thread's 8 columns)

        // slices the stage vector x (this thread's 8 elements) into shared memory; returns through eb the column scales
        auto slice_stage = [&](const double2 (&x)[8]) {
            unsigned m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = __reduce_max_sync(0xffffffffu, max(abs_hi(x[j].x), abs_hi(x[j].y)));  // over the warp's 32 rows
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) red[qd * NCOL + 8 * oc + j] = m[j];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            unsigned wl[NS * 3], wh[NS * 3];  // the thread's 8 bytes of each of the 18 planes (columns 0-3 / 4-7 of its octet)
#pragma unroll
            for (int i = 0; i < NS * 3; ++i) wl[i] = wh[i] = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * oc + j;
                const unsigned mm = max(max(red[c], red[NCOL + c]), max(red[2 * NCOL + c], red[3 * NCOL + c]));
                eb[j] = slice_exponent_hi(mm);
                const double scale = pow2(7 * NS - eb[j]);
                int qr[NS], qi[NS];
                slice7(x[j].x, scale, qr);
                slice7(x[j].y, scale, qi);
#pragma unroll
                for (int p = 0; p < NS; ++p) {
                    unsigned& w0 = j < 4 ? wl[p * 3 + 0] : wh[p * 3 + 0];
                    unsigned& w1 = j < 4 ? wl[p * 3 + 1] : wh[p * 3 + 1];
                    unsigned& w2 = j < 4 ? wl[p * 3 + 2] : wh[p * 3 + 2];
                    asm("bfi.b32 %0, %1, %0, %2, 8;" : "+r"(w0) : "r"(qr[p]), "r"(8 * (j & 3)));
                    asm("bfi.b32 %0, %1, %0, %2, 8;" : "+r"(w1) : "r"(qi[p]), "r"(8 * (j & 3)));
                    asm("bfi.b32 %0, %1, %0, %2, 8;" : "+r"(w2) : "r"(-qi[p]), "r"(8 * (j & 3)));
                }
            }
            {
                const int off = bplane_off8(oc, row);
#pragma unroll
                for (int i = 0; i < NS * 3; ++i) *reinterpret_cast<uint2*>(bsl + i * BPLANE + off) = make_uint2(wl[i], wh[i]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(b_ready);
        };

        double2 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * oc + j, col = col0 + c;
            double2 v = make_double2(0.0, 0.0);
            if (row < n && col < B) v = y[(size_t)row * ldy + col];
            x[j] = v;
            ysm[c * KD + row] = v;
            ksm[c * KD + row] = make_double2(0.0, 0.0);
        }
        slice_stage(x);

        unsigned pf[2] = {0u, 0u};
#pragma unroll 1
        for (int sidx = 0; sidx < total; ++sidx) {
            const int stage = sidx & 3, entry = stage_entry(sidx);
            // groups g = 2 .. NS + 1 arrive in order of decreasing weight 2^(-7g): T = sum_g D_g 128^(NS + 1 - g) by Horner in
            // int64 (|D_g| < 2^25, |T| < 2^61: exact), ONE conversion to fp64 per value at the end
            long long tr[8], ti[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tr[j] = ti[j] = 0;
            if (warp == 0) OZ_DBG(8);
#pragma unroll 1
            for (int g = 2; g <= NS + 1; ++g) {
                const int b = g & 1;
                mbar_wait(full + b, pf[b]);
                pf[b] ^= 1u;
                tc_fence_after();
                int vr[8], vi[8];
                tmem_ld8(lane_base + (uint32_t)((2 * b) * NCOL + 8 * oc), vr);
                tmem_ld8(lane_base + (uint32_t)((2 * b + 1) * NCOL + 8 * oc), vi);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                if (lane == 0) mbar_arrive(empty + b);
                if (warp == 0) OZ_DBG(8 + g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    tr[j] = (tr[j] << 7) + vr[j];
                    ti[j] = (ti[j] << 7) + vi[j];
                }
            }
            // k = 2^(eA[row] + eB[col]) acc; RK4 combine; next stage input
            const int ea = ea_s[(entry & 1) * KD + row];
            const StageCoef sc(stage, h);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * oc + j;
                const double s = pow2(ea + eb[j] - 7 * (NS + 1));
                const double k_r = (double)tr[j] * s, k_i = (double)ti[j] * s;
                double2 ks = stage == 0 ? make_double2(0.0, 0.0) : ksm[c * KD + row];
                ks.x = sc.keep * ks.x + sc.wk * k_r;
                ks.y = sc.keep * ks.y + sc.wk * k_i;
                const double v_r = sc.last ? ks.x : k_r, v_i = sc.last ? ks.y : k_i;
                const double2 yv = ysm[c * KD + row];
                x[j] = make_double2(yv.x + sc.astep * v_r, yv.y + sc.astep * v_i);
                if (sc.last) ysm[c * KD + row] = x[j]; else ksm[c * KD + row] = ks;
            }
            if (warp == 0) OZ_DBG(20);
            if (sidx + 1 < total) slice_stage(x);
            if (warp == 0) OZ_DBG(21);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * oc + j, col = col0 + c;
            if (row < n && col < B) y[(size_t)row * ldy + col] = ysm[c * KD + row];
        }
    } else {
        // ============ warp 16: MMA issuer (+ generator loader of lane quarter 0); warps 17-19: loaders of quarters 1-3 ============
        // thread = row of the generator: its 12 slice planes (128 B each) go global -> registers -> TMEM (tcgen05.st), one plane
        // ahead; the first plane of the next entry is requested before the previous entry is released
        const int qd = warp & 3;  // 16 -> 0, 17 -> 1, 18 -> 2, 19 -> 3
        const int row = 32 * qd + lane;
        const uint32_t a_lane_base = TMEM_A + ((uint32_t)(32 * qd) << 16);
        auto load_entry = [&](int e) {
            const int8_t* src = planes + (size_t)e * 2 * NS * KD * KD + (size_t)row * KD;
            uint4 w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
            ea_s[(e & 1) * KD + row] = expo[(size_t)e * KD + row];
#pragma unroll 1
            for (int pl = 0; pl < 2 * NS; ++pl) {
                uint4 wn[8];
                const int8_t* nxt = src + (size_t)(pl + 1 < 2 * NS ? pl + 1 : pl) * KD * KD;
#pragma unroll
                for (int i = 0; i < 8; ++i) wn[i] = __ldg(reinterpret_cast<const uint4*>(nxt) + i);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t v[8] = {w[2 * i].x, w[2 * i].y, w[2 * i].z, w[2 * i].w, w[2 * i + 1].x, w[2 * i + 1].y, w[2 * i + 1].z, w[2 * i + 1].w};
                    tmem_st8(a_lane_base + (uint32_t)(pl * 32 + 8 * i), v);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = wn[i];
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        };
        const int last_entry = 2 * S;
        if (warp != MMA_WARP) {
            unsigned pfree = 0u;
#pragma unroll 1
            for (int e = 0; e <= last_entry; ++e) {
                if (e > 0) {
                    mbar_wait(a_free, pfree);
                    pfree ^= 1u;
                    tc_fence_after();
                }
                load_entry(e);
            }
        } else {
            const uint32_t bs_addr = (uint32_t)__cvta_generic_to_shared(bsl);
            constexpr uint32_t LBO = (NCOL / 16) * 128, SBO = 128;  // between the 8-row k groups / between the 16-column cores
            const uint64_t bdesc0 = smem_desc(bs_addr, LBO, SBO);
            unsigned pe[2] = {1u, 1u}, pb = 0u, pa = 0u, pfree = 0u;
            load_entry(0);
#pragma unroll 1
            for (int sidx = 0; sidx < total; ++sidx) {
                const int entry = stage_entry(sidx);
                OZ_DBG(0);
                if (sidx == 0 || stage_entry(sidx - 1) != entry) {
                    mbar_wait(a_ready, pa);
                    pa ^= 1u;
                }
                OZ_DBG(1);
                mbar_wait(b_ready, pb);
                pb ^= 1u;
                tc_fence_after();
                OZ_DBG(2);
                // fully unrolled: every TMEM / shared-memory operand is a constant or a base plus a compile-time offset (the
                // descriptor's address field is bits [0, 14) of its low word in 16 B units: an offset never carries out of it)
#pragma unroll
                for (int g = 2; g <= NS + 1; ++g) {
                    const int b = g & 1;
                    mbar_wait(empty + b, pe[b]);
                    pe[b] ^= 1u;
                    tc_fence_after();
                    constexpr uint32_t dummy = 0;
                    (void)dummy;
                    const uint32_t d_re = (uint32_t)((2 * b) * NCOL), d_im = (uint32_t)((2 * b + 1) * NCOL);
#pragma unroll
                    for (int p = 1; p < g; ++p) {
                        const int q = g - p;
                        const uint32_t a_re = TMEM_A + (uint32_t)((0 * NS + (p - 1)) * 32), a_im = TMEM_A + (uint32_t)((1 * NS + (p - 1)) * 32);
#pragma unroll
                        for (int ks = 0; ks < KD / 32; ++ks) {
                            const uint64_t dre = bdesc0 + (uint64_t)((((q - 1) * 3 + 0) * BPLANE + ks * 4 * (int)LBO) >> 4);
                            const uint64_t dim = bdesc0 + (uint64_t)((((q - 1) * 3 + 1) * BPLANE + ks * 4 * (int)LBO) >> 4);
                            const uint64_t dnm = bdesc0 + (uint64_t)((((q - 1) * 3 + 2) * BPLANE + ks * 4 * (int)LBO) >> 4);
                            const uint32_t acc = (p == 1 && ks == 0) ? 0u : 1u;
                            mma_ts(d_re, a_re + 8 * ks, dre, acc);
                            mma_ts(d_re, a_im + 8 * ks, dnm, 1u);
                            mma_ts(d_im, a_re + 8 * ks, dim, acc);
                            mma_ts(d_im, a_im + 8 * ks, dre, 1u);
                        }
                    }
                    umma_commit(full + b);
                }
                OZ_DBG(3);
                if (sidx + 1 < total && stage_entry(sidx + 1) != entry) {
                    umma_commit(a_free);  // completes when every MMA that reads this entry is done
                    mbar_wait(a_free, pfree);
                    pfree ^= 1u;
                    tc_fence_after();
                    load_entry(stage_entry(sidx + 1));  // this warp's quarter of the next entry
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace

bool rk4_ozaki_supported(int n) { return n >= 121 && n <= 128; }

void rk4_ozaki_debug(long long* host64) { cudaMemcpyFromSymbol(host64, g_oz_dbg, sizeof(long long) * 64); }

// bytes of the int8 slice planes + row exponents of T table entries
size_t rk4_ozaki_table_bytes(int T) { return (size_t)T * (2 * NS * KD * KD + KD * sizeof(int)); }

// gen_rowmajor: [2S+1][n][n] generator table (QDB_LAYOUT_ROWMAJOR); ws: rk4_ozaki_table_bytes(2S+1) of scratch
int launch_rk4_ozaki(int n, int B, int S, const double2* gen_rowmajor, double h, double2* y, int ldy, void* ws, cudaStream_t st) {
    const int T = 2 * S + 1;
    int8_t* planes = reinterpret_cast<int8_t*>(ws);
    int* expo = reinterpret_cast<int*>(planes + (size_t)T * 2 * NS * KD * KD);
    for (int t0 = 0; t0 < T; t0 += kMaxGridY) {
        const int Tc = T - t0 < kMaxGridY ? T - t0 : kMaxGridY;
        ozaki_gslice_kernel<<<dim3(KD, Tc), 128, 0, st>>>(n, gen_rowmajor + (size_t)t0 * n * n, planes + (size_t)t0 * 2 * NS * KD * KD,
                                                          expo + (size_t)t0 * KD);
        QDB_LAUNCH_CHECK("ozaki_gslice_kernel");
    }
    QDB_CUDA(cudaFuncSetAttribute(rk4_ozaki_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    rk4_ozaki_kernel<<<(B + NCOL - 1) / NCOL, NWARPS * 32, SM_TOTAL, st>>>(n, B, S, planes, expo, h, y, ldy);
    QDB_LAUNCH_CHECK("rk4_ozaki_kernel");
    return QDB_OK;
}

}  // namespace qdb
