// Complex GEMM  C = alpha (A diag(pre) B) .* post / colscale  + beta C  with the fp64 contraction EMULATED on the int8 tensor cores
// (tcgen05.mma kind::i8) -- the error-free byte split of rk4_ozaki.cu (see there for the arithmetic) as a general product.
//
// Why.  zgemm3m_kernel sits at 1.07x the fp64 DMMA roof (1.06x cuBLAS ZGEMM) on the vectorised-Lindblad shapes
// (729 x 4096 x 729, 729^3); the only way further is to leave the fp64 pipe.  Every real operand is cut into NS = 5 signed bytes
// against a power-of-two scale per ROW of A and per COLUMN of B (over the whole contraction length, so that the int32 slice
// products of all k chunks add up exactly in int64); 15 slice pairs x (re | im) are kept.  Normwise error 2^-40 per operand.
//
// Three launches:
//   zg_aslice_kernel   A (M x K) -> int8 planes per (row tile of 128, k chunk of 128), chunk-major like the generator planes
//                      of rk4_ozaki.cu (a loader lane = a row reads 16 B next to its neighbours'), + row exponents;
//   zg_colmax_kernel + zg_bslice_kernel   B (K x N, rows scaled by `pre`) -> per (column tile of 32, k chunk) the
//                      shared-memory IMAGE of the operand planes (B_re | B_im), (-B_im | B_re) of every slice in the MN-major
//                      no-swizzle core-matrix layout (12 KB per slice: the two planes overlap in B_re), + column exponents;
//   zgemm_ozaki_kernel one CTA per 128 x 32 tile of C: per k chunk the A planes go L2 -> (TMA, four-deep shared-memory ring)
//                      -> registers -> TMEM (four loader warps; a plane of the previous chunk is overwritten as soon as
//                      its last group has completed), the B image arrives by ONE TMA bulk copy into a double-buffered slot, warp 16 issues 120 N = 64 MMAs, sixteen
//                      epilogue warps drain the five weight groups (tcgen05.ld) into int64 sums that run over all chunks;
//                      one conversion to fp64 and the standard epilogue at the end.
#include <cstdint>
#include <cstdlib>

#include "ozaki_device.cuh"

namespace qdb {
namespace {

constexpr int TN = 32;                 // columns of C per CTA (MMA N = 64: re | im)
// B image of a (column tile, k chunk): BImage<TN> per slice (ozaki_device.cuh): 12 KB instead of the 16 KB of two planes
using BI = BImage<TN>;
constexpr int BKG = BI::KG, BSL = BI::SLICE;
__device__ __forceinline__ int bimg_off8(int oc, int k, int which) { return BI::off8(oc, k, which); }
constexpr int APL = KD * KD;           // bytes of one A plane of a (row tile, k chunk)
constexpr int Z_EPI_WARPS = 16, Z_MMA_WARP = 16, Z_PRODUCER = 21, Z_NWARPS = 22;  // 17-20: loaders (smem -> TMEM), 21: TMA producer of the A planes

// NSL byte slices per operand: 5 -> 2^-40, 15 slice pairs, three accumulator buffers, a six-deep A ring;
// 6 -> 2^-48, 21 pairs -- TMEM (12 A planes) then leaves room for two accumulator buffers, shared memory (two 72 KB B images)
// for a five-deep A ring
template <int NSL>
struct ZCfg {
    static constexpr int NACC = NSL == 5 ? 3 : 2;
    static constexpr int ASTAGES = NSL == 5 ? 6 : 5;
    static constexpr int BCHUNK = NSL * BSL;       // bytes of the B image of a (column tile, k chunk)
    static constexpr int ACHUNK = 2 * NSL * APL;
    static constexpr uint32_t TMEM_A = NACC * 2 * TN;  // accumulators below, A planes above
    static_assert(TMEM_A + 2 * NSL * 32 <= 512, "TMEM");
    static constexpr int S_B = 0, S_A = 2 * BCHUNK, S_BAR = S_A + ASTAGES * APL, S_TMEM = S_BAR + 32 * 8, S_TOTAL = S_TMEM + 16;
    static_assert(S_TOTAL <= 232448, "shared memory");
    __host__ __device__ static constexpr int acc(int g) { return (NSL + 1 - g) % NACC; }
};

struct EpiZ {
    double2* C;
    int ldc;
    double2 alpha, beta;
    const double* colscale;
    const double2* post;
    // RK4 stage of the generic large-n stepper (zgemm.cu EpiRk4): k = A B;  yout = ybase + a_next k;  acc = (first ? 0 : acc) + w k
    const double2* ybase;
    double2* yout;  // non-null selects this epilogue
    double2* acc;
    double a_next, w;
    int first;
};

// independent products side by side (grid.z / grid.y of the slicing kernels): element strides of the operands (0 = shared)
// and byte / element strides of the sliced scratch
struct ZBatch {
    long long sA, sB, sC;
    size_t a_planes, b_images;  // bytes per product
    int ea, eb;                 // exponents per product
};

// ---- A -> planes[rt][kc][part][p][k/16][row][k%16], expo[row]; one block per (padded) row ----
template <int NSL>
__global__ void __launch_bounds__(128) zg_aslice_kernel(int M, int K, int KC, const double2* __restrict__ A, int lda,
                                                         int8_t* __restrict__ planes, int* __restrict__ expo, ZBatch zb) {
    const int row = blockIdx.x, tid = threadIdx.x;
    A += (size_t)blockIdx.y * zb.sA;
    planes += (size_t)blockIdx.y * zb.a_planes;
    expo += (size_t)blockIdx.y * zb.ea;
    __shared__ unsigned wmax[4];
    unsigned m = 0u;
    if (row < M)
        for (int k = tid; k < K; k += 128) {
            const double2 v = A[(size_t)row * lda + k];
            m = max(m, max(abs_hi(v.x), abs_hi(v.y)));
        }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((tid & 31) == 0) wmax[tid >> 5] = m;
    __syncthreads();
    m = max(max(wmax[0], wmax[1]), max(wmax[2], wmax[3]));
    const int e = slice_exponent_hi(m);
    if (tid == 0) expo[row] = e;
    const double scale = pow2(8 * NSL - e);
    const int rt = row >> 7, r = row & 127;
    for (int kc = 0; kc < KC; ++kc) {
        const int k = kc * KD + tid;
        double2 v = make_double2(0.0, 0.0);
        if (row < M && k < K) v = A[(size_t)row * lda + k];
        const long long dr = digits_of<NSL>(v.x, scale), di = digits_of<NSL>(v.y, scale);
        int8_t* base = planes + (size_t)(rt * KC + kc) * ZCfg<NSL>::ACHUNK + (size_t)(tid >> 4) * (KD * 16) + r * 16 + (tid & 15);
#pragma unroll
        for (int p = 0; p < NSL; ++p) {
            base[(size_t)(0 * NSL + p) * APL] = (int8_t)(dr >> (8 * (NSL - 1 - p)));
            base[(size_t)(1 * NSL + p) * APL] = (int8_t)(di >> (8 * (NSL - 1 - p)));
        }
    }
}

__device__ __forceinline__ double2 scaled_b(const double2* __restrict__ B, int ldb, const double2* __restrict__ pre, int k, int c) {
    double2 v = B[(size_t)k * ldb + c];
    if (pre) v = cmul(pre[k], v);
    return v;
}

// ---- column maxima of diag(pre) B: colmax[c] (high words), zeroed by the caller ----
__global__ void __launch_bounds__(256) zg_colmax_kernel(int K, int N, const double2* __restrict__ B, int ldb, const double2* __restrict__ pre,
                                                         unsigned* __restrict__ colmax, ZBatch zb) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= N) return;
    B += (size_t)blockIdx.z * zb.sB;
    colmax += (size_t)blockIdx.z * zb.eb;
    const int k0 = blockIdx.y * 64, k1 = min(K, k0 + 64);
    unsigned m = 0u;
    for (int k = k0; k < k1; ++k) {
        const double2 v = scaled_b(B, ldb, pre, k, c);
        m = max(m, max(abs_hi(v.x), abs_hi(v.y)));
    }
    atomicMax(colmax + c, m);
}

// ---- diag(pre) B -> images[ct][kc][slice] (12 KB each), expo[c].  Block = column tile x 32 k rows; a warp = one k group of 8
// rows, lane = (k row, column octet): a store instruction of the warp fills two adjacent cores (256 contiguous bytes) ----
template <int NSL>
__global__ void __launch_bounds__(128) zg_bslice_kernel(int K, int N, int KC, int CT, const double2* __restrict__ B, int ldb,
                                                         const double2* __restrict__ pre, const unsigned* __restrict__ colmax,
                                                         int8_t* __restrict__ images, int* __restrict__ expo, ZBatch zb) {
    B += (size_t)blockIdx.z * zb.sB;
    colmax += (size_t)blockIdx.z * zb.eb;
    images += (size_t)blockIdx.z * zb.b_images;
    expo += (size_t)blockIdx.z * zb.eb;
    const int ct = blockIdx.x, lane = threadIdx.x & 31;
    const int k = (blockIdx.y * 4 + (threadIdx.x >> 5)) * 8 + (lane >> 2), oc = lane & 3;
    const int c0 = ct * TN + 8 * oc;
    (void)CT;
    unsigned wl[3][NSL], wh[3][NSL];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        unsigned lo[3][4], hi[3][4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int c = c0 + 4 * hh + jj;
            double2 v = make_double2(0.0, 0.0);
            int e = 0;
            if (c < N) {
                e = slice_exponent_hi(colmax[c]);
                if (k < K) v = scaled_b(B, ldb, pre, k, c);
                if (k == 0) expo[c] = e;
            } else if (k == 0) {
                expo[c] = 0;
            }
            const double scale = pow2(8 * NSL - e);
            const long long d0 = digits_of<NSL>(v.x, scale), d1 = digits_of<NSL>(v.y, scale), d2 = digits_of_negated<NSL>(v.y, scale);
            lo[0][jj] = (unsigned)d0, hi[0][jj] = (unsigned)((unsigned long long)d0 >> 32);
            lo[1][jj] = (unsigned)d1, hi[1][jj] = (unsigned)((unsigned long long)d1 >> 32);
            lo[2][jj] = (unsigned)d2, hi[2][jj] = (unsigned)((unsigned long long)d2 >> 32);
        }
#pragma unroll
        for (int part = 0; part < 3; ++part) {
            unsigned wlo[4], whi[4];  // byte j of the digits = slice NSL - j
            transpose4(lo[part][0], lo[part][1], lo[part][2], lo[part][3], wlo);
            transpose4(hi[part][0], hi[part][1], hi[part][2], hi[part][3], whi);
#pragma unroll
            for (int p = 1; p <= NSL; ++p) {
                const int byte = NSL - p;
                const unsigned w = byte < 4 ? wlo[byte] : whi[byte - 4];
                if (hh == 0) wl[part][p - 1] = w; else wh[part][p - 1] = w;
            }
        }
    }
    const int kc = k >> 7, kk = k & 127;
    int8_t* base = images + (size_t)(ct * KC + kc) * ZCfg<NSL>::BCHUNK;
#pragma unroll
    for (int p = 0; p < NSL; ++p) {
        int8_t* sl = base + p * BSL;
        *reinterpret_cast<uint2*>(sl + bimg_off8(oc, kk, 0)) = make_uint2(wl[2][p], wh[2][p]);  // -im
        *reinterpret_cast<uint2*>(sl + bimg_off8(oc, kk, 1)) = make_uint2(wl[0][p], wh[0][p]);  // re
        *reinterpret_cast<uint2*>(sl + bimg_off8(oc, kk, 2)) = make_uint2(wl[1][p], wh[1][p]);  // im
    }
}

template <int NSL>
__global__ void __launch_bounds__(Z_NWARPS * 32, 1)
zgemm_ozaki_kernel(int M, int N, int KC, const int8_t* __restrict__ aplanes, const int* __restrict__ expoA,
                   const int8_t* __restrict__ bimages, const int* __restrict__ expoB, EpiZ epi, ZBatch zb) {
    aplanes += (size_t)blockIdx.z * zb.a_planes;
    bimages += (size_t)blockIdx.z * zb.b_images;
    expoA += (size_t)blockIdx.z * zb.ea;
    expoB += (size_t)blockIdx.z * zb.eb;
    epi.C += (size_t)blockIdx.z * zb.sC;
    using Z = ZCfg<NSL>;
    constexpr int NS = NSL, NACC = Z::NACC, ASTAGES = Z::ASTAGES, BCHUNK = Z::BCHUNK, ACHUNK = Z::ACHUNK;
    constexpr int ZS_B = Z::S_B, ZS_A = Z::S_A, ZS_BAR = Z::S_BAR, ZS_TMEM = Z::S_TMEM;
    constexpr uint32_t TMEM_A = Z::TMEM_A;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + ZS_BAR);
    uint64_t *full = bars, *empty = bars + NACC, *a_ready = bars + 2 * NACC, *p_free = a_ready + 1, *b_full = p_free + NS, *b_free = b_full + 2, *as_full = b_free + 2, *as_free = as_full + ASTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + ZS_TMEM);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ct = blockIdx.x, rt = blockIdx.y;

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NACC; ++b) {
            mbar_init(full + b, 1);
            mbar_init(empty + b, Z_EPI_WARPS);
        }
        mbar_init(a_ready, LOADERS);
#pragma unroll
        for (int p = 0; p < NS; ++p) mbar_init(p_free + p, 1);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            mbar_init(b_full + b, 1);
            mbar_init(b_free + b, 1);
        }
#pragma unroll
        for (int b = 0; b < ASTAGES; ++b) {
            mbar_init(as_full + b, 1);
            mbar_init(as_free + b, LOADERS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == Z_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0u) __trap();  // all 512 columns of the only resident CTA: the allocation starts at column 0, lane 0
    constexpr uint32_t tmem = 0u;

    if (warp < Z_EPI_WARPS) {
        // =========================== epilogue warps: thread = (row, 8 columns) ===========================
        const int qd = warp & 3, oc = warp >> 2;
        const int row = 32 * qd + lane;
        const uint32_t lane_base = ((uint32_t)(32 * qd) << 16);
        unsigned pf[NACC] = {};
        // sums over the groups AND the k chunks in int64 (integer pipe; the conversions of a per-group fp64 sum throttled the
        // fp64 pipe: 22 % of the stall samples).  Range: a chunk contributes < 2^(24 + 8 (NS - 1)); with six slices the sums are
        // kept in units of 2^8 -- the least significant group is rounded to its upper 24 bits, an error of 2^-49 of the
        // operand scales, below their 2^-48 truncation -- so that k <= 4096 stays below 2^60 for either slice count
        constexpr int DROP = NS == 6 ? 8 : 0, HALF = NS == 6 ? 128 : 0;
        long long tr[8], ti[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tr[j] = ti[j] = 0;
#pragma unroll 1
        for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
            for (int g = NS + 1; g >= 2; --g) {
                const int b = Z::acc(g), sh = 8 * (NS + 1 - g);
                mbar_wait(full + b, pf[b]);
                pf[b] ^= 1u;
                tc_fence_after();
                int vr[8], vi[8];
                tmem_ld8(lane_base + (uint32_t)(2 * b * TN + 8 * oc), vr);
                tmem_ld8(lane_base + (uint32_t)((2 * b + 1) * TN + 8 * oc), vi);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                if (lane == 0) mbar_arrive(empty + b);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (sh >= DROP) {
                        tr[j] += (long long)vr[j] << (sh >= DROP ? sh - DROP : 0);
                        ti[j] += (long long)vi[j] << (sh >= DROP ? sh - DROP : 0);
                    } else {  // round to nearest, ties up
                        tr[j] += (long long)((vr[j] + HALF) >> DROP);
                        ti[j] += (long long)((vi[j] + HALF) >> DROP);
                    }
                }
            }
        }
        const int r = rt * KD + row;
        if (r < M) {
            const int ea = expoA[r];
            const double2 po = epi.post ? epi.post[r] : make_double2(1.0, 0.0);
            const bool use_beta = epi.beta.x != 0.0 || epi.beta.y != 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = ct * TN + 8 * oc + j;
                if (c < N) {
                    const double s = pow2(ea + expoB[c] - 8 * (NS + 1) + DROP);
                    double2 v = make_double2((double)tr[j] * s, (double)ti[j] * s);
                    if (epi.yout) {
                        const size_t i = (size_t)r * epi.ldc + c;
                        const double2 yb = epi.ybase[i];
                        epi.yout[i] = make_double2(fma(epi.a_next, v.x, yb.x), fma(epi.a_next, v.y, yb.y));
                        double2 a = make_double2(epi.w * v.x, epi.w * v.y);
                        if (!epi.first) {
                            const double2 old = epi.acc[i];
                            a.x += old.x;
                            a.y += old.y;
                        }
                        epi.acc[i] = a;
                        continue;
                    }
                    if (epi.post) v = cmul(po, v);
                    double2 a = epi.alpha;
                    if (epi.colscale) {
                        const double cs = epi.colscale[c];
                        a.x *= cs;
                        a.y *= cs;
                    }
                    v = cmul(a, v);
                    double2* dst = epi.C + (size_t)r * epi.ldc + c;
                    if (use_beta) v = cadd(v, cmul(epi.beta, *dst));
                    *dst = v;
                }
            }
        }
    } else if (warp == Z_PRODUCER) {
        // ============ TMA producer: the A planes of every k chunk, in the order the loaders consume them, four in flight ============
        if (lane == 0) {
            const int8_t* abase = aplanes + (size_t)rt * KC * ACHUNK;
            const int nplanes = KC * 2 * NS;
#pragma unroll 1
            for (int t = 0; t < nplanes; ++t) {
                const int kc = t / (2 * NS), it = t % (2 * NS);
                const int p = NS - (it >> 1), part = it & 1;
                const int slot = t % ASTAGES, use = t / ASTAGES;
                if (use >= 1) mbar_wait(as_free + slot, (unsigned)(use - 1) & 1u);
                mbar_expect_tx(as_full + slot, APL);
                tma_bulk_g2s(sm + ZS_A + slot * APL, abase + (size_t)kc * ACHUNK + (size_t)(part * NS + p - 1) * APL, APL, as_full + slot);
            }
        }
    } else if (warp != Z_MMA_WARP) {
        // ============ loaders: thread = row of the A tile, shared-memory ring -> registers -> TMEM (tcgen05.st), least significant
        // slice first: slice p of the previous chunk is free once its group p + 1 has completed (p_free[p - 1]); lane 0 of the
        // first loader also feeds the B images by TMA ============
        const int qd = warp & 3;
        const int row = 32 * qd + lane;
        const uint32_t a_lane_base = TMEM_A + ((uint32_t)(32 * qd) << 16);
        const bool producer = warp == Z_MMA_WARP + 1 && lane == 0;
        const int8_t* bsrc = bimages + (size_t)ct * KC * BCHUNK;
        if (producer) {
            mbar_expect_tx(b_full, BCHUNK);
            tma_bulk_g2s(sm + ZS_B, bsrc, BCHUNK, b_full);
        }
        int t = 0;
#pragma unroll 1
        for (int kc = 0; kc < KC; ++kc) {
            const unsigned par = (unsigned)(kc - 1) & 1u;
#pragma unroll 1
            for (int it = 0; it < 2 * NS; ++it, ++t) {
                const int p = NS - (it >> 1), part = it & 1;
                const int slot = t % ASTAGES, use = t / ASTAGES;
                mbar_wait(as_full + slot, (unsigned)use & 1u);
                const uint4* src = reinterpret_cast<const uint4*>(sm + ZS_A + slot * APL) + row;
                uint4 w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = src[i * KD];
                // the slot goes back to the TMA producer only when the loads have LANDED: an arrive does not wait for loads
                // in flight (the empty asm statements consume the registers, i.e. wait on their scoreboard)
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("" ::"r"(w[i].x), "r"(w[i].y), "r"(w[i].z), "r"(w[i].w) : "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(as_free + slot);  // the plane is in registers
                if (kc > 0 && part == 0) {
                    mbar_wait(p_free + (p - 1), par);
                    tc_fence_after();
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t v[8] = {w[2 * i].x, w[2 * i].y, w[2 * i].z, w[2 * i].w, w[2 * i + 1].x, w[2 * i + 1].y, w[2 * i + 1].z, w[2 * i + 1].w};
                    tmem_st8(a_lane_base + (uint32_t)((part * NS + p - 1) * 32 + 8 * i), v);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            // the B image of the next chunk: its slot was read by chunk kc - 1, whose last group has completed by now
            if (producer && kc + 1 < KC) {
                const int buf = (kc + 1) & 1;
                if (kc >= 1) mbar_wait(b_free + buf, (unsigned)((kc - 1) >> 1) & 1u);
                mbar_expect_tx(b_full + buf, BCHUNK);
                tma_bulk_g2s(sm + ZS_B + buf * BCHUNK, bsrc + (size_t)(kc + 1) * BCHUNK, BCHUNK, b_full + buf);
            }
            __syncwarp();
        }
    } else {
        // =========================== MMA issuer ===========================
        const uint32_t bs_addr = (uint32_t)__cvta_generic_to_shared(sm + ZS_B);
        constexpr uint32_t LBO = BKG, SBO = 128;
        const uint64_t bdesc0 = smem_desc(bs_addr, LBO, SBO);
        const uint32_t bd_hi = (uint32_t)(bdesc0 >> 32);
        unsigned pe[NACC];
#pragma unroll
        for (int b = 0; b < NACC; ++b) pe[b] = 1u;
        // ONE elected thread runs the whole issue loop (waits, MMAs, commits): no election and predicate shuffling per MMA
        if (elect_one())
#pragma unroll 1
        for (int kc = 0; kc < KC; ++kc) {
            const int buf = kc & 1;
            const bool release = kc + 1 < KC;
            uint32_t bd_lo = (uint32_t)bdesc0 + (uint32_t)(buf * (BCHUNK >> 4));
            asm volatile("" : "+r"(bd_lo));  // opaque per chunk: descriptor words are base + immediate, not hoisted registers
            mbar_wait(a_ready, (unsigned)kc & 1u);
            mbar_wait(b_full + buf, (unsigned)(kc >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int g = NS + 1; g >= 2; --g) {
                const int b = Z::acc(g);
                mbar_wait(empty + b, pe[b]);
                pe[b] ^= 1u;
                tc_fence_after();
                const uint32_t d = (uint32_t)(b * 2 * TN);
#pragma unroll
                for (int p = 1; p < g; ++p) {
                    const int q = g - p;
                    const uint32_t a_re = TMEM_A + (uint32_t)((0 * NS + (p - 1)) * 32), a_im = TMEM_A + (uint32_t)((1 * NS + (p - 1)) * 32);
#pragma unroll
                    for (int ks = 0; ks < KD / 32; ++ks) {
                        const uint32_t b1 = bd_lo + (uint32_t)(((q - 1) * BSL + BI::RE_IM + ks * 4 * (int)LBO) >> 4);  // (re | im): cores 2..5
                        const uint32_t b2 = bd_lo + (uint32_t)(((q - 1) * BSL + ks * 4 * (int)LBO) >> 4);        // (-im | re): cores 0..3
                        mma_ts1<idesc_for(2 * TN)>(d, a_re + 8 * ks, b1, bd_hi, (p == 1 && ks == 0) ? 0u : 1u);
                        mma_ts1<idesc_for(2 * TN)>(d, a_im + 8 * ks, b2, bd_hi, 1u);
                    }
                }
                umma_commit1(full + b);
                if (release) umma_commit1(p_free + (g - 2));
            }
            umma_commit1(b_free + buf);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == Z_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int NSL>
int launch_int8(int M, int N, int Kd, const double2* A, int lda, long long sA, const double2* B, int ldb, long long sB, double2* C, int ldc,
                long long sC, int count, double2 alpha, double2 beta, const double* colscale, const double2* pre, const double2* post,
                cudaStream_t st, const EpiZ* rk4 = nullptr) {
    using Z = ZCfg<NSL>;
    const int RT = (M + KD - 1) / KD, CT = (N + TN - 1) / TN, KC = (Kd + KD - 1) / KD;
    const size_t a_one = (size_t)RT * KC * Z::ACHUNK, b_one = (size_t)CT * KC * Z::BCHUNK;
    const size_t a_bytes = a_one * count, b_bytes = b_one * count;
    const size_t ea_bytes = (size_t)RT * KD * sizeof(int) * count, eb_bytes = (size_t)CT * TN * sizeof(int) * count,
                 cm_bytes = (size_t)CT * TN * sizeof(unsigned) * count;
    const ZBatch zb{sA, sB, sC, a_one, b_one, RT * KD, CT * TN};
    // stream-ordered scratch from the device's default pool, which is told once to keep what it has been given
    static const bool pool_ready = [] {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        return true;
    }();
    (void)pool_ready;
    uint8_t* ws = nullptr;
    QDB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), a_bytes + b_bytes + ea_bytes + eb_bytes + cm_bytes, st));
    int8_t* aplanes = reinterpret_cast<int8_t*>(ws);
    int8_t* bimages = reinterpret_cast<int8_t*>(ws + a_bytes);
    int* expoA = reinterpret_cast<int*>(ws + a_bytes + b_bytes);
    int* expoB = reinterpret_cast<int*>(ws + a_bytes + b_bytes + ea_bytes);
    unsigned* colmax = reinterpret_cast<unsigned*>(ws + a_bytes + b_bytes + ea_bytes + eb_bytes);
    QDB_CUDA(cudaMemsetAsync(colmax, 0, cm_bytes, st));
    zg_aslice_kernel<NSL><<<dim3(RT * KD, count), 128, 0, st>>>(M, Kd, KC, A, lda, aplanes, expoA, zb);
    QDB_LAUNCH_CHECK("zg_aslice_kernel");
    zg_colmax_kernel<<<dim3((N + 255) / 256, (Kd + 63) / 64, count), 256, 0, st>>>(Kd, N, B, ldb, pre, colmax, zb);
    QDB_LAUNCH_CHECK("zg_colmax_kernel");
    zg_bslice_kernel<NSL><<<dim3(CT, KC * KD / 32, count), 128, 0, st>>>(Kd, N, KC, CT, B, ldb, pre, colmax, bimages, expoB, zb);
    QDB_LAUNCH_CHECK("zg_bslice_kernel");
    QDB_CUDA(cudaFuncSetAttribute(zgemm_ozaki_kernel<NSL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Z::S_TOTAL));
    EpiZ epi{C, ldc, alpha, beta, colscale, post, nullptr, nullptr, nullptr, 0.0, 0.0, 0};
    if (rk4) epi = *rk4;
    zgemm_ozaki_kernel<NSL><<<dim3(CT, RT, count), Z_NWARPS * 32, Z::S_TOTAL, st>>>(M, N, KC, aplanes, expoA, bimages, expoB, epi, zb);
    QDB_LAUNCH_CHECK("zgemm_ozaki_kernel");
    QDB_CUDA(cudaFreeAsync(ws, st));
    return QDB_OK;
}

}  // namespace

// The emulated product pays three passes over the operands and a serial drain of the weight groups per k chunk: it wins on
// products that fill the chip with 128 x 32 tiles and are long enough in k (measured: profiles/r02_s_zgemm_int8.jsonl).
// QDB_ZGEMM_INT8=0 keeps every product on the fp64 DMMA kernels, =2 forces the emulation (tests);
// QDB_ZGEMM_SLICES=5 selects five byte slices per operand (2^-40) instead of six (2^-48).
bool zgemm_int8_preferred(int M, int N, int Kd) {
    const char* env = getenv("QDB_ZGEMM_INT8");  // read per call: the tests switch it
    const int mode = env ? atoi(env) : 1;
    if (mode == 0) return false;
    if (Kd > 4096) return false;  // the int64 sums over the k chunks stay below 2^60
    if (mode == 2) return Kd >= 1 && M >= 1 && N >= 1;
    const long tiles = (long)((M + KD - 1) / KD) * ((N + TN - 1) / TN);
    return Kd >= 384 && tiles >= sm_count() / 2;
}

int launch_zgemm_int8(int M, int N, int Kd, const double2* A, int lda, const double2* B, int ldb, double2* C, int ldc, double2 alpha,
                      double2 beta, const double* colscale, const double2* pre, const double2* post, cudaStream_t st) {
    const char* env = getenv("QDB_ZGEMM_SLICES");
    const int slices = env && atoi(env) == 5 ? 5 : 6;
    if (slices == 5) return launch_int8<5>(M, N, Kd, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, alpha, beta, colscale, pre, post, st);
    return launch_int8<6>(M, N, Kd, A, lda, 0, B, ldb, 0, C, ldc, 0, 1, alpha, beta, colscale, pre, post, st);
}

// one RK4 stage of the generic (n > 256) stepper: k = G yin, yout = ybase + a_next k, acc = (first ? 0 : acc) + w k
int launch_zgemm_int8_rk4stage(int n, int B, const double2* G, const double2* yin, int ldy, const double2* ybase, double2* yout, double2* acc,
                               double a_next, double w, int first, cudaStream_t st) {
    const char* env = getenv("QDB_ZGEMM_SLICES");
    const int slices = env && atoi(env) == 5 ? 5 : 6;
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
    const EpiZ e{nullptr, ldy, one, zero, nullptr, nullptr, ybase, yout, acc, a_next, w, first};
    if (slices == 5) return launch_int8<5>(n, B, n, G, n, 0, yin, ldy, 0, nullptr, ldy, 0, 1, one, zero, nullptr, nullptr, nullptr, st, &e);
    return launch_int8<6>(n, B, n, G, n, 0, yin, ldy, 0, nullptr, ldy, 0, 1, one, zero, nullptr, nullptr, nullptr, st, &e);
}

// count independent products side by side (operand strides in elements, 0 = shared): one launch of each of the four kernels,
// in slices of at most 64 products (scratch: ~17 MB per 729^3 product)
int launch_zgemm_int8_batched(int M, int N, int Kd, const double2* A, int lda, long long sA, const double2* B, int ldb, long long sB,
                              double2* C, int ldc, long long sC, double2 alpha, double2 beta, int count, cudaStream_t st) {
    const char* env = getenv("QDB_ZGEMM_SLICES");
    const int slices = env && atoi(env) == 5 ? 5 : 6;
    for (int z0 = 0; z0 < count; z0 += 64) {
        const int c = count - z0 < 64 ? count - z0 : 64;
        const double2 *a = A + (size_t)z0 * sA, *b = B + (size_t)z0 * sB;
        double2* cc = C + (size_t)z0 * sC;
        const int rc = slices == 5 ? launch_int8<5>(M, N, Kd, a, lda, sA, b, ldb, sB, cc, ldc, sC, c, alpha, beta, nullptr, nullptr, nullptr, st)
                                   : launch_int8<6>(M, N, Kd, a, lda, sA, b, ldb, sB, cc, ldc, sC, c, alpha, beta, nullptr, nullptr, nullptr, st);
        if (rc != QDB_OK) return rc;
    }
    return QDB_OK;
}

}  // namespace qdb
