// General complex128 GEMM on the fp64 tensor pipe (DMMA m8n8k4), with the fused pro/epilogues of
// the hot path.
//
// CTA tile 64 x BN (BN = 64 or 32), k-chunk 16, 8 warps as 4 (rows) x 2 (cols); each warp owns a 16 x BN/2
// complex tile = 2 x BN/16 DMMA tiles, i.e. 32 (16) real DMMAs per k4-step fed by 2 + 4 (2 + 2) LDS.128 -- the
// DMMA pipe (16 issue cycles per DMMA per SM sub-partition) is the bound, not shared memory.
// Operand tiles are staged with a 3-deep cp.async ring; rows are padded so that every fragment
// load is bank-conflict free (A stride = 64 mod 128 B, B stride = 32 mod 128 B).
// Two CTAs fit an SM: the narrow tile is chosen when the 64 x 64 grid would leave half of those 296 slots
// empty (the 729^3 products of the vectorised-Lindblad expm: 144 CTAs -> 276).
#include <cstdlib>

#include "qdb_common.cuh"

namespace qdb {

namespace {

constexpr int BM = 64, BK = 16, STAGES = 3;
constexpr int A_LD = BK + 4;   // complex elements per smem row of A  (320 B)
constexpr int A_TILE = BM * A_LD;
template <int BN>
struct TileB {
    static constexpr int LD = BN + 2;  // complex elements per smem row of B  (1056 B / 544 B: 32 mod 128 B)
    static constexpr int TILE = BK * LD;
    static constexpr size_t SMEM = (size_t)STAGES * (A_TILE + TILE) * sizeof(double2);
};

struct EpiStd {
    double2* C;
    int ldc;
    double2 alpha, beta;
    const double* colscale;
    const double2* post;
    long long sC;  // batch stride of C (elements); 0 for a single product
    __device__ __forceinline__ void shift(unsigned z) { C += (size_t)z * (size_t)sC; }
};

struct EpiRk4 {
    const double2* ybase;
    double2* yout;
    double2* acc;
    int ld;
    double a_next, w;
    int first;
    __device__ __forceinline__ void shift(unsigned) {}
};

// batched products (grid.z): operand strides in elements, 0 = shared by every product of the batch
struct BatchStride {
    long long sA, sB;
};

__device__ __forceinline__ void epilogue(const EpiStd& e, int r, int c, double2 v) {
    if (e.post) v = cmul(e.post[r], v);
    double2 a = e.alpha;
    if (e.colscale) {
        const double s = e.colscale[c];
        a.x *= s;
        a.y *= s;
    }
    v = cmul(a, v);
    double2* dst = e.C + (size_t)r * e.ldc + c;
    if (e.beta.x != 0.0 || e.beta.y != 0.0) v = cadd(v, cmul(e.beta, *dst));
    *dst = v;
}

__device__ __forceinline__ void epilogue(const EpiRk4& e, int r, int c, double2 k) {
    const size_t i = (size_t)r * e.ld + c;
    const double2 yb = e.ybase[i];
    e.yout[i] = make_double2(fma(e.a_next, k.x, yb.x), fma(e.a_next, k.y, yb.y));
    double2 a = make_double2(e.w * k.x, e.w * k.y);
    if (!e.first) {
        const double2 old = e.acc[i];
        a.x += old.x;
        a.y += old.y;
    }
    e.acc[i] = a;
}

template <typename Epi, int BN>
__global__ void __launch_bounds__(256) zgemm_kernel(int M, int N, int Kd, const double2* __restrict__ A0, int lda,
                                                     const double2* __restrict__ Bm0, int ldb,
                                                     const double2* __restrict__ pre, Epi epi, BatchStride bs) {
    const double2* __restrict__ A = A0 + (size_t)blockIdx.z * (size_t)bs.sA;
    const double2* __restrict__ Bm = Bm0 + (size_t)blockIdx.z * (size_t)bs.sB;
    epi.shift(blockIdx.z);
    constexpr int B_LD = TileB<BN>::LD, B_TILE = TileB<BN>::TILE;
    constexpr int NC = BN / 16;  // column tiles per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* sA = reinterpret_cast<double2*>(smem_raw);
    double2* sB = sA + STAGES * A_TILE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int numK = (Kd + BK - 1) / BK;

    auto load_stage = [&](int kt, int slot) {
        const int k0 = kt * BK;
        double2* a_dst = sA + slot * A_TILE;
        double2* b_dst = sB + slot * B_TILE;
#pragma unroll
        for (int i = 0; i < (BM * BK) / 256; ++i) {
            const int idx = tid + 256 * i;
            const int r = idx / BK, c = idx % BK;
            const bool ok = (m0 + r < M) && (k0 + c < Kd);
            const double2* src = ok ? A + (size_t)(m0 + r) * lda + k0 + c : A;
            cp_async16(a_dst + r * A_LD + c, src, ok);
        }
#pragma unroll
        for (int i = 0; i < (BK * BN) / 256; ++i) {
            const int idx = tid + 256 * i;
            const int r = idx / BN, c = idx % BN;
            const bool ok = (k0 + r < Kd) && (n0 + c < N);
            const double2* src = ok ? Bm + (size_t)(k0 + r) * ldb + n0 + c : Bm;
            cp_async16(b_dst + r * B_LD + c, src, ok);
        }
    };

    double cr[2][NC][2], ci[2][NC][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int c = 0; c < NC; ++c) cr[m][c][0] = cr[m][c][1] = ci[m][c][0] = ci[m][c][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < numK) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < numK; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < numK) load_stage(nk, nk % STAGES);
            cp_async_commit();
        }
        const double2* a_s = sA + (kt % STAGES) * A_TILE + (wm * 16 + g) * A_LD + q;
        const double2* b_s = sB + (kt % STAGES) * B_TILE + q * B_LD + wn * (BN / 2) + g;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double2 a[2], b[NC];
#pragma unroll
            for (int m = 0; m < 2; ++m) a[m] = a_s[m * 8 * A_LD + kk * 4];
#pragma unroll
            for (int c = 0; c < NC; ++c) b[c] = b_s[kk * 4 * B_LD + c * 8];
            if (pre != nullptr) {
                const int k = kt * BK + kk * 4 + q;
                const double2 p = k < Kd ? pre[k] : make_double2(0.0, 0.0);
#pragma unroll
                for (int m = 0; m < 2; ++m) a[m] = cmul(a[m], p);
            }
            double nai[2];
#pragma unroll
            for (int m = 0; m < 2; ++m) nai[m] = negate(a[m].y);
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    dmma(cr[m][c][0], cr[m][c][1], a[m].x, b[c].x);
                    dmma(ci[m][c][0], ci[m][c][1], a[m].x, b[c].y);
                }
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    dmma(cr[m][c][0], cr[m][c][1], nai[m], b[c].y);
                    dmma(ci[m][c][0], ci[m][c][1], a[m].y, b[c].x);
                }
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int r = m0 + wm * 16 + m * 8 + g;
        if (r >= M) continue;
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int col = n0 + wn * (BN / 2) + c * 8 + 2 * q + i;
                if (col < N) epilogue(epi, r, col, make_double2(cr[m][c][i], ci[m][c][i]));
            }
    }
}

// ------------------------------------------------------------------------------------------------
// 3-product variant ("3M"): the DMMA issue rate is the roof of every fp64 GEMM here, and a complex tile
// product needs only three real ones:  p1 += ar br,  p2 += ai bi,  p3 += (ar + ai)(br + bi);
// re = p1 - p2, im = p3 - p1 - p2 formed once in the epilogue (the trick of rk4_shared3m_kernel, csrc/rk4_fused.cu).
// A DADD inside the k loop runs on the same fp64 pipe as the DMMAs, so the (re + im) sums are formed ONCE per
// element while the operand tiles pass through registers on their way to shared memory (global -> registers ->
// shared, the `pre` phases of the A columns are applied there too) instead of once per fragment load (2x for A,
// 4x for B).  The k loop is then 12 LDS + 24 DMMAs per k4-step and warp.
//   CTA tile 64 x 64, k-chunk 16, 8 warps as 4 x 2, warp tile 16 x 32 (2 x 4 DMMA tiles, 3 accumulators each).
//   Three shared-memory stages + one chunk in registers: chunk kt+2 is stored (and chunk kt+3 requested) in the
//   middle of the DMMAs of chunk kt, one __syncthreads per chunk; fragment loads run one k4-step ahead of the DMMAs,
//   across chunk boundaries (with one CTA per SM there are only two warps per sub-partition to hide a bubble).
//   Per stage: A complex [64][20] double2, A sums [64][20] double, B complex [16][66] double2, B sums [16][68]
//   double -- the row strides make every fragment load conflict free (128-bit loads per quarter warp: rows 2j and
//   2j+1 sit 16 banks apart; 64-bit loads per half warp: four rows tile the 32 banks).
// Error bound: normwise (a few ulp of |A||B|) instead of componentwise.
// ------------------------------------------------------------------------------------------------
constexpr int T_STAGES = 3;
constexpr int TA_LD = 20, TAS_LD = 20, TB_LD = 66, TBS_LD = 68;
constexpr size_t T_AC = (size_t)BM * TA_LD * sizeof(double2);    // 20480
constexpr size_t T_AS = (size_t)BM * TAS_LD * sizeof(double);    // 10240
constexpr size_t T_BC = (size_t)BK * TB_LD * sizeof(double2);    // 16896
constexpr size_t T_BS = (size_t)BK * TBS_LD * sizeof(double);    //  8704
constexpr size_t T_STAGE = T_AC + T_AS + T_BC + T_BS;            // 56320
constexpr size_t T_SMEM = T_STAGES * T_STAGE;                    // 168960

// SPLIT = 1: one CTA per 64 x 64 tile, 2-D grid (tiles with linear index >= tile_limit, when tile_limit > 0, are left to
// the tail launch).  SPLIT = 2, 4, 8: the TAIL of a product whose tile count leaves the last wave mostly empty -- a cluster of
// SPLIT CTAs shares one tile, each CTA multiplies 1 / SPLIT of the k range, the partial tiles are summed through
// distributed shared memory in FIXED rank order (every CTA reduces its band of rows, reading the bands of the others), so
// the result does not depend on scheduling.  729 x 4096 x 729: 768 tiles = 5 waves + 28 tiles; those 28 run as 28 x 4 CTAs.
template <typename Epi, int SPLIT>
__global__ void __launch_bounds__(256, 1) zgemm3m_kernel(int M, int N, int Kd, const double2* __restrict__ A0, int lda,
                                                          const double2* __restrict__ Bm0, int ldb,
                                                          const double2* __restrict__ pre, Epi epi, BatchStride bs,
                                                          int tile_limit, int tile0, int ntx) {
    const double2* __restrict__ A = A0 + (size_t)blockIdx.z * (size_t)bs.sA;
    const double2* __restrict__ Bm = Bm0 + (size_t)blockIdx.z * (size_t)bs.sB;
    epi.shift(blockIdx.z);
    constexpr int BN = 64;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps
    int bx = blockIdx.x, by = blockIdx.y;
    unsigned krank = 0;
    if constexpr (SPLIT > 1) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(krank));
        const int tile = tile0 + (int)(blockIdx.x / SPLIT);
        by = tile / ntx;
        bx = tile - by * ntx;
    } else if (tile_limit > 0 && (int)(blockIdx.y * gridDim.x + blockIdx.x) >= tile_limit) {
        return;
    }
    const int m0 = by * BM, n0 = bx * BN;
    const int numK_all = (Kd + BK - 1) / BK;
    // this CTA's share of the k chunks
    const int kc0 = SPLIT > 1 ? (int)((long long)numK_all * krank / SPLIT) : 0;
    const int kc1 = SPLIT > 1 ? (int)((long long)numK_all * (krank + 1) / SPLIT) : numK_all;
    const int numK = kc1 - kc0;

    // staging map: A element (ar + 16 i, ac), B element (br + 4 i, bc), i < 4
    const int ar = tid >> 4, ac = tid & 15, br = tid >> 6, bc = tid & 63;
    double2 ra[4], rb[4], rp;
    auto gload = [&](int kt) {
        const int k0 = (kc0 + kt) * BK;
        const bool kok = k0 + ac < Kd;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = m0 + ar + 16 * i;
            ra[i] = (kok && r < M) ? ldg_stream(A + (size_t)r * lda + k0 + ac) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = k0 + br + 4 * i;
            rb[i] = (r < Kd && n0 + bc < N) ? ldg_stream(Bm + (size_t)r * ldb + n0 + bc) : make_double2(0.0, 0.0);
        }
        if (pre != nullptr) rp = kok ? pre[k0 + ac] : make_double2(0.0, 0.0);
    };
    auto sstore = [&](int slot) {
        unsigned char* st = smem_raw + (size_t)slot * T_STAGE;
        double2* a_c = reinterpret_cast<double2*>(st);
        double* a_s = reinterpret_cast<double*>(st + T_AC);
        double2* b_c = reinterpret_cast<double2*>(st + T_AC + T_AS);
        double* b_s = reinterpret_cast<double*>(st + T_AC + T_AS + T_BC);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double2 v = ra[i];
            if (pre != nullptr) v = cmul(v, rp);
            a_c[(ar + 16 * i) * TA_LD + ac] = v;
            a_s[(ar + 16 * i) * TAS_LD + ac] = v.x + v.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            b_c[(br + 4 * i) * TB_LD + bc] = rb[i];
            b_s[(br + 4 * i) * TBS_LD + bc] = rb[i].x + rb[i].y;
        }
    };

    double p[3][2][4][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int c = 0; c < 4; ++c) p[a][m][c][0] = p[a][m][c][1] = 0.0;

    // fragments of one k4-step: 6 LDS.128 + 6 LDS.64, consumed by 24 DMMAs
    struct Frags {
        double2 a[2], b[4];
        double as[2], bs[4];
    };
    const int a_off = (wm * 16 + g) * TA_LD + q, as_off = (wm * 16 + g) * TAS_LD + q;
    const int b_off = q * TB_LD + wn * 32 + g, bs_off = q * TBS_LD + wn * 32 + g;
    auto load_frags = [&](int slot, int kk, Frags& f) {
        const unsigned char* st = smem_raw + (size_t)slot * T_STAGE;
        const double2* a_c = reinterpret_cast<const double2*>(st) + a_off + kk * 4;
        const double* a_s = reinterpret_cast<const double*>(st + T_AC) + as_off + kk * 4;
        const double2* b_c = reinterpret_cast<const double2*>(st + T_AC + T_AS) + b_off + kk * 4 * TB_LD;
        const double* b_s = reinterpret_cast<const double*>(st + T_AC + T_AS + T_BC) + bs_off + kk * 4 * TBS_LD;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            f.a[m] = a_c[m * 8 * TA_LD];
            f.as[m] = a_s[m * 8 * TAS_LD];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            f.b[c] = b_c[c * 8];
            f.bs[c] = b_s[c * 8];
        }
    };
    auto mma = [&](const Frags& f) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int c = 0; c < 4; ++c) dmma(p[0][m][c][0], p[0][m][c][1], f.a[m].x, f.b[c].x);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int c = 0; c < 4; ++c) dmma(p[1][m][c][0], p[1][m][c][1], f.a[m].y, f.b[c].y);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int c = 0; c < 4; ++c) dmma(p[2][m][c][0], p[2][m][c][1], f.as[m], f.bs[c]);
    };

    // Software pipeline: the fragments of k4-step i+1 are requested before the DMMAs of step i are issued -- across
    // chunk boundaries too: chunk kt+1 was stored during iteration kt-1 and published by the barrier that ended it,
    // so its first fragments can be fetched in the last k4-step of chunk kt, ahead of the barrier.  The barrier at
    // the end of iteration kt publishes chunk kt+2 (stored in the middle of the iteration, while DMMAs are queued)
    // and retires every read of stage kt % 3 before iteration kt+1 overwrites it with chunk kt+3.
    if (numK > 0) {
        gload(0);
        sstore(0);
    }
    if (numK > 1) {
        gload(1);
        sstore(1);
    }
    if (numK > 2) gload(2);
    __syncthreads();
    Frags f0, f1;
    if (numK > 0) load_frags(0, 0, f0);
    int slot = 0;
#pragma unroll 1
    for (int kt = 0; kt < numK; ++kt) {
        const int nslot = slot == T_STAGES - 1 ? 0 : slot + 1;
        load_frags(slot, 1, f1);
        mma(f0);
        load_frags(slot, 2, f0);
        if (kt + 2 < numK) sstore(nslot == T_STAGES - 1 ? 0 : nslot + 1);  // stage (kt + 2) % 3
        mma(f1);
        load_frags(slot, 3, f1);
        if (kt + 3 < numK) gload(kt + 3);
        mma(f0);
        if (kt + 1 < numK) load_frags(nslot, 0, f0);
        mma(f1);
        __syncthreads();
        slot = nslot;
    }

    if constexpr (SPLIT > 1) {
        // partial tile of this k range -> own shared memory (the stage buffers are free now), [64][65] complex
        constexpr int PLD = 65;
        double2* part = reinterpret_cast<double2*>(smem_raw);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    part[(wm * 16 + m * 8 + g) * PLD + wn * 32 + c * 8 + 2 * q + i] =
                        make_double2(p[0][m][c][i] - p[1][m][c][i], (p[2][m][c][i] - p[0][m][c][i]) - p[1][m][c][i]);
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        // rank r reduces rows [r * 64 / SPLIT, (r + 1) * 64 / SPLIT) over all ranks, in rank order
        constexpr int ROWS = 64 / SPLIT;
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(part);
        for (int e = tid; e < ROWS * 64; e += 256) {
            const int rl = (int)krank * ROWS + e / 64, cl = e % 64;
            double2 sum = make_double2(0.0, 0.0);
#pragma unroll
            for (int rk = 0; rk < SPLIT; ++rk) {
                uint32_t ra_;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra_) : "r"(base + (uint32_t)((rl * PLD + cl) * sizeof(double2))), "r"(rk));
                double2 v;
                asm volatile("ld.shared::cluster.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(ra_));
                sum.x += v.x;
                sum.y += v.y;
            }
            const int r = m0 + rl, col = n0 + cl;
            if (r < M && col < N) epilogue(epi, r, col, sum);
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const int r = m0 + wm * 16 + m * 8 + g;
            if (r >= M) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int col = n0 + wn * 32 + c * 8 + 2 * q + i;
                    if (col < N)
                        epilogue(epi, r, col, make_double2(p[0][m][c][i] - p[1][m][c][i], (p[2][m][c][i] - p[0][m][c][i]) - p[1][m][c][i]));
                }
        }
    }
}

template <typename Epi, int BN>
int launch_bn(int M, int N, int Kd, const double2* A, int lda, const double2* B, int ldb, const double2* pre,
              const Epi& epi, BatchStride bs, int count, cudaStream_t st) {
    // per launch, not cached: the attribute belongs to the current device, and one process may drive several
    QDB_CUDA(cudaFuncSetAttribute(zgemm_kernel<Epi, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TileB<BN>::SMEM));
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, count);
    zgemm_kernel<Epi, BN><<<grid, 256, TileB<BN>::SMEM, st>>>(M, N, Kd, A, lda, B, ldb, pre, epi, bs);
    QDB_LAUNCH_CHECK("zgemm_kernel");
    return QDB_OK;
}

template <typename Epi, int SPLIT>
int launch_3m_tail(int M, int N, int Kd, const double2* A, int lda, const double2* B, int ldb, const double2* pre,
                   const Epi& epi, BatchStride bs, int tile0, int tail, int ntx, cudaStream_t st) {
    auto kern = zgemm3m_kernel<Epi, SPLIT>;
    QDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(tail * SPLIT);
    lc.blockDim = dim3(256);
    lc.dynamicSmemBytes = T_SMEM;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SPLIT;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    QDB_CUDA(cudaLaunchKernelEx(&lc, kern, M, N, Kd, A, lda, B, ldb, pre, epi, bs, 0, tile0, ntx));
    QDB_LAUNCH_CHECK("zgemm3m_kernel<split>");
    return QDB_OK;
}

// QDB_ZGEMM_NO_TAIL=1 switches the split-k tail off (the whole product on one-CTA-per-tile launches: the bit-level reference)
bool zgemm_tail_enabled() {
    const char* e = getenv("QDB_ZGEMM_NO_TAIL");
    return !(e && e[0] == '1');
}

template <typename Epi>
int launch_3m(int M, int N, int Kd, const double2* A, int lda, const double2* B, int ldb, const double2* pre,
              const Epi& epi, BatchStride bs, int count, cudaStream_t st) {
    QDB_CUDA(cudaFuncSetAttribute(zgemm3m_kernel<Epi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM));
    const int ntx = (N + 63) / 64, nty = (M + BM - 1) / BM;
    dim3 grid(ntx, nty, count);
    // Wave quantisation: with one CTA per SM a product of T tiles takes ceil(T / #SM) tile times.  When the last wave is
    // less than half full, its tiles are split along k over clusters of 2 / 4 / 8 CTAs (deterministic reduction through
    // DSMEM) so that it costs 1/2 .. 1/8 of a tile time instead of a whole one.
    const int SMS = sm_count();
    const int tiles = ntx * nty;
    const int tail = tiles % SMS;
    int split = 1;
    if (count == 1 && tiles > SMS && tail > 0 && 2 * tail <= SMS && Kd >= 256 && zgemm_tail_enabled()) {
        split = 2;
        while (split < 8 && 2 * split * tail <= SMS) split *= 2;
    }
    if (split == 1) {
        zgemm3m_kernel<Epi, 1><<<grid, 256, T_SMEM, st>>>(M, N, Kd, A, lda, B, ldb, pre, epi, bs, 0, 0, ntx);
        QDB_LAUNCH_CHECK("zgemm3m_kernel");
        return QDB_OK;
    }
    zgemm3m_kernel<Epi, 1><<<grid, 256, T_SMEM, st>>>(M, N, Kd, A, lda, B, ldb, pre, epi, bs, tiles - tail, 0, ntx);
    QDB_LAUNCH_CHECK("zgemm3m_kernel");
    if (split == 2) return launch_3m_tail<Epi, 2>(M, N, Kd, A, lda, B, ldb, pre, epi, bs, tiles - tail, tail, ntx, st);
    if (split == 4) return launch_3m_tail<Epi, 4>(M, N, Kd, A, lda, B, ldb, pre, epi, bs, tiles - tail, tail, ntx, st);
    return launch_3m_tail<Epi, 8>(M, N, Kd, A, lda, B, ldb, pre, epi, bs, tiles - tail, tail, ntx, st);
}

// QDB_ZGEMM_4M=1 pins the 4-product kernel (bit-level reference of the 3-product one in the tests)
bool zgemm_3m_enabled() {
    const char* e = getenv("QDB_ZGEMM_4M");
    return !(e && e[0] == '1');
}

template <typename Epi>
int launch(int M, int N, int Kd, const double2* A, int lda, const double2* B, int ldb, const double2* pre,
           const Epi& epi, cudaStream_t st, BatchStride bs = BatchStride{0, 0}, int count = 1) {
    // 3-product kernel (one CTA per SM) once its 64 x 64 grid (times the batch) fills at least half of the SMs and the
    // k loop is long enough to amortise the three-stage fill; small products stay on the 4-product kernel (two CTAs
    // per SM, narrow tiles)
    const long tiles64 = (long)((N + 63) / 64) * ((M + BM - 1) / BM) * count;
    if (zgemm_3m_enabled() && 2 * tiles64 >= sm_count() && Kd >= 64)
        return launch_3m<Epi>(M, N, Kd, A, lda, B, ldb, pre, epi, bs, count, st);
    // 64 x 64 tiles unless they would fill fewer than the 2 CTA slots per SM
    if (tiles64 < 2L * sm_count() && N > 32) return launch_bn<Epi, 32>(M, N, Kd, A, lda, B, ldb, pre, epi, bs, count, st);
    return launch_bn<Epi, 64>(M, N, Kd, A, lda, B, ldb, pre, epi, bs, count, st);
}

}  // namespace

int launch_zgemm(int M, int N, int Kd, const double2* A, int lda, const double2* B, int ldb, double2* C, int ldc,
                 double2 alpha, double2 beta, const double* colscale, const double2* pre, const double2* post,
                 cudaStream_t st) {
    if (M == 0 || N == 0) return QDB_OK;
    // large products: the int8 tensor-core emulation (zgemm_ozaki.cu)
    if (zgemm_int8_preferred(M, N, Kd)) return launch_zgemm_int8(M, N, Kd, A, lda, B, ldb, C, ldc, alpha, beta, colscale, pre, post, st);
    EpiStd e{C, ldc, alpha, beta, colscale, post, 0};
    return launch(M, N, Kd, A, lda, B, ldb, pre, e, st);
}

// count independent products C_z = alpha A_z B_z + beta C_z, z < count, operands sA / sB / sC elements apart (grid.z; a
// stride of 0 shares the operand).  One launch fills the chip with products that are far too small to do so alone:
// the step propagators of the time-parallel solvers (propagator.cu).
int launch_zgemm_batched(int M, int N, int Kd, const double2* A, int lda, long long sA, const double2* B, int ldb, long long sB,
                         double2* C, int ldc, long long sC, double2 alpha, double2 beta, int count, cudaStream_t st) {
    if (M == 0 || N == 0 || count == 0) return QDB_OK;
    // products that are large enough on their own: the int8 tensor-core emulation, side by side
    if (zgemm_int8_preferred(M, N, Kd)) return launch_zgemm_int8_batched(M, N, Kd, A, lda, sA, B, ldb, sB, C, ldc, sC, alpha, beta, count, st);
    for (int z0 = 0; z0 < count; z0 += 65535) {  // grid.z limit
        const int c = count - z0 < 65535 ? count - z0 : 65535;
        EpiStd e{C + (size_t)z0 * sC, ldc, alpha, beta, nullptr, nullptr, sC};
        const int rc = launch(M, N, Kd, A + (size_t)z0 * sA, lda, B + (size_t)z0 * sB, ldb, (const double2*)nullptr, e, st,
                              BatchStride{sA, sB}, c);
        if (rc != QDB_OK) return rc;
    }
    return QDB_OK;
}

int launch_zgemm_rk4stage(int n, int B, const double2* G, const double2* yin, int ldy, const double2* ybase,
                          double2* yout, double2* acc, double a_next, double w, int first, cudaStream_t st) {
    if (n == 0 || B == 0) return QDB_OK;
    if (zgemm_int8_preferred(n, B, n)) return launch_zgemm_int8_rk4stage(n, B, G, yin, ldy, ybase, yout, acc, a_next, w, first, st);
    EpiRk4 e{ybase, yout, acc, ldy, a_next, w, first};
    return launch(n, B, n, G, n, yin, ldy, (const double2*)nullptr, e, st);
}

}  // namespace qdb
