// Fused fixed-step RK4: S steps x 4 stages in ONE launch, state tile resident on chip
// (SURVEY.md 8(a) rows a2 + a3 + a7 + a8 inner loop).
//
// Decomposition.  Every state column evolves independently, so a CTA owns 8*NCT whole columns for
// the entire launch: no inter-CTA communication, HBM traffic = read y once + write y once.
//   * A operand (the generator) streams L2 -> registers directly in DMMA A-fragment order
//     (QDB_LAYOUT_PACKED: one coalesced 512 B LDG.128 per warp per fragment) through a 4-deep
//     register ring that runs 3 k-tiles ahead and keeps streaming across stage boundaries.  Each
//     warp owns distinct row tiles, so there is no intra-CTA reuse that shared-memory staging
//     could exploit.
//   * B operand (the stage vector) lives in shared memory in DMMA B-fragment order, double
//     buffered; fragments are fetched one k-tile ahead; the epilogue of stage s writes the stage
//     s+1 input (XOR-swizzled so both the fragment loads and the scattered epilogue stores are
//     bank-conflict free per quarter warp, the unit in which 16 B shared accesses are served).
//   * The RK4 k-sum lives in registers, y itself in a thread-private shared-memory slab.
//
// Shared-signal mode: A = precomputed generator table entry G_frame(t_stage) (frame phases folded
// in by generator_kernel).  Sweep mode: A = the K+1 stored operators, the per-column signal value
// scales the B fragment (so the operator sum accumulates in the same DMMA accumulators) and the
// frame phases are applied to B rows on write and to C rows on read.
//
// Split mode (shared-signal kernel): whole-column CTAs quantise the batch in units of 8 NCW columns
// per SM (4096 columns = 512 octets over 148 SMs -> 4 octets on the busiest SM = 86.5 % of the
// chip).  A 2-CTA cluster instead owns 2 NCW + 1 octets: NCW octets per CTA plus one octet whose
// output ROWS are split between the two CTAs (row tile m < MR/2 on rank 0, m >= MR/2 on rank 1).
// Only that octet's stage vector crosses DSMEM (8 KiB per CTA per stage at n=128), and the per-stage
// __syncthreads becomes a cluster barrier.  74 clusters x 7 octets cover 4096 columns with 98.8 %
// of the SM-time busy.
//
// Roofline: fp64 tensor pipe.  Algorithmic flops per column per step = 4(8n^2 + 12n) + 28n.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "qdb_common.cuh"
#include "rk4_device.cuh"

namespace qdb {

namespace {

constexpr int RING = 4;  // A-fragment register ring; prefetch distance RING-1 k-tiles (KT % RING == 0)

struct Geometry {
    int n, npad, KT, RT;  // KT = kpad/4 k-tiles (kpad = n rounded to 16), RT = npad/8 row tiles
    int WR, WC;           // warps along rows / columns
    int NCT;              // column tiles per CTA
};

// position of state element (row tile rt, row-in-tile g, column tile ct, column-in-tile cin) in the
__device__ __forceinline__ int yin_pos(int NCT, int rt, int g, int ct, int cin) {
    const int kt = 2 * rt + (g >> 2);
    const int lane_b = (g & 3) + 4 * cin;
    return (kt * NCT + ct) * 32 + frag_swizzle(lane_b);
}

// Complex tile product as real DMMAs.
//   M3 = false ("4M"): cr += ar br - ai bi, ci += ar bi + ai br            -> 4 DMMAs per k-tile
//   M3 = true  ("3M"): p1 += ar br, p2 += ai bi, p3 += (ar+ai)(br+bi);     -> 3 DMMAs per k-tile
//                      re = p1 - p2, im = p3 - p1 - p2 (formed once per stage in the epilogue)
// 3M trades a quarter of the tensor-pipe work for one more accumulator per tile, two DADDs per fragment
// and a normwise (instead of componentwise) error bound of a few ulp of |A||B| -- five orders of
// magnitude inside the 1e-8 parity bar.
template <int MR, int NCW, bool M3>
struct Accum {
    static constexpr int NA = M3 ? 3 : 2;
    double p[NA][MR][NCW][2];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int m = 0; m < MR; ++m)
#pragma unroll
                for (int c = 0; c < NCW; ++c) p[a][m][c][0] = p[a][m][c][1] = 0.0;
    }
    __device__ __forceinline__ double re(int m, int c, int i) const { return M3 ? p[0][m][c][i] - p[1][m][c][i] : p[0][m][c][i]; }
    __device__ __forceinline__ double im(int m, int c, int i) const {
        return M3 ? (p[NA - 1][m][c][i] - p[0][m][c][i]) - p[1][m][c][i] : p[1][m][c][i];
    }
};

// own tiles (MR x NCW) plus, in split mode, MS row tiles of the shared octet (B fragment b[NCW], A fragments a_s
// = this rank's half).  Every accumulator is touched once (3M) or twice half a block apart (4M) per k-tile.
template <int MR, int NCW, int MS, bool SPLIT>
__device__ __forceinline__ void mma_block(Accum<MR, NCW, false>& acc, Accum<MS, 1, false>& accs, const double2 (&a)[MR],
                                          const double2 (&a_s)[MS], const double2 (&b)[NCW + (SPLIT ? 1 : 0)]) {
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) {
            dmma(acc.p[0][m][c][0], acc.p[0][m][c][1], a[m].x, b[c].x);
            dmma(acc.p[1][m][c][0], acc.p[1][m][c][1], a[m].x, b[c].y);
        }
    if constexpr (SPLIT) {
#pragma unroll
        for (int m = 0; m < MS; ++m) {
            dmma(accs.p[0][m][0][0], accs.p[0][m][0][1], a_s[m].x, b[NCW].x);
            dmma(accs.p[1][m][0][0], accs.p[1][m][0][1], a_s[m].x, b[NCW].y);
        }
    }
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) {
            dmma(acc.p[0][m][c][0], acc.p[0][m][c][1], -a[m].y, b[c].y);  // SASS: DMMA with negated operand
            dmma(acc.p[1][m][c][0], acc.p[1][m][c][1], a[m].y, b[c].x);
        }
    if constexpr (SPLIT) {
#pragma unroll
        for (int m = 0; m < MS; ++m) {
            dmma(accs.p[0][m][0][0], accs.p[0][m][0][1], -a_s[m].y, b[NCW].y);
            dmma(accs.p[1][m][0][0], accs.p[1][m][0][1], a_s[m].y, b[NCW].x);
        }
    }
}

// 3M block: fragments arrive with their (re + im) sums precomputed (A: third plane of the generator table,
// B: third plane of the stage buffer written by the epilogue), so the k loop issues DMMAs only -- a DADD in
// the loop costs about 9 cycles of the same fp64 pipe the DMMAs need.
struct Frag3 {
    double2 c;  // (re, im)
    double s;   // re + im
};
template <int MR, int NCW, int MS, bool SPLIT>
__device__ __forceinline__ void mma_block3(Accum<MR, (NCW > 0 ? NCW : 1), true>& acc, Accum<MS, 1, true>& accs, const Frag3 (&a)[MR],
                                           const Frag3 (&a_s)[MS], const Frag3 (&b)[NCW + (SPLIT ? 1 : 0)]) {
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) dmma(acc.p[0][m][c][0], acc.p[0][m][c][1], a[m].c.x, b[c].c.x);
    if constexpr (SPLIT) {
#pragma unroll
        for (int m = 0; m < MS; ++m) dmma(accs.p[0][m][0][0], accs.p[0][m][0][1], a_s[m].c.x, b[NCW].c.x);
    }
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) dmma(acc.p[1][m][c][0], acc.p[1][m][c][1], a[m].c.y, b[c].c.y);
    if constexpr (SPLIT) {
#pragma unroll
        for (int m = 0; m < MS; ++m) dmma(accs.p[1][m][0][0], accs.p[1][m][0][1], a_s[m].c.y, b[NCW].c.y);
    }
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) dmma(acc.p[2][m][c][0], acc.p[2][m][c][1], a[m].s, b[c].s);
    if constexpr (SPLIT) {
#pragma unroll
        for (int m = 0; m < MS; ++m) dmma(accs.p[2][m][0][0], accs.p[2][m][0][1], a_s[m].s, b[NCW].s);
    }
}

// ---- cluster helpers (split mode) ----
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 B into the peer CTA's shared memory at the offset of local pointer p, completing on the peer's copy of bar
__device__ __forceinline__ void st_async_peer(const double2* p, const uint64_t* bar, unsigned peer, double2 v) {
    uint32_t rp, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rp) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(peer));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(peer));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1,%2}, [%3];" ::"r"(rp), "d"(v.x),
                 "d"(v.y), "r"(rb)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------
// shared-signal mode
// ------------------------------------------------------------------------------------------------
template <int MR, int NCW, bool SPLIT, int SKT>
struct StaticGeo {
    int n;
    static constexpr int npad = 4 * SKT, KT = SKT, RT = 8 * MR, WR = 8, WC = 1, NCT = NCW + (SPLIT ? 1 : 0);
};

// SKT > 0 fixes the geometry at compile time (KT = SKT k-tiles, 8 row warps x 1 column warp, 256 threads,
// RT = 8 MR row tiles, NCT = NCW (+1 in split mode)): addresses become immediates and about 25 registers
// that otherwise carry precomputed epilogue offsets across the main loop are freed.
template <int MR, int NCW, bool SPLIT, int SKT>
__global__ void __launch_bounds__(256, 1)
rk4_shared_kernel(Geometry geo_rt, int B, int S, const double2* __restrict__ gen, double h, double2* __restrict__ y,
                  int ldy) {
    static_assert(!SPLIT || MR % 2 == 0, "split mode halves the row tiles of the shared octet");
    typename std::conditional<(SKT > 0), StaticGeo<MR, NCW, SPLIT, SKT>, Geometry>::type geo;
    if constexpr (SKT > 0) geo.n = geo_rt.n; else geo = geo_rt;
    constexpr int MS = SPLIT ? MR / 2 : 1;         // row tiles of the shared octet per warp (1 = unused dummy)
    constexpr int NB = NCW + (SPLIT ? 1 : 0);      // B fragments per k-tile per warp
    constexpr int SLOTS = (MR * NCW + (SPLIT ? MS : 0)) * 2;  // y-slab entries per thread
    extern __shared__ __align__(16) double2 sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int swl = frag_swizzle(lane);
    const int wr = warp % geo.WR, wc = warp / geo.WR;  // split mode: WC == 1
    const int KT = geo.KT, NCT = geo.NCT, n = geo.n;
    const size_t entry_elems = (size_t)geo.npad * KT * 4;
    const int yin_elems = KT * NCT * 32;
    const int yst_off = 2 * yin_elems;  // thread-private y slab: [SLOTS][blockDim]
    const int nthr = SKT > 0 ? 256 : blockDim.x;
    // first global column of this warp's own tiles / of the shared octet; local stage-buffer tile index
    const unsigned rank = SPLIT ? cluster_ctarank() : 0u;
    const int ct0 = wc * NCW;  // local index of own tile 0
    int colw, cols = 0;
    if (SPLIT) {
        const int base = (blockIdx.x >> 1) * (2 * NCW + 1);
        colw = 8 * (base + (int)rank * (NCW + 1));
        cols = 8 * (base + NCW);
    } else {
        colw = 8 * (blockIdx.x * NCT + ct0);
    }

    int rt[MR], rtl[MR];  // row tile owned / loaded (surplus warps recompute the last tile, results dropped)
    bool mvalid[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        rt[m] = wr + geo.WR * m;
        mvalid[m] = rt[m] < geo.RT;
        rtl[m] = mvalid[m] ? rt[m] : geo.RT - 1;
    }
    int rts[MS];  // this rank's row tiles of the shared octet
    bool svalid[MS];
#pragma unroll
    for (int mm = 0; mm < MS; ++mm) {
        rts[mm] = wr + geo.WR * ((int)rank * MS + mm);
        svalid[mm] = SPLIT && rts[mm] < geo.RT;
    }

    // ---- zero both stage buffers (k rows beyond npad stay zero), load y ----
    for (int i = tid; i < 2 * yin_elems; i += nthr) sm[i] = make_double2(0.0, 0.0);
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = colw + 8 * c + 2 * q + i;
                double2 v = make_double2(0.0, 0.0);
                if (mvalid[m] && row < n && col < B) v = y[(size_t)row * ldy + col];
                sm[yst_off + ((m * NCW + c) * 2 + i) * nthr + tid] = v;
                if (mvalid[m]) sm[yin_pos(NCT, rt[m], g, ct0 + c, 2 * q + i)] = v;
            }
    if (SPLIT) {
        // shared octet: every CTA stages ALL rows, but keeps only its own half in the y slab
#pragma unroll
        for (int m = 0; m < MR; ++m)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = cols + 2 * q + i;
                double2 v = make_double2(0.0, 0.0);
                if (mvalid[m] && row < n && col < B) v = y[(size_t)row * ldy + col];
                if (mvalid[m]) sm[yin_pos(NCT, rt[m], g, NCW, 2 * q + i)] = v;
                const int mm = m - (int)rank * MS;
                if (mm >= 0 && mm < MS) sm[yst_off + ((MR * NCW + mm) * 2 + i) * nthr + tid] = v;
            }
    }

    Accum<MR, NCW, false> acc;
    acc.zero();
    Accum<MS, 1, false> accs;  // shared octet (split mode)
    accs.zero();
    double kr[MR][NCW][2], ki[MR][NCW][2];  // running k1 + 2 k2 + 2 k3 + k4
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) kr[m][c][0] = kr[m][c][1] = ki[m][c][0] = ki[m][c][1] = 0.0;
    double ksr[MS][2], ksi[MS][2];
#pragma unroll
    for (int mm = 0; mm < MS; ++mm) ksr[mm][0] = ksr[mm][1] = ksi[mm][0] = ksi[mm][1] = 0.0;

    // A-fragment offsets of this warp's row tiles inside one table entry
    size_t aoff[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) aoff[m] = (size_t)rtl[m] * KT * 32 + lane;

    // prime the ring with k-tiles 0 .. RING-2 of the first stage
    double2 ring[RING][MR];
#pragma unroll
    for (int u = 0; u < RING - 1; ++u)
#pragma unroll
        for (int m = 0; m < MR; ++m) ring[u][m] = ldg_stream(gen + aoff[m] + (size_t)u * 32);

    int cur = 0;
    // split mode: two mbarriers (stage parity) count the bytes the peer sends per stage
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + yst_off + SLOTS * nthr);
    unsigned tx_bytes = 0;
    if (SPLIT) {
        const int first = (int)(rank ^ 1u) * MS * geo.WR;  // peer's row tiles of the shared octet
        const int cnt = max(0, min(geo.RT - first, MS * geo.WR));
        tx_bytes = (unsigned)cnt * 64u * (unsigned)sizeof(double2);
        if (tid == 0) {
            mbar_init(mbar, 1);
            mbar_init(mbar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_barrier();  // peer's buffers and mbarriers are initialised before the first remote store can land
    } else {
        __syncthreads();
    }

    const int total_stages = 4 * S;
#pragma unroll 1
    for (int sidx = 0; sidx < total_stages; ++sidx) {
        const int step = sidx >> 2, stage = sidx & 3;
        const int entry = 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
        // table entry of the NEXT stage (the ring streams into it during the last RING-1 k-tiles)
        const int nstage = (stage + 1) & 3, nstep = step + (stage == 3 ? 1 : 0);
        const int nentry = (sidx + 1 < total_stages) ? 2 * nstep + (nstage == 0 ? 0 : (nstage == 3 ? 2 : 1)) : entry;
        const double2* gcur = gen + (size_t)entry * entry_elems;
        const double2* gnxt = gen + (size_t)nentry * entry_elems;
        const int ybase = cur * yin_elems + ct0 * 32;
        if (SPLIT && tid == 0) mbar_expect_tx(mbar + (sidx & 1), tx_bytes);

        double2 bfrag[2][NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) bfrag[0][c] = sm[ybase + c * 32 + swl];

#pragma unroll 1
        for (int kt0 = 0; kt0 < KT; kt0 += RING) {
#pragma unroll
            for (int u = 0; u < RING; ++u) {
                const int kt = kt0 + u;
                // A fragments of k-tile kt + RING-1 into the ring slot the previous k-tile just released
                {
                    const int ktn = kt + RING - 1;
                    const double2* src = ktn < KT ? gcur + (size_t)ktn * 32 : gnxt + (size_t)(ktn - KT) * 32;
#pragma unroll
                    for (int m = 0; m < MR; ++m) ring[(u + RING - 1) % RING][m] = ldg_stream(src + aoff[m]);
                }
                // B fragments of k-tile kt + 1 (clamped at the stage end: a harmless reload)
                {
                    const int ktb = min(kt + 1, KT - 1);
#pragma unroll
                    for (int c = 0; c < NB; ++c) bfrag[(u + 1) & 1][c] = sm[ybase + (ktb * NCT + c) * 32 + swl];
                }
                double2 a_s[MS];
                if constexpr (SPLIT) {
#pragma unroll
                    for (int mm = 0; mm < MS; ++mm) a_s[mm] = rank ? ring[u][MS + mm] : ring[u][mm];
                }
                mma_block<MR, NCW, MS, SPLIT>(acc, accs, ring[u], a_s, bfrag[u & 1]);
            }
        }

        // ---- epilogue: RK4 stage combine, write next stage input ----
        const StageCoef sc(stage, h);
        const int ydst = (cur ^ 1) * yin_elems;
        if constexpr (SPLIT) {
            // shared octet first: its remote stores are in flight while the own tiles are combined
#pragma unroll
            for (int mm = 0; mm < MS; ++mm)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double k_r = accs.re(mm, 0, i), k_i = accs.im(mm, 0, i);
                    const int slab = yst_off + ((MR * NCW + mm) * 2 + i) * nthr + tid;
                    const double2 yv = sm[slab];
                    ksr[mm][i] = sc.keep * ksr[mm][i] + sc.wk * k_r;
                    ksi[mm][i] = sc.keep * ksi[mm][i] + sc.wk * k_i;
                    const double v_r = sc.last ? ksr[mm][i] : k_r, v_i = sc.last ? ksi[mm][i] : k_i;
                    const double2 nxt = make_double2(yv.x + sc.astep * v_r, yv.y + sc.astep * v_i);
                    if (sc.last) sm[slab] = nxt;
                    if (svalid[mm]) {
                        double2* dst = sm + ydst + yin_pos(NCT, rts[mm], g, NCW, 2 * q + i);
                        *dst = nxt;
                        st_async_peer(dst, mbar + (sidx & 1), rank ^ 1u, nxt);
                    }
                }
            accs.zero();
        }
        // one row tile at a time: all y-slab loads first, then the stores (the compiler cannot hoist a shared
        // load above a shared store it cannot disambiguate, which would serialise on the LDS latency)
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            double2 yv[NCW][2];
#pragma unroll
            for (int c = 0; c < NCW; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i) yv[c][i] = sm[yst_off + ((m * NCW + c) * 2 + i) * nthr + tid];
#pragma unroll
            for (int c = 0; c < NCW; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double k_r = acc.re(m, c, i), k_i = acc.im(m, c, i);
                    kr[m][c][i] = sc.keep * kr[m][c][i] + sc.wk * k_r;
                    ki[m][c][i] = sc.keep * ki[m][c][i] + sc.wk * k_i;
                    const double v_r = sc.last ? kr[m][c][i] : k_r, v_i = sc.last ? ki[m][c][i] : k_i;
                    const double2 nxt = make_double2(yv[c][i].x + sc.astep * v_r, yv[c][i].y + sc.astep * v_i);
                    if (sc.last) sm[yst_off + ((m * NCW + c) * 2 + i) * nthr + tid] = nxt;
                    if (mvalid[m]) sm[ydst + yin_pos(NCT, rt[m], g, ct0 + c, 2 * q + i)] = nxt;
                }
        }
        acc.zero();
        cur ^= 1;
        __syncthreads();
        if (SPLIT) mbar_wait(mbar + (sidx & 1), (sidx >> 1) & 1);  // the peer's half of the shared octet has landed
    }
    if (SPLIT) cluster_barrier();  // neither CTA retires while the other could still address its shared memory

    // ---- store y ----
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = colw + 8 * c + 2 * q + i;
                if (mvalid[m] && row < n && col < B)
                    y[(size_t)row * ldy + col] = sm[yst_off + ((m * NCW + c) * 2 + i) * nthr + tid];
            }
    if (SPLIT) {
#pragma unroll
        for (int mm = 0; mm < MS; ++mm)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rts[mm] + g;
                const int col = cols + 2 * q + i;
                if (svalid[mm] && row < n && col < B)
                    y[(size_t)row * ldy + col] = sm[yst_off + ((MR * NCW + mm) * 2 + i) * nthr + tid];
            }
    }
}

// ------------------------------------------------------------------------------------------------
// shared-signal mode, 3M complex products
// ------------------------------------------------------------------------------------------------
// Same decomposition as rk4_shared_kernel, with three changes that make the 3-DMMA complex product pay:
//   * the generator table is QDB_LAYOUT_PACKED3M: every entry carries a third plane re + im, streamed
//     beside the complex plane (LDG.128 + LDG.64 per fragment);
//   * the stage buffer carries a third plane too, written ONCE per element by the epilogue instead of being
//     re-derived by every warp in the k loop;
//   * to pay for that plane the own-tile stage buffer is single buffered (barrier before and after the
//     epilogue stores); only the shared octet, which the peer CTA writes asynchronously, stays double buffered.
template <int MR, int NCW, bool SPLIT, int SKT>
struct StaticGeo3 {
    int n;
    static constexpr int npad = 4 * SKT, KT = SKT, RT = 8 * MR, WR = 8, WC = 1, NCT = NCW;  // NCT = own tiles
};

__device__ __forceinline__ double ldg_stream_f64(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_async_peer_f64(const double* p, const uint64_t* bar, unsigned peer, double v) {
    uint32_t rp, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rp) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(peer));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(peer));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(rp), "d"(v), "r"(rb)
                 : "memory");
}

template <int MR, int NCW, bool SPLIT, int SKT>
__global__ void __launch_bounds__(256, 1)
rk4_shared3m_kernel(Geometry geo_rt, int B, int S, const double2* __restrict__ gen, double h, double2* __restrict__ y,
                    int ldy) {
    static_assert(!SPLIT || MR % 2 == 0, "split mode halves the row tiles of the shared octet");
    typename std::conditional<(SKT > 0), StaticGeo3<MR, NCW, SPLIT, SKT>, Geometry>::type geo;
    if constexpr (SKT > 0) geo.n = geo_rt.n; else geo = geo_rt;
    static_assert(NCW > 0 || SPLIT, "a CTA without own column tiles only exists in split mode");
    constexpr int MS = SPLIT ? MR / 2 : 1;
    constexpr int NB = NCW + (SPLIT ? 1 : 0);
    constexpr int SLOTS = (MR * NCW + (SPLIT ? MS : 0)) * 2;
    // NCW == 0: pure row split -- the cluster owns ONE column octet, each CTA half of its rows (small batches:
    // 64 octets keep 128 SMs busy instead of 64).  Only this rank's row tiles of A are streamed then, and the k-tiles
    // are so short (3 DMMAs per warp) that the fragment ring is deepened to stay ahead of the L2 latency.
    constexpr int NC1 = NCW > 0 ? NCW : 1;             // array extents
    constexpr int MA = NCW > 0 ? MR : MS;              // row tiles of A a warp loads
    constexpr int RG = (NCW <= 2 && !(SPLIT && NCW == 2) && SKT > 0 && SKT % 8 == 0) ? 8 : RING;  // few DMMAs per k-tile: deeper ring against the L2 latency
    extern __shared__ __align__(16) double2 sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int swl = frag_swizzle(lane);
    const int wr = warp % geo.WR, wc = warp / geo.WR;
    const int KT = geo.KT, NCO = geo.NCT, n = geo.n;  // NCO = own column tiles of the CTA
    const size_t entry_elems = (size_t)geo.npad * KT * 4;
    const int nthr = SKT > 0 ? 256 : blockDim.x;
    // shared-memory carve-up: complex planes first, then the sum planes
    const int own_elems = KT * NCO * 32, sh_elems = SPLIT ? KT * 32 : 0;
    double2* own_c = sm;
    double2* sh_c = own_c + own_elems;     // [2][sh_elems]
    double2* slab = sh_c + 2 * sh_elems;   // [SLOTS][nthr] thread-private y
    double* own_s = reinterpret_cast<double*>(slab + SLOTS * nthr);
    double* sh_s = own_s + own_elems;      // [2][sh_elems]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sh_s + 2 * sh_elems);

    const unsigned rank = SPLIT ? cluster_ctarank() : 0u;
    const int ct0 = wc * NCW;
    int colw, cols = 0;
    if (SPLIT) {
        const int base = (blockIdx.x >> 1) * (2 * NCW + 1);
        colw = 8 * (base + (int)rank * (NCW + 1));
        cols = 8 * (base + NCW);
    } else {
        colw = 8 * (blockIdx.x * NCO + ct0);
    }

    int rt[MR], rtl[MR];
    bool mvalid[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        rt[m] = wr + geo.WR * m;
        mvalid[m] = rt[m] < geo.RT;
        rtl[m] = mvalid[m] ? rt[m] : geo.RT - 1;
    }
    int rts[MS];
    bool svalid[MS];
#pragma unroll
    for (int mm = 0; mm < MS; ++mm) {
        rts[mm] = wr + geo.WR * ((int)rank * MS + mm);
        svalid[mm] = SPLIT && rts[mm] < geo.RT;
    }

    // ---- zero the stage planes (k rows beyond n stay zero), load y ----
    for (int i = tid; i < own_elems + 2 * sh_elems; i += nthr) sm[i] = make_double2(0.0, 0.0);
    for (int i = tid; i < own_elems + 2 * sh_elems; i += nthr) own_s[i] = 0.0;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = colw + 8 * c + 2 * q + i;
                double2 v = make_double2(0.0, 0.0);
                if (mvalid[m] && row < n && col < B) v = y[(size_t)row * ldy + col];
                slab[((m * NCW + c) * 2 + i) * nthr + tid] = v;
                if (mvalid[m]) {
                    const int pos = yin_pos(NCO, rt[m], g, ct0 + c, 2 * q + i);
                    own_c[pos] = v;
                    own_s[pos] = v.x + v.y;
                }
            }
    if (SPLIT) {
#pragma unroll
        for (int m = 0; m < MR; ++m)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = cols + 2 * q + i;
                double2 v = make_double2(0.0, 0.0);
                if (mvalid[m] && row < n && col < B) v = y[(size_t)row * ldy + col];
                if (mvalid[m]) {
                    const int pos = yin_pos(1, rt[m], g, 0, 2 * q + i);
                    sh_c[pos] = v;
                    sh_s[pos] = v.x + v.y;
                }
                const int mm = m - (int)rank * MS;
                if (mm >= 0 && mm < MS) slab[((MR * NCW + mm) * 2 + i) * nthr + tid] = v;
            }
    }

    Accum<MR, NC1, true> acc;
    acc.zero();
    Accum<MS, 1, true> accs;
    accs.zero();
    double kr[MR][NC1][2], ki[MR][NC1][2];
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) kr[m][c][0] = kr[m][c][1] = ki[m][c][0] = ki[m][c][1] = 0.0;
    double ksr[MS][2], ksi[MS][2];
#pragma unroll
    for (int mm = 0; mm < MS; ++mm) ksr[mm][0] = ksr[mm][1] = ksi[mm][0] = ksi[mm][1] = 0.0;

    size_t aoff[MA];
#pragma unroll
    for (int m = 0; m < MA; ++m) {
        const int tile = NCW > 0 ? rtl[m] : (svalid[m % MS] ? rts[m % MS] : geo.RT - 1);
        aoff[m] = (size_t)tile * KT * 32 + lane;
    }
    // entry e: complex plane at gen + e * (3/2) entry_elems, sum plane right behind it
    const size_t entry_stride = entry_elems + entry_elems / 2;  // in double2 units (entry_elems is even)

    Frag3 ring[RG][MA];
#pragma unroll
    for (int u = 0; u < RG - 1; ++u)
#pragma unroll
        for (int m = 0; m < MA; ++m) {
            ring[u][m].c = ldg_stream(gen + aoff[m] + (size_t)u * 32);
            ring[u][m].s = ldg_stream_f64(reinterpret_cast<const double*>(gen + entry_elems) + aoff[m] + (size_t)u * 32);
        }
    const double2* pc[MA];  // running prefetch pointers: next fragment of the complex / sum plane
    const double* ps[MA];
#pragma unroll
    for (int m = 0; m < MA; ++m) {
        pc[m] = gen + aoff[m] + (RG - 1) * 32;
        ps[m] = reinterpret_cast<const double*>(gen + entry_elems) + aoff[m] + (RG - 1) * 32;
    }

    int cur = 0;
    unsigned tx_bytes = 0;
    if (SPLIT) {
        const int first = (int)(rank ^ 1u) * MS * geo.WR;
        const int cnt = max(0, min(geo.RT - first, MS * geo.WR));
        tx_bytes = (unsigned)cnt * 64u * 24u;  // complex + sum plane
        if (tid == 0) {
            mbar_init(mbar, 1);
            mbar_init(mbar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_barrier();
    } else {
        __syncthreads();
    }

    const int total_stages = 4 * S;
#pragma unroll 1
    for (int sidx = 0; sidx < total_stages; ++sidx) {
        const int step = sidx >> 2, stage = sidx & 3;
        const int entry = 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
        const int nstage = (stage + 1) & 3, nstep = step + (stage == 3 ? 1 : 0);
        const int nentry = (sidx + 1 < total_stages) ? 2 * nstep + (nstage == 0 ? 0 : (nstage == 3 ? 2 : 1)) : entry;
        const double2* gnxt = gen + (size_t)nentry * entry_stride;
        const double2* bc = own_c + ct0 * 32 + swl;
        const double* bs = own_s + ct0 * 32 + swl;
        const double2* shc = sh_c + cur * sh_elems + swl;
        const double* shs = sh_s + cur * sh_elems + swl;
        if (SPLIT && tid == 0) mbar_expect_tx(mbar + (sidx & 1), tx_bytes);

        Frag3 bfrag[2][NB];
#pragma unroll
        for (int c = 0; c < NCW; ++c) {
            bfrag[0][c].c = bc[c * 32];
            bfrag[0][c].s = bs[c * 32];
        }
        if constexpr (SPLIT) {
            bfrag[0][NCW].c = shc[0];
            bfrag[0][NCW].s = shs[0];
        }

        // k loop in blocks of RG k-tiles.  The fragment ring runs RG-1 k-tiles ahead through running pointers
        // (pc / ps: next fragment of the complex / sum plane); only the LAST block of a stage crosses into the next
        // table entry, so it is peeled and the steady blocks carry no entry-selection arithmetic at all.
        const double2* nc = gnxt;                                                // next entry, complex plane
        const double* ns = reinterpret_cast<const double*>(gnxt + entry_elems);  // next entry, sum plane
        auto k_block = [&](auto last_tag, int kt0) {
            constexpr bool LAST = decltype(last_tag)::value;
            const double2* bck = bc + (size_t)kt0 * NCO * 32;
            const double* bsk = bs + (size_t)kt0 * NCO * 32;
            const double2* shck = shc + kt0 * 32;
            const double* shsk = shs + kt0 * 32;
#pragma unroll
            for (int u = 0; u < RG; ++u) {
                if (!LAST || u == 0) {  // fragment kt + RG-1 of this entry
#pragma unroll
                    for (int m = 0; m < MA; ++m) {
                        ring[(u + RG - 1) % RG][m].c = ldg_stream(pc[m]);
                        ring[(u + RG - 1) % RG][m].s = ldg_stream_f64(ps[m]);
                        pc[m] += 32;
                        ps[m] += 32;
                    }
                } else {  // fragments 0 .. RG-2 of the next entry
#pragma unroll
                    for (int m = 0; m < MA; ++m) {
                        ring[(u + RG - 1) % RG][m].c = ldg_stream(nc + aoff[m] + (u - 1) * 32);
                        ring[(u + RG - 1) % RG][m].s = ldg_stream_f64(ns + aoff[m] + (u - 1) * 32);
                    }
                }
                if (!LAST || u + 1 < RG) {  // B fragments of k-tile kt + 1 (none after the last k-tile of the stage)
#pragma unroll
                    for (int c = 0; c < NCW; ++c) {
                        bfrag[(u + 1) & 1][c].c = bck[((u + 1) * NCO + c) * 32];
                        bfrag[(u + 1) & 1][c].s = bsk[((u + 1) * NCO + c) * 32];
                    }
                    if constexpr (SPLIT) {
                        bfrag[(u + 1) & 1][NCW].c = shck[(u + 1) * 32];
                        bfrag[(u + 1) & 1][NCW].s = shsk[(u + 1) * 32];
                    }
                }
                Frag3 a_s[MS];
                if constexpr (NCW == 0) {
                    Frag3 a_none[MR];  // no own tiles: never read
#pragma unroll
                    for (int mm = 0; mm < MS; ++mm) a_s[mm] = ring[u][mm];
                    mma_block3<MR, NCW, MS, SPLIT>(acc, accs, a_none, a_s, bfrag[u & 1]);
                } else {
                    if constexpr (SPLIT) {
#pragma unroll
                        for (int mm = 0; mm < MS; ++mm) a_s[mm] = rank ? ring[u][MS + mm] : ring[u][mm];
                    }
                    mma_block3<MR, NCW, MS, SPLIT>(acc, accs, ring[u], a_s, bfrag[u & 1]);
                }
            }
        };
#pragma unroll 1
        for (int kt0 = 0; kt0 < KT - RG; kt0 += RG) k_block(std::false_type{}, kt0);
        k_block(std::true_type{}, KT - RG);
#pragma unroll
        for (int m = 0; m < MA; ++m) {  // the ring now holds fragments 0 .. RG-2 of the next entry
            pc[m] = nc + aoff[m] + (RG - 1) * 32;
            ps[m] = ns + aoff[m] + (RG - 1) * 32;
        }

        // ---- epilogue ----
        const StageCoef sc(stage, h);
        if constexpr (SPLIT) {
            // shared octet: goes to the OTHER buffer here and in the peer CTA, no hazard with the reads above
            double2* dc = sh_c + (cur ^ 1) * sh_elems;
            double* ds = sh_s + (cur ^ 1) * sh_elems;
#pragma unroll
            for (int mm = 0; mm < MS; ++mm)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double k_r = accs.re(mm, 0, i), k_i = accs.im(mm, 0, i);
                    double2* sl = slab + ((MR * NCW + mm) * 2 + i) * nthr + tid;
                    const double2 yv = *sl;
                    ksr[mm][i] = sc.keep * ksr[mm][i] + sc.wk * k_r;
                    ksi[mm][i] = sc.keep * ksi[mm][i] + sc.wk * k_i;
                    const double v_r = sc.last ? ksr[mm][i] : k_r, v_i = sc.last ? ksi[mm][i] : k_i;
                    const double2 nxt = make_double2(yv.x + sc.astep * v_r, yv.y + sc.astep * v_i);
                    if (sc.last) *sl = nxt;
                    if (svalid[mm]) {
                        const int pos = yin_pos(1, rts[mm], g, 0, 2 * q + i);
                        const double sum = nxt.x + nxt.y;
                        dc[pos] = nxt;
                        ds[pos] = sum;
                        st_async_peer(dc + pos, mbar + (sidx & 1), rank ^ 1u, nxt);
                        st_async_peer_f64(ds + pos, mbar + (sidx & 1), rank ^ 1u, sum);
                    }
                }
            accs.zero();
        }
        // own tiles: combine into registers first (the accumulators die here), then wait until every warp has
        // finished reading the single-buffered stage planes, then overwrite them
        double2 nx[MR][NC1][2];
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            double2 yv[NC1][2];
#pragma unroll
            for (int c = 0; c < NCW; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i) yv[c][i] = slab[((m * NCW + c) * 2 + i) * nthr + tid];
#pragma unroll
            for (int c = 0; c < NCW; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double k_r = acc.re(m, c, i), k_i = acc.im(m, c, i);
                    kr[m][c][i] = sc.keep * kr[m][c][i] + sc.wk * k_r;
                    ki[m][c][i] = sc.keep * ki[m][c][i] + sc.wk * k_i;
                    const double v_r = sc.last ? kr[m][c][i] : k_r, v_i = sc.last ? ki[m][c][i] : k_i;
                    nx[m][c][i] = make_double2(yv[c][i].x + sc.astep * v_r, yv[c][i].y + sc.astep * v_i);
                    if (sc.last) slab[((m * NCW + c) * 2 + i) * nthr + tid] = nx[m][c][i];
                }
        }
        acc.zero();
        __syncthreads();
#pragma unroll
        for (int m = 0; m < MR; ++m)
#pragma unroll
            for (int c = 0; c < NCW; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    if (mvalid[m]) {
                        const int pos = yin_pos(NCO, rt[m], g, ct0 + c, 2 * q + i);
                        own_c[pos] = nx[m][c][i];
                        own_s[pos] = nx[m][c][i].x + nx[m][c][i].y;
                    }
        cur ^= 1;
        __syncthreads();
        if (SPLIT) mbar_wait(mbar + (sidx & 1), (sidx >> 1) & 1);
    }
    if (SPLIT) cluster_barrier();

    // ---- store y ----
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = colw + 8 * c + 2 * q + i;
                if (mvalid[m] && row < n && col < B) y[(size_t)row * ldy + col] = slab[((m * NCW + c) * 2 + i) * nthr + tid];
            }
    if (SPLIT) {
#pragma unroll
        for (int mm = 0; mm < MS; ++mm)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rts[mm] + g;
                const int col = cols + 2 * q + i;
                if (svalid[mm] && row < n && col < B) y[(size_t)row * ldy + col] = slab[((MR * NCW + mm) * 2 + i) * nthr + tid];
            }
    }
}

// ------------------------------------------------------------------------------------------------
// sweep mode: per-column signal values
// ------------------------------------------------------------------------------------------------
template <int MR, int NCW>
__global__ void __launch_bounds__(256, 1)
rk4_sweep_kernel(Geometry geo, int K, int B, int S, const double2* __restrict__ stat /*packed or null*/,
                 const double2* __restrict__ ops /*[K] packed*/, const double* __restrict__ coeff /*[2S+1][K][ldc]*/,
                 int ldc, const double* __restrict__ mu, const double* __restrict__ times /*[2S+1]*/, double h,
                 double2* __restrict__ y, int ldy) {
    extern __shared__ __align__(16) double2 sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int swl = frag_swizzle(lane);
    const int wr = warp % geo.WR, wc = warp / geo.WR;
    const int KT = geo.KT, NCT = geo.NCT, n = geo.n;
    const size_t entry_elems = (size_t)geo.npad * KT * 4;
    const int yin_elems = KT * NCT * 32;
    const int nthr = blockDim.x;
    const int ncols = 8 * NCT;
    const int yst_off = 2 * yin_elems;                                              // [MR*NCW*2][nthr]
    double* scoef = reinterpret_cast<double*>(sm + yst_off + MR * NCW * 2 * nthr);  // [K][ncols]
    const int col0 = blockIdx.x * ncols;
    const bool framed = (mu != nullptr);

    int rt[MR], rtl[MR];
    bool mvalid[MR];
    double mu_row[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        rt[m] = wr + geo.WR * m;
        mvalid[m] = rt[m] < geo.RT;
        rtl[m] = mvalid[m] ? rt[m] : geo.RT - 1;
        const int row = 8 * rt[m] + g;
        mu_row[m] = (framed && mvalid[m] && row < n) ? mu[row] : 0.0;
    }

    // phases of this thread's rows at the current stage time: p = exp(-i mu t)
    double2 ph[MR];
    {
        const double t0 = framed ? times[0] : 0.0;
#pragma unroll
        for (int m = 0; m < MR; ++m) ph[m] = framed ? frame_phase(mu_row[m], t0) : make_double2(1.0, 0.0);
    }

    for (int i = tid; i < 2 * yin_elems; i += nthr) sm[i] = make_double2(0.0, 0.0);
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int ct = wc * NCW + c;
                const int col = col0 + 8 * ct + 2 * q + i;
                double2 v = make_double2(0.0, 0.0);
                if (mvalid[m] && row < n && col < B) v = y[(size_t)row * ldy + col];
                sm[yst_off + ((m * NCW + c) * 2 + i) * nthr + tid] = v;
                if (mvalid[m]) sm[yin_pos(NCT, rt[m], g, ct, 2 * q + i)] = cmul(ph[m], v);  // pre-phase
            }

    Accum<MR, NCW, false> acc;
    acc.zero();
    Accum<1, 1, false> acc_unused;
    double kr[MR][NCW][2], ki[MR][NCW][2];
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) kr[m][c][0] = kr[m][c][1] = ki[m][c][0] = ki[m][c][1] = 0.0;

    const int has_static = stat != nullptr ? 1 : 0;
    const int J = K + has_static;  // operator passes per k-tile; pass 0 = static operator if present
    // k-tiles that hold data: ceil(n / 4) <= KT.  The pass list (kt, j) is padded to a multiple of RING with passes
    // over k-tile KTE, which exists and is all zero whenever KTE < KT (and KTE == KT makes KTE * J a multiple
    // of RING already) -- n = 81 runs 21 k-tiles instead of the 24 the fragment ring is padded to.
    const int KTE = (n + 3) >> 2;
    const int total = (KTE * J + RING - 1) / RING * RING;

    size_t aoff[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) aoff[m] = (size_t)rtl[m] * KT * 32 + lane;
    auto a_src = [&](int kt, int j) -> const double2* {
        const double2* base = (has_static && j == 0) ? stat : ops + (size_t)(j - has_static) * entry_elems;
        return base + (size_t)kt * 32;
    };

    // ring primed with the first RING-1 (kt, j) passes; (pk, pj) = next pass to prefetch
    double2 ring[RING][MR];
    int pk = 0, pj = 0, pit = 0;  // pit = index of the next prefetched pass within the stage
    auto advance_prefetch = [&]() {
        if (++pit == total) {
            pit = 0;
            pk = 0;
            pj = 0;
        } else if (++pj == J) {
            pj = 0;
            ++pk;
        }
    };
#pragma unroll
    for (int u = 0; u < RING - 1; ++u) {
        const double2* src = a_src(pk, pj);
#pragma unroll
        for (int m = 0; m < MR; ++m) ring[u][m] = __ldg(src + aoff[m]);
        advance_prefetch();
    }

    int cur = 0;
    const int total_stages = 4 * S;
#pragma unroll 1
    for (int sidx = 0; sidx < total_stages; ++sidx) {
        const int step = sidx >> 2, stage = sidx & 3;
        const int entry = 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
        // signal values of this CTA's columns at this stage time
        for (int idx = tid; idx < K * ncols; idx += nthr) {
            const int j = idx / ncols, cc = idx - j * ncols;
            const int col = col0 + cc;
            scoef[idx] = col < B ? coeff[((size_t)entry * K + j) * ldc + col] : 0.0;
        }
        __syncthreads();  // scoef + previous epilogue's stage-buffer writes visible

        const int ybase = cur * yin_elems + (wc * NCW) * 32;
        const double* csrc = scoef + (wc * NCW) * 8 + g;  // B-fragment lane holds column 8 ct + g

        int kt = 0, j = 0;
        double2 b[NCW];
#pragma unroll 1
        for (int it0 = 0; it0 < total; it0 += RING) {
#pragma unroll
            for (int u = 0; u < RING; ++u) {
                {   // prefetch pass it + RING-1 (wraps to the start: the operators are time independent)
                    const double2* src = a_src(pk, pj);
#pragma unroll
                    for (int m = 0; m < MR; ++m) ring[(u + RING - 1) % RING][m] = __ldg(src + aoff[m]);
                    advance_prefetch();
                }
                if (j == 0) {
#pragma unroll
                    for (int c = 0; c < NCW; ++c) b[c] = sm[ybase + (kt * NCT + c) * 32 + swl];
                }
                const int sig = j - has_static;  // -1 -> static operator, coefficient 1
                double2 bs[NCW];
#pragma unroll
                for (int c = 0; c < NCW; ++c) {
                    const double s = sig < 0 ? 1.0 : csrc[sig * ncols + c * 8];
                    bs[c] = make_double2(b[c].x * s, b[c].y * s);
                }
                {
                    const double2 a_unused[1] = {};
                    mma_block<MR, NCW, 1, false>(acc, acc_unused, ring[u], a_unused, bs);
                }
                if (++j == J) { j = 0; ++kt; }
            }
        }

        // ---- epilogue: post-phase conj(p(t_stage)) on k; pre-phase p(t_next) on the next stage input ----
        double2 ph_next[MR];
        {
            const int nstage = (stage + 1) & 3, nstep = step + (stage == 3 ? 1 : 0);
            const int next_entry = (sidx + 1 < total_stages) ? 2 * nstep + (nstage == 0 ? 0 : (nstage == 3 ? 2 : 1)) : entry;
            const double tn = framed ? times[next_entry] : 0.0;
#pragma unroll
            for (int m = 0; m < MR; ++m)
                ph_next[m] = (framed && next_entry != entry) ? frame_phase(mu_row[m], tn) : ph[m];
        }
        const StageCoef sc(stage, h);
        const int ydst = (cur ^ 1) * yin_elems;
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            double2 yv[NCW][2];  // all y-slab loads of the row tile ahead of its stores (see rk4_shared_kernel)
#pragma unroll
            for (int c = 0; c < NCW; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i) yv[c][i] = sm[yst_off + ((m * NCW + c) * 2 + i) * nthr + tid];
#pragma unroll
            for (int c = 0; c < NCW; ++c)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double2 k = cmul_conj_a(ph[m], make_double2(acc.re(m, c, i), acc.im(m, c, i)));
                    kr[m][c][i] = sc.keep * kr[m][c][i] + sc.wk * k.x;
                    ki[m][c][i] = sc.keep * ki[m][c][i] + sc.wk * k.y;
                    const double v_r = sc.last ? kr[m][c][i] : k.x, v_i = sc.last ? ki[m][c][i] : k.y;
                    const double2 nxt = make_double2(yv[c][i].x + sc.astep * v_r, yv[c][i].y + sc.astep * v_i);
                    if (sc.last) sm[yst_off + ((m * NCW + c) * 2 + i) * nthr + tid] = nxt;
                    if (mvalid[m]) sm[ydst + yin_pos(NCT, rt[m], g, wc * NCW + c, 2 * q + i)] = cmul(ph_next[m], nxt);
                }
        }
#pragma unroll
        for (int m = 0; m < MR; ++m) ph[m] = ph_next[m];
        acc.zero();
        cur ^= 1;
        __syncthreads();  // all warps done reading scoef / the old stage buffer before they are rewritten
    }

#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = col0 + 8 * (wc * NCW + c) + 2 * q + i;
                if (mvalid[m] && row < n && col < B)
                    y[(size_t)row * ldy + col] = sm[yst_off + ((m * NCW + c) * 2 + i) * nthr + tid];
            }
}

// ------------------------------------------------------------------------------------------------
// host-side configuration
// ------------------------------------------------------------------------------------------------
struct Config {
    Geometry geo;
    int MR, NCW, threads;
    size_t smem;
    int grid;
    bool split;  // 2-CTA clusters, 2 NCW + 1 column tiles per cluster (shared-signal kernel only)
    bool m3;     // rk4_shared3m_kernel (needs a QDB_LAYOUT_PACKED3M table); geo.NCT = own tiles only
};

constexpr int kMaxFusedNpad = 256;
constexpr size_t kSmemLimit = 227 * 1024;

bool pick_config(int n, int B, int K_sweep /*0 for shared*/, Config& cfg, double* cost_out = nullptr) {
    const int SMS = sm_count();
    const int npad = round_up8(n);
    if (npad > kMaxFusedNpad || n < 1) return false;
    Geometry geo;
    geo.n = n;
    geo.npad = npad;
    geo.KT = round_up16(n) / 4;
    geo.RT = npad / 8;
    const int CT = (B + 7) / 8;  // column tiles in the batch
    int WR, WC, MR;
    if (geo.RT >= 8) {
        WR = 8;
        WC = 1;
        MR = (geo.RT + 7) / 8;
        // fewer wasted row slots with 4 row-warps x 2 column-warps?
        const int MR4 = (geo.RT + 3) / 4;
        if (MR4 <= 4 && MR4 * 4 < MR * 8 && CT >= 2 * SMS) {
            WR = 4;
            WC = 2;
            MR = MR4;
        }
    } else {
        WR = 1;
        while (WR < geo.RT) WR *= 2;
        MR = 1;
        // small problems: 4 warps per CTA so that more CTAs exist; else fill 8 warps with columns
        const int warps = (CT >= 2 * SMS * (8 / WR)) ? 8 : (WR > 4 ? 8 : 4);
        WC = warps / WR;
        if (WC < 1) WC = 1;
    }
    int NCWmax = MR <= 2 ? 4 : 2;
    // experiment hook: QDB_FORCE_CFG="WR,WC,MR,NCW" pins the tiling (profiling only)
    if (const char* force = getenv("QDB_FORCE_CFG")) {
        int fwr, fwc, fmr, fncw;
        if (sscanf(force, "%d,%d,%d,%d", &fwr, &fwc, &fmr, &fncw) == 4 && fwr * fmr >= geo.RT && fwr * fwc <= 8) {
            WR = fwr;
            WC = fwc;
            MR = fmr;
            NCWmax = fncw;
        }
    }
    // The busiest SM runs ceil(ctas / #SMs) CTAs one after the other, each costing a fixed part (epilogue,
    // barrier, pipeline fill; also the poorer A-fragment reuse of narrow tiles) plus a part proportional to
    // its column tiles.  Measured at n=128 (profiles/probe/cfg_sweep.py): 28.0 / 41.2 / 75.4 us per step for
    // NCW = 1 / 2 / 4, i.e. about 9.5 + 16.5 NCW.  Ties go to the wider tile.
    constexpr double kFixedTiles = 0.6;
    double best_cost = -1;
    bool found = false;
    for (int NCW = NCWmax; NCW >= 1; NCW /= 2) {
        Geometry g2 = geo;
        g2.WR = WR;
        g2.WC = WC;
        g2.NCT = NCW * WC;
        const int threads = 32 * WR * WC;
        size_t smem = (size_t)2 * g2.KT * g2.NCT * 32 * sizeof(double2) + (size_t)MR * NCW * 2 * threads * sizeof(double2);
        if (K_sweep > 0) smem += (size_t)K_sweep * 8 * g2.NCT * sizeof(double);
        if (smem > kSmemLimit) continue;
        const int ctas = (CT + g2.NCT - 1) / g2.NCT;
        const double cost = (double)((ctas + SMS - 1) / SMS) * (kFixedTiles + NCW);
        if (!found || cost < best_cost) {
            found = true;
            best_cost = cost;
            cfg.geo = g2;
            cfg.MR = MR;
            cfg.NCW = NCW;
            cfg.threads = threads;
            cfg.smem = smem;
            cfg.grid = ctas;
            cfg.split = false;
        }
    }
    // Split candidates (shared-signal kernel, all warps along rows, two row tiles per warp): a 2-CTA cluster
    // owns 2 NCW + 1 column tiles, the busiest SM computes NCW + 1/2 tiles per row tile and wave.
    const char* nosplit = getenv("QDB_NO_SPLIT");
    if (found && K_sweep == 0 && WC == 1 && MR == 2 && SMS >= 2 && !(nosplit && nosplit[0] == '1')) {
        for (int NCW = 3; NCW >= 1; --NCW) {
            Geometry g2 = geo;
            g2.WR = WR;
            g2.WC = 1;
            g2.NCT = NCW + 1;
            const int threads = 32 * WR;
            const size_t smem = (size_t)2 * g2.KT * g2.NCT * 32 * sizeof(double2) +
                                (size_t)(MR * NCW + MR / 2) * 2 * threads * sizeof(double2) + 16 /*mbarriers*/;
            if (smem > kSmemLimit) continue;
            const int clusters = (CT + 2 * NCW) / (2 * NCW + 1);
            const int slots = SMS / 2;
            const double cost = (double)((clusters + slots - 1) / slots) * (kFixedTiles + NCW + 0.5);
            if (cost < best_cost) {
                best_cost = cost;
                cfg.geo = g2;
                cfg.MR = MR;
                cfg.NCW = NCW;
                cfg.threads = threads;
                cfg.smem = smem;
                cfg.grid = 2 * clusters;
                cfg.split = true;
            }
        }
    }
    cfg.m3 = false;
    if (cost_out) *cost_out = best_cost;
    return found;
}

// 3M kernel (rk4_shared3m_kernel): 8 row warps, MR = ceil(RT / 8) row tiles per warp, at most 7 tiles per warp
// (three accumulators per tile).  Cost in the units of pick_config: a 3M column tile costs about 0.78 of a 4M
// one (3 of 4 DMMAs plus the shared per-k-tile overhead), the second barrier adds a little to the fixed part.
bool pick_config3m(int n, int B, Config& cfg, double* cost_out) {
    const int SMS = sm_count();
    const int npad = round_up8(n);
    if (npad > kMaxFusedNpad || n < 1) return false;
    Geometry geo;
    geo.n = n;
    geo.npad = npad;
    geo.KT = round_up16(n) / 4;
    geo.RT = npad / 8;
    if (geo.RT < 8) return false;  // small operators stay on the 4M kernel
    const int MR = (geo.RT + 7) / 8;
    // measured (profiles/probe/m3_check.py): 3M wins 15-25 % for one or two row tiles per warp; with three or
    // four the narrow column tiles it is left with (7 accumulator tiles per warp at most) give the gain back
    if (MR > 2) return false;
    const int CT = (B + 7) / 8;
    const int threads = 256;
    constexpr double kFixed = 0.7, kTile = 0.78;
    bool found = false;
    double best = 0;
    auto consider = [&](int NCW, bool split) {
        Geometry g2 = geo;
        g2.WR = 8;
        g2.WC = 1;
        g2.NCT = NCW;  // own tiles
        const int MS = split ? MR / 2 : 0;
        if (MR * NCW + MS > 7) return;
        const size_t smem = (size_t)g2.KT * NCW * 32 * 24 + (split ? (size_t)2 * g2.KT * 32 * 24 : 0) +
                            (size_t)(MR * NCW + MS) * 2 * threads * sizeof(double2) + 16;
        if (smem > kSmemLimit) return;
        int ctas;
        double cost;
        if (split) {
            const int clusters = (CT + 2 * NCW) / (2 * NCW + 1);
            const int slots = SMS / 2;
            ctas = 2 * clusters;
            cost = (double)((clusters + slots - 1) / slots) * (kFixed + kTile * (NCW + 0.5));
        } else {
            ctas = (CT + NCW - 1) / NCW;
            cost = (double)((ctas + SMS - 1) / SMS) * (kFixed + kTile * NCW);
        }
        if (!found || cost < best) {
            found = true;
            best = cost;
            cfg.geo = g2;
            cfg.MR = MR;
            cfg.NCW = NCW;
            cfg.threads = threads;
            cfg.smem = smem;
            cfg.grid = ctas;
            cfg.split = split;
            cfg.m3 = true;
        }
    };
    const char* nosplit = getenv("QDB_NO_SPLIT");
    const bool allow_split = MR == 2 && SMS >= 2 && !(nosplit && nosplit[0] == '1');
    for (int NCW = (MR == 1 ? 4 : 3); NCW >= 1; --NCW) {
        consider(NCW, false);
        if (allow_split) consider(NCW, true);
    }
    if (allow_split) consider(0, true);  // pure row split: one octet per cluster (small batches)
    if (found && cost_out) *cost_out = best;
    return found;
}

// Best tiling for the shared-signal solve: 3M when it is available and predicted faster (QDB_NO_3M=1 pins 4M).
bool pick_config_shared(int n, int B, Config& cfg) {
    double c4 = 0, c3 = 0;
    Config cfg4, cfg3;
    const bool ok4 = pick_config(n, B, 0, cfg4, &c4);
    const char* no3 = getenv("QDB_NO_3M");
    const bool ok3 = !(no3 && no3[0] == '1') && pick_config3m(n, B, cfg3, &c3);
    if (ok3 && (!ok4 || c3 < c4)) {
        cfg = cfg3;
        return true;
    }
    if (ok4) cfg = cfg4;
    return ok4;
}

template <int MR, int NCW>
int launch_shared_t(const Config& cfg, int B, int S, const double2* gen, double h, double2* y, int ldy, cudaStream_t st) {
    QDB_CUDA(cudaFuncSetAttribute(rk4_shared_kernel<MR, NCW, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    rk4_shared_kernel<MR, NCW, false, 0><<<cfg.grid, cfg.threads, cfg.smem, st>>>(cfg.geo, B, S, gen, h, y, ldy);
    QDB_LAUNCH_CHECK("rk4_shared_kernel");
    return QDB_OK;
}

template <int MR, int NCW, int SKT>
int launch_split_t(const Config& cfg, int B, int S, const double2* gen, double h, double2* y, int ldy, cudaStream_t st) {
    auto kern = rk4_shared_kernel<MR, NCW, true, SKT>;
    QDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(cfg.grid);
    lc.blockDim = dim3(cfg.threads);
    lc.dynamicSmemBytes = cfg.smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    QDB_CUDA(cudaLaunchKernelEx(&lc, kern, cfg.geo, B, S, gen, h, y, ldy));
    QDB_LAUNCH_CHECK("rk4_shared_kernel<split>");
    return QDB_OK;
}

template <int MR, int NCW>
int launch_sweep_t(const Config& cfg, int K, int B, int S, const double2* stat, const double2* ops, const double* coeff, int ldc,
                   const double* mu, const double* times, double h, double2* y, int ldy, cudaStream_t st) {
    QDB_CUDA(cudaFuncSetAttribute(rk4_sweep_kernel<MR, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    rk4_sweep_kernel<MR, NCW><<<cfg.grid, cfg.threads, cfg.smem, st>>>(cfg.geo, K, B, S, stat, ops, coeff, ldc, mu, times, h, y, ldy);
    QDB_LAUNCH_CHECK("rk4_sweep_kernel");
    return QDB_OK;
}

template <int MR, int NCW, bool SPLIT, int SKT>
int launch_3m_t(const Config& cfg, int B, int S, const double2* gen, double h, double2* y, int ldy, cudaStream_t st) {
    auto kern = rk4_shared3m_kernel<MR, NCW, SPLIT, SKT>;
    QDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(cfg.grid);
    lc.blockDim = dim3(cfg.threads);
    lc.dynamicSmemBytes = cfg.smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SPLIT ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    QDB_CUDA(cudaLaunchKernelEx(&lc, kern, cfg.geo, B, S, gen, h, y, ldy));
    QDB_LAUNCH_CHECK("rk4_shared3m_kernel");
    return QDB_OK;
}

// QDB_ROWSPLIT_OLD=1 pins rk4_shared3m_kernel<2,0,split> for small batches at n = 121..128 (the bit-level reference of
// rk4_rowsplit3m_kernel in the tests: same products in the same order per accumulator, k halves swapped on rank 1)
bool rowsplit_enabled() {
    const char* e = getenv("QDB_ROWSPLIT_OLD");
    return !(e && e[0] == '1');
}

#define QDB_DISPATCH(MRv, NCWv, CALL)                \
    if (cfg.MR == MRv && cfg.NCW == NCWv) return CALL

}  // namespace

bool rk4_fused_supported(int n) { return n >= 1 && round_up8(n) <= kMaxFusedNpad; }

bool rk4_fused_tiling(int n, int B, int sweep_K, int* out) {
    if (sweep_K > 0 && rk4_sweepf_selected(n, sweep_K, B, rk4_sweep_small_supported(n, sweep_K, true)))
        return rk4_sweepf_tiling(n, B, sweep_K, out);
    Config cfg;
    const bool ok = sweep_K > 0 ? pick_config(n, B, sweep_K, cfg)
                                : (sweep_K < 0 ? pick_config(n, B, 0, cfg) : pick_config_shared(n, B, cfg));
    if (!ok) return false;
    out[0] = cfg.geo.WR;
    out[1] = cfg.geo.WC;
    out[2] = cfg.MR;
    out[3] = cfg.NCW;
    out[4] = cfg.split ? 1 : 0;
    out[5] = cfg.grid;
    out[6] = cfg.threads;
    out[7] = (int)cfg.smem;
    out[8] = cfg.m3 ? 1 : 0;
    return true;
}

// table_layout: QDB_LAYOUT_PACKED -> 4M kernels, QDB_LAYOUT_PACKED3M -> 3M kernel (see rk4_fused_table_layout)
int launch_rk4_fused_shared(int n, int B, int S, const double2* gen_table, int table_layout, double h, double2* y, int ldy,
                            cudaStream_t st) {
    Config cfg;
    bool ok;
    if (table_layout == QDB_LAYOUT_PACKED3M) {
        ok = pick_config3m(n, B, cfg, nullptr);
    } else {
        ok = pick_config(n, B, 0, cfg);
    }
    if (!ok) {
        set_error("rk4 fused: unsupported shape n=%d B=%d for table layout %d", n, B, table_layout);
        return QDB_E_UNSUPPORTED;
    }
#define ARGS cfg, B, S, gen_table, h, y, ldy, st
    if (cfg.m3) {
        const bool static128 = cfg.geo.KT == 32 && cfg.geo.RT == 16 && cfg.geo.npad == 128;
        if (cfg.split) {
            if (static128) QDB_DISPATCH(2, 3, (launch_3m_t<2, 3, true, 32>(ARGS)));
            if (static128 && cfg.NCW == 0 && rowsplit_enabled()) return launch_rk4_rowsplit3m(n, B, S, gen_table, h, y, ldy, st);
            if (static128) QDB_DISPATCH(2, 0, (launch_3m_t<2, 0, true, 32>(ARGS)));
            QDB_DISPATCH(2, 0, (launch_3m_t<2, 0, true, 0>(ARGS)));
            QDB_DISPATCH(2, 1, (launch_3m_t<2, 1, true, 0>(ARGS)));
            QDB_DISPATCH(2, 2, (launch_3m_t<2, 2, true, 0>(ARGS)));
            QDB_DISPATCH(2, 3, (launch_3m_t<2, 3, true, 0>(ARGS)));
        } else {
            if (static128) QDB_DISPATCH(2, 1, (launch_3m_t<2, 1, false, 32>(ARGS)));
            if (static128) QDB_DISPATCH(2, 2, (launch_3m_t<2, 2, false, 32>(ARGS)));
            QDB_DISPATCH(1, 1, (launch_3m_t<1, 1, false, 0>(ARGS)));
            QDB_DISPATCH(1, 2, (launch_3m_t<1, 2, false, 0>(ARGS)));
            QDB_DISPATCH(1, 3, (launch_3m_t<1, 3, false, 0>(ARGS)));
            QDB_DISPATCH(1, 4, (launch_3m_t<1, 4, false, 0>(ARGS)));
            QDB_DISPATCH(2, 1, (launch_3m_t<2, 1, false, 0>(ARGS)));
            QDB_DISPATCH(2, 2, (launch_3m_t<2, 2, false, 0>(ARGS)));
            QDB_DISPATCH(2, 3, (launch_3m_t<2, 3, false, 0>(ARGS)));
        }
        set_error("rk4 fused 3M: no kernel for MR=%d NCW=%d split=%d", cfg.MR, cfg.NCW, (int)cfg.split);
        return QDB_E_UNSUPPORTED;
    }
    if (cfg.split) {
        // headline geometry (n = 121..128: 32 k-tiles, 16 row tiles) compiled with static geometry
        const bool static128 = cfg.geo.KT == 32 && cfg.geo.RT == 16 && cfg.geo.WR == 8 && cfg.geo.npad == 128;
        if (static128) QDB_DISPATCH(2, 3, (launch_split_t<2, 3, 32>(ARGS)));
        QDB_DISPATCH(2, 1, (launch_split_t<2, 1, 0>(ARGS)));
        QDB_DISPATCH(2, 2, (launch_split_t<2, 2, 0>(ARGS)));
        QDB_DISPATCH(2, 3, (launch_split_t<2, 3, 0>(ARGS)));
    }
    QDB_DISPATCH(1, 1, (launch_shared_t<1, 1>(ARGS)));
    QDB_DISPATCH(1, 2, (launch_shared_t<1, 2>(ARGS)));
    QDB_DISPATCH(1, 4, (launch_shared_t<1, 4>(ARGS)));
    QDB_DISPATCH(2, 1, (launch_shared_t<2, 1>(ARGS)));
    QDB_DISPATCH(2, 2, (launch_shared_t<2, 2>(ARGS)));
    QDB_DISPATCH(2, 4, (launch_shared_t<2, 4>(ARGS)));
    QDB_DISPATCH(3, 1, (launch_shared_t<3, 1>(ARGS)));
    QDB_DISPATCH(3, 2, (launch_shared_t<3, 2>(ARGS)));
    QDB_DISPATCH(4, 1, (launch_shared_t<4, 1>(ARGS)));
    QDB_DISPATCH(4, 2, (launch_shared_t<4, 2>(ARGS)));
#undef ARGS
    set_error("rk4 fused: no kernel for MR=%d NCW=%d", cfg.MR, cfg.NCW);
    return QDB_E_UNSUPPORTED;
}

// table layout the shared-signal solve should be fed for this shape
int rk4_fused_table_layout(int n, int B) {
    Config cfg;
    if (!pick_config_shared(n, B, cfg)) return QDB_LAYOUT_PACKED;
    return cfg.m3 ? QDB_LAYOUT_PACKED3M : QDB_LAYOUT_PACKED;
}

int launch_rk4_fused_sweep(int n, int K, int B, int S, const double2* stat_packed, const double2* ops_packed,
                           const double* coeff, int ldc,
                           const double* mu, const double* times_dev, double h, double2* y, int ldy, void* ws, cudaStream_t st) {
    // K >= 3: generator formed per column on the tensor pipe (2 Kpad + 4 FMAs per element instead of 4 (K + 1))
    const bool small_ok = rk4_sweep_small_supported(n, K, stat_packed != nullptr);
    if (ws != nullptr && rk4_sweepf_selected(n, K, B, small_ok))
        return launch_rk4_sweepf(n, K, B, S, stat_packed, ops_packed, coeff, ldc, mu, times_dev, h, y, ldy, ws, st);
    // small operators: operators resident in shared memory, pre-scaled stage vectors, operator sum split over warps
    if (small_ok)
        return launch_rk4_sweep_small(n, K, B, S, stat_packed, ops_packed, coeff, ldc, mu, times_dev, h, y, ldy, st);
    Config cfg;
    if (!pick_config(n, B, K > 0 ? K : 1, cfg)) {
        set_error("rk4 sweep: unsupported shape n=%d B=%d K=%d", n, B, K);
        return QDB_E_UNSUPPORTED;
    }
#define ARGS cfg, K, B, S, stat_packed, ops_packed, coeff, ldc, mu, times_dev, h, y, ldy, st
    QDB_DISPATCH(1, 1, (launch_sweep_t<1, 1>(ARGS)));
    QDB_DISPATCH(1, 2, (launch_sweep_t<1, 2>(ARGS)));
    QDB_DISPATCH(1, 4, (launch_sweep_t<1, 4>(ARGS)));
    QDB_DISPATCH(2, 1, (launch_sweep_t<2, 1>(ARGS)));
    QDB_DISPATCH(2, 2, (launch_sweep_t<2, 2>(ARGS)));
    QDB_DISPATCH(2, 4, (launch_sweep_t<2, 4>(ARGS)));
    QDB_DISPATCH(3, 1, (launch_sweep_t<3, 1>(ARGS)));
    QDB_DISPATCH(3, 2, (launch_sweep_t<3, 2>(ARGS)));
    QDB_DISPATCH(4, 1, (launch_sweep_t<4, 1>(ARGS)));
    QDB_DISPATCH(4, 2, (launch_sweep_t<4, 2>(ARGS)));
#undef ARGS
    set_error("rk4 sweep: no kernel for MR=%d NCW=%d", cfg.MR, cfg.NCW);
    return QDB_E_UNSUPPORTED;
}

}  // namespace qdb
