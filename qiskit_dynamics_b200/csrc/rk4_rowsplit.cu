// Fused shared-signal RK4 for SMALL batches at the headline dimension (n = 121..128): pure row split.
//
// Below ~600 columns there are fewer column octets than SMs (the per-GPU share of a strong-scaled batch of 4096 on
// 8 GPUs is 512 columns = 64 octets), so a 2-CTA cluster owns ONE octet and each CTA half of its output rows: warp w of
// rank r computes row tile 8 r + w (8 rows x 8 columns, three DMMA accumulators: re*re, im*im, (re+im)(re+im)).
// rk4_shared3m_kernel<2,0,true,32> does the same; the round-1 ncu capture of it (profiles/r02_b_*) shows what a warp
// with 3 DMMAs per k-tile is bound by: not the tensor pipe (55 % of active cycles) and not L2 bandwidth (11 %), but
// the L2 LATENCY of the generator stream (long-scoreboard stalls on the first DMMA of every ring block, 31 % of the
// samples) plus the per-stage exchange: barrier + mbarrier wait, 10 %.  This kernel changes three things:
//
//   * the A-fragment register ring is 16 k-tiles deep (15 ahead ~ 1400 cycles of DMMA issue of the two warps of a
//     sub-partition), still streaming across stage boundaries;
//   * a CTA walks the k dimension OWN HALF FIRST: the rows it computed itself are in its shared memory right after its
//     epilogue, the peer's rows arrive through DSMEM (st.async + complete_tx on this CTA's mbarrier) while the first
//     16 k-tiles are multiplied -- the mbarrier wait sits in the middle of the stage, where it has already completed;
//   * y and the RK4 k-sum live in registers (one row tile per warp), one __syncthreads per stage.
//
// Table layout: QDB_LAYOUT_PACKED3M (complex plane + (re + im) plane), as for rk4_shared3m_kernel.
#include <type_traits>

#include "qdb_common.cuh"
#include "rk4_device.cuh"

namespace qdb {

namespace {

constexpr int KT = 32;    // k-tiles (n padded to 128)
constexpr int HALF = 16;  // k-tiles per rank = ring depth
constexpr int PLANE = KT * 32;

struct F3 {
    double2 c;
    double s;
};

__device__ __forceinline__ unsigned ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ldg_f64(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t map_peer(const void* p, unsigned peer) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(peer));
    return r;
}
__device__ __forceinline__ void st_async_c(uint32_t dst, uint32_t bar, double2 v) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1,%2}, [%3];" ::"r"(dst), "d"(v.x),
                 "d"(v.y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_async_s(uint32_t dst, uint32_t bar, double v) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(dst), "d"(v), "r"(bar)
                 : "memory");
}

// The k loop is scheduled by hand: every load and every DMMA is a volatile asm, so ptxas keeps the written order
// (A fragment 15 k-tiles ahead, B fragment 3 ahead, then the three DMMAs of the k-tile).  Left to itself the compiler
// sinks the shared-memory loads to three DMMAs before their use (ncu: 15 % short-scoreboard stalls).
__device__ __forceinline__ void dmma_v(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds_c(const double2* p) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ double lds_s(const double* p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

// stage-buffer slot of (row tile rt, row g, column cin) of the octet: B-fragment order, k-tile = 2 rt + g / 4
__device__ __forceinline__ int slot(int rt, int g, int cin) {
    return (2 * rt + (g >> 2)) * 32 + frag_swizzle((g & 3) + 4 * cin);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
rk4_rowsplit3m_kernel(int n, int B, int S, const double2* __restrict__ gen, double h, double2* __restrict__ y, int ldy) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double2 (*sh_c)[PLANE] = reinterpret_cast<double2 (*)[PLANE]>(smraw);                        // stage vector of the octet, complex plane, [2]
    double (*sh_s)[PLANE] = reinterpret_cast<double (*)[PLANE]>(smraw + 2 * PLANE * sizeof(double2));  // re + im plane, [2]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smraw + 2 * PLANE * (sizeof(double2) + sizeof(double)));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int swl = frag_swizzle(lane);
    const unsigned rank = ctarank(), peer = rank ^ 1u;
    const int rt = (int)rank * 8 + warp;          // own row tile
    const int rtp = (int)peer * 8 + warp;         // the peer row tile this warp stages at start-up
    const int col0 = 8 * (blockIdx.x >> 1) + 2 * q;
    const size_t entry_elems = (size_t)128 * KT * 4;              // complex plane of one table entry
    const size_t entry_stride = entry_elems + entry_elems / 2;    // + sum plane, in double2 units

    // ---- load y: own rows into registers and the stage buffer, the peer's rows into the stage buffer only ----
    double2 yv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = 8 * rt + g, rowp = 8 * rtp + g, col = col0 + i;
        double2 v = make_double2(0.0, 0.0), vp = make_double2(0.0, 0.0);
        if (row < n && col < B) v = y[(size_t)row * ldy + col];
        if (rowp < n && col < B) vp = y[(size_t)rowp * ldy + col];
        yv[i] = v;
        const int p0 = slot(rt, g, 2 * q + i), p1 = slot(rtp, g, 2 * q + i);
        sh_c[0][p0] = v;
        sh_s[0][p0] = v.x + v.y;
        sh_c[0][p1] = vp;
        sh_s[0][p1] = vp.x + vp.y;
    }
    if (tid == 0) {
        mbar_init(mbar, 1);
        mbar_init(mbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // remote addresses of this thread's two output slots (both buffers) and of the peer's mbarriers
    uint32_t rc[2][2], rs[2][2], rbar[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        rbar[b] = map_peer(mbar + b, peer);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int p0 = slot(rt, g, 2 * q + i);
            rc[b][i] = map_peer(&sh_c[b][p0], peer);
            rs[b][i] = map_peer(&sh_s[b][p0], peer);
        }
    }

    // A-fragment stream: this warp's row tile, k-tiles in the order own half, peer half
    const size_t aoff = (size_t)rt * KT * 32 + lane;
    const int lo = 16 * (int)rank * 32, hi = 16 * (int)peer * 32;  // element offsets of the two k halves
    F3 ring[HALF];
    {
        const double2* c = gen + aoff + lo;
        const double* s = reinterpret_cast<const double*>(gen + entry_elems) + aoff + lo;
#pragma unroll
        for (int u = 0; u < HALF - 1; ++u) {
            ring[u].c = ldg_stream(c + u * 32);
            ring[u].s = ldg_f64(s + u * 32);
        }
    }
    double p0[2] = {0.0, 0.0}, p1[2] = {0.0, 0.0}, p2[2] = {0.0, 0.0};  // re*re, im*im, (re+im)(re+im)
    double ksr[2] = {0.0, 0.0}, ksi[2] = {0.0, 0.0};
    const unsigned tx_bytes = 8u * 64u * 24u;  // the peer's 8 row tiles, complex + sum plane

    cluster_sync_all();  // stage buffers and mbarriers of both CTAs are initialised before any remote store can land

    const int total = 4 * S;
    int cur = 0;
#pragma unroll 1
    for (int sidx = 0; sidx < total; ++sidx) {
        const int step = sidx >> 2, stage = sidx & 3;
        const int entry = 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
        const int nstage = (stage + 1) & 3, nstep = step + (stage == 3 ? 1 : 0);
        const bool final_stage = sidx + 1 == total;
        const int nentry = final_stage ? entry : 2 * nstep + (nstage == 0 ? 0 : (nstage == 3 ? 2 : 1));
        const double2* ec = gen + (size_t)entry * entry_stride + aoff;
        const double* es = reinterpret_cast<const double*>(gen + (size_t)entry * entry_stride + entry_elems) + aoff;
        const double2* nc = gen + (size_t)nentry * entry_stride + aoff + lo;
        const double* ns = reinterpret_cast<const double*>(gen + (size_t)nentry * entry_stride + entry_elems) + aoff + lo;
        if (tid == 0 && !final_stage) mbar_expect_tx(mbar + (sidx & 1), tx_bytes);  // rows of stage sidx + 1 from the peer

        // The 32 k-tiles of the stage as one sequence: positions 0..15 = own half, 16..31 = peer half.  Ring slot p % 16
        // holds the A fragment of position p; position p + 15 is fetched into the slot position p - 1 released.  B fragments
        // run BD positions ahead through a small register ring; the mbarrier wait for the peer's rows sits right before
        // the first B load of the second half, where it has long completed.
        {
            constexpr int BD = 4;
            const double2* bc_lo = &sh_c[cur][lo + swl];
            const double* bs_lo = &sh_s[cur][lo + swl];
            const double2* bc_hi = &sh_c[cur][hi + swl];
            const double* bs_hi = &sh_s[cur][hi + swl];
            F3 b[BD];
#pragma unroll
            for (int j = 0; j < BD - 1; ++j) {
                b[j].c = lds_c(bc_lo + j * 32);
                b[j].s = lds_s(bs_lo + j * 32);
            }
#pragma unroll
            for (int p = 0; p < 2 * HALF; ++p) {
                {   // A prefetch, 15 positions ahead
                    F3& dst = ring[(p + HALF - 1) % HALF];
                    const int t = p + HALF - 1;  // target position: < 32 this entry, else next entry (own half first)
                    if (t < HALF) {
                        dst.c = ldg_stream(ec + lo + t * 32);
                        dst.s = ldg_f64(es + lo + t * 32);
                    } else if (t < 2 * HALF) {
                        dst.c = ldg_stream(ec + hi + (t - HALF) * 32);
                        dst.s = ldg_f64(es + hi + (t - HALF) * 32);
                    } else {
                        dst.c = ldg_stream(nc + (t - 2 * HALF) * 32);
                        dst.s = ldg_f64(ns + (t - 2 * HALF) * 32);
                    }
                }
                {   // B prefetch, BD - 1 positions ahead
                    const int t = p + BD - 1;
                    if (t == HALF && sidx > 0) mbar_wait(mbar + ((sidx - 1) & 1), ((sidx - 1) >> 1) & 1);
                    if (t < HALF) {
                        b[t % BD].c = lds_c(bc_lo + t * 32);
                        b[t % BD].s = lds_s(bs_lo + t * 32);
                    } else if (t < 2 * HALF) {
                        b[t % BD].c = lds_c(bc_hi + (t - HALF) * 32);
                        b[t % BD].s = lds_s(bs_hi + (t - HALF) * 32);
                    }
                }
                const F3& a = ring[p % HALF];
                const F3& bb = b[p % BD];
                dmma_v(p0[0], p0[1], a.c.x, bb.c.x);
                dmma_v(p1[0], p1[1], a.c.y, bb.c.y);
                dmma_v(p2[0], p2[1], a.s, bb.s);
            }
        }

        // ---- epilogue: RK4 stage combine; next stage input to both CTAs ----
        const StageCoef sc(stage, h);
        const int nb = cur ^ 1;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const double k_r = p0[i] - p1[i], k_i = (p2[i] - p0[i]) - p1[i];
            ksr[i] = sc.keep * ksr[i] + sc.wk * k_r;
            ksi[i] = sc.keep * ksi[i] + sc.wk * k_i;
            const double v_r = sc.last ? ksr[i] : k_r, v_i = sc.last ? ksi[i] : k_i;
            const double2 nxt = make_double2(yv[i].x + sc.astep * v_r, yv[i].y + sc.astep * v_i);
            if (sc.last) yv[i] = nxt;
            if (!final_stage) {
                const int ps = slot(rt, g, 2 * q + i);
                const double sum = nxt.x + nxt.y;
                sh_c[nb][ps] = nxt;
                sh_s[nb][ps] = sum;
                const uint32_t bar = (sidx & 1) ? rbar[1] : rbar[0];
                st_async_c(nb ? rc[1][i] : rc[0][i], bar, nxt);
                st_async_s(nb ? rs[1][i] : rs[0][i], bar, sum);
            }
            p0[i] = p1[i] = p2[i] = 0.0;
        }
        cur = nb;
        __syncthreads();  // own rows of the next stage are visible to every warp of this CTA
    }
    cluster_sync_all();  // neither CTA retires while the other could still address its shared memory

#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = 8 * rt + g, col = col0 + i;
        if (row < n && col < B) y[(size_t)row * ldy + col] = yv[i];
    }
}

}  // namespace

// n = 121..128 (32 k-tiles, 16 row tiles), table in QDB_LAYOUT_PACKED3M; one 2-CTA cluster per column octet
int launch_rk4_rowsplit3m(int n, int B, int S, const double2* gen_table, double h, double2* y, int ldy, cudaStream_t st) {
    const int octets = (B + 7) / 8;
    constexpr size_t smem = 2 * PLANE * (sizeof(double2) + sizeof(double)) + 16;
    QDB_CUDA(cudaFuncSetAttribute(rk4_rowsplit3m_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rk4_rowsplit3m_kernel<<<2 * octets, 256, smem, st>>>(n, B, S, gen_table, h, y, ldy);
    QDB_LAUNCH_CHECK("rk4_rowsplit3m_kernel");
    return QDB_OK;
}

}  // namespace qdb
