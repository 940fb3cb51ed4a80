// Fused fixed-step RK4: S steps x 4 stages in ONE launch, state tile resident on chip
// (SURVEY.md 8(a) rows a2 + a3 + a7 + a8 inner loop).
//
// Decomposition.  Every state column evolves independently, so a CTA owns 8*NCT whole columns for
// the entire launch: no inter-CTA communication, HBM traffic = read y once + write y once.
//   * A operand (the generator) streams L2 -> registers directly in DMMA A-fragment order
//     (QDB_LAYOUT_PACKED: one coalesced 512 B LDG.128 per warp per fragment).  Each warp owns
//     distinct row tiles, so there is no intra-CTA reuse that shared-memory staging could exploit.
//   * B operand (the stage vector) lives in shared memory in DMMA B-fragment order, double
//     buffered; the epilogue of stage s writes the stage s+1 input there (XOR-swizzled so both the
//     fragment loads and the scattered epilogue stores are bank-conflict free).
//   * The RK4 accumulator lives in registers, y itself in a thread-private shared-memory slab.
//
// Shared-signal mode: A = precomputed generator table entry G_frame(t_stage) (frame phases folded
// in by generator_kernel).  Sweep mode: A = the K+1 stored operators, the per-column signal value
// scales the B fragment (so the operator sum accumulates in the same DMMA accumulators) and the
// frame phases are applied to B rows on write and to C rows on read.
//
// Roofline: fp64 tensor pipe.  Algorithmic flops per column per step = 4(8n^2 + 12n) + 28n.
#include "qdb_common.cuh"

namespace qdb {

namespace {

constexpr int PF = 4;  // A-fragment prefetch depth (k4-steps)

struct Geometry {
    int n, npad, KT, RT;
    int WR, WC;   // warps along rows / columns
    int NCT;      // column tiles per CTA
};

// position of state element (row, col-in-CTA) in the B-fragment-ordered stage buffer
__device__ __forceinline__ int yin_pos(int NCT, int rt, int g, int ct, int cin) {
    const int kt = 2 * rt + (g >> 2);
    const int lane_b = (g & 3) + 4 * cin;
    return (kt * NCT + ct) * 32 + (lane_b ^ ((g >> 2) << 2));
}

template <int MR, int NCW>
struct Accum {
    double cr[MR][NCW][2], ci[MR][NCW][2];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int m = 0; m < MR; ++m)
#pragma unroll
            for (int c = 0; c < NCW; ++c) cr[m][c][0] = cr[m][c][1] = ci[m][c][0] = ci[m][c][1] = 0.0;
    }
};

template <int MR, int NCW>
__device__ __forceinline__ void mma_block(Accum<MR, NCW>& acc, const double2 (&a)[MR], const double2 (&b)[NCW]) {
    double nai[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) nai[m] = negate(a[m].y);
#pragma unroll
    for (int m = 0; m < MR; ++m) {
#pragma unroll
        for (int c = 0; c < NCW; ++c) {
            dmma(acc.cr[m][c][0], acc.cr[m][c][1], a[m].x, b[c].x);
            dmma(acc.ci[m][c][0], acc.ci[m][c][1], a[m].x, b[c].y);
        }
    }
#pragma unroll
    for (int m = 0; m < MR; ++m) {
#pragma unroll
        for (int c = 0; c < NCW; ++c) {
            dmma(acc.cr[m][c][0], acc.cr[m][c][1], nai[m], b[c].y);
            dmma(acc.ci[m][c][0], acc.ci[m][c][1], a[m].y, b[c].x);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// shared-signal mode
// ------------------------------------------------------------------------------------------------
template <int MR, int NCW>
__global__ void __launch_bounds__(256, 1)
rk4_shared_kernel(Geometry geo, int B, int S, const double2* __restrict__ gen, double h, double2* __restrict__ y,
                  int ldy) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wr = warp % geo.WR, wc = warp / geo.WR;
    const int KT = geo.KT, NCT = geo.NCT, n = geo.n;
    const size_t npad2 = (size_t)geo.npad * geo.npad;
    const int yin_elems = KT * NCT * 32;
    double2* yin[2] = {reinterpret_cast<double2*>(smem_raw), reinterpret_cast<double2*>(smem_raw) + yin_elems};
    double2* yst = reinterpret_cast<double2*>(smem_raw) + 2 * yin_elems;  // [MR*NCW*2][blockDim]
    const int nthr = blockDim.x;
    const int col0 = blockIdx.x * 8 * NCT;

    int rt[MR], rtl[MR];  // row tile owned / row tile loaded (clamped: surplus warps recompute the last tile)
    bool mvalid[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        rt[m] = wr + geo.WR * m;
        mvalid[m] = rt[m] < geo.RT;
        rtl[m] = mvalid[m] ? rt[m] : geo.RT - 1;
    }

    // ---- load y tile: thread-private slab + stage-0 input ----
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
                yst[((m * NCW + c) * 2 + i) * nthr + tid] = v;
                if (mvalid[m]) yin[0][yin_pos(NCT, rt[m], g, ct, 2 * q + i)] = v;
            }
    __syncthreads();

    Accum<MR, NCW> acc;
    acc.zero();
    double kr[MR][NCW][2], ki[MR][NCW][2];  // running k1 + 2 k2 + 2 k3 + k4
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) kr[m][c][0] = kr[m][c][1] = ki[m][c][0] = ki[m][c][1] = 0.0;
    int cur = 0;
    const double h2 = 0.5 * h;
    const double h6 = (1.0 / 6) * h;  // reference: div6 * h * (...)  (fixed_step_solvers.py:60,73)

    for (int step = 0; step < S; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const int entry = 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
            const double2* gsrc = gen + (size_t)entry * npad2 + lane;
            const double2* ysrc = yin[cur] + (wc * NCW) * 32;
            // ---- main loop over k4 tiles, A fragments prefetched PF tiles ahead ----
            double2 abuf[PF][MR];
#pragma unroll
            for (int u = 0; u < PF; ++u)
#pragma unroll
                for (int m = 0; m < MR; ++m)
                    if (u < KT) abuf[u][m] = ldg_stream(gsrc + ((size_t)rtl[m] * KT + u) * 32);
            for (int kt0 = 0; kt0 < KT; kt0 += PF) {
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    const int kt = kt0 + u;
                    if (kt < KT) {
                        double2 a[MR], b[NCW];
#pragma unroll
                        for (int m = 0; m < MR; ++m) a[m] = abuf[u][m];
#pragma unroll
                        for (int m = 0; m < MR; ++m)
                            if (kt + PF < KT) abuf[u][m] = ldg_stream(gsrc + ((size_t)rtl[m] * KT + kt + PF) * 32);
                        const int sw = lane ^ ((kt & 1) << 2);
#pragma unroll
                        for (int c = 0; c < NCW; ++c) b[c] = ysrc[(kt * NCT + c) * 32 + sw];
                        mma_block<MR, NCW>(acc, a, b);
                    }
                }
            }
            // ---- epilogue: RK4 stage combine, write next stage input ----
            double2* ydst = yin[cur ^ 1];
            // k-sum weights 1,2,2,1; next-input step h/2, h/2, h; final update (1/6) h * ksum
            const bool last = (stage == 3);
            const double keep = stage == 0 ? 0.0 : 1.0;
            const double wk = (stage == 1 || stage == 2) ? 2.0 : 1.0;
            const double astep = stage < 2 ? h2 : (stage == 2 ? h : h6);
#pragma unroll
            for (int m = 0; m < MR; ++m) {
#pragma unroll
                for (int c = 0; c < NCW; ++c) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double k_r = acc.cr[m][c][i], k_i = acc.ci[m][c][i];
                        const int slab = ((m * NCW + c) * 2 + i) * nthr + tid;
                        const double2 yv = yst[slab];
                        kr[m][c][i] = keep * kr[m][c][i] + wk * k_r;
                        ki[m][c][i] = keep * ki[m][c][i] + wk * k_i;
                        const double v_r = last ? kr[m][c][i] : k_r, v_i = last ? ki[m][c][i] : k_i;
                        const double2 nxt = make_double2(yv.x + astep * v_r, yv.y + astep * v_i);
                        if (last) yst[slab] = nxt;
                        if (mvalid[m]) ydst[yin_pos(NCT, rt[m], g, wc * NCW + c, 2 * q + i)] = nxt;
                    }
                }
            }
            acc.zero();
            cur ^= 1;
            __syncthreads();
        }
    }

    // ---- store y ----
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = col0 + 8 * (wc * NCW + c) + 2 * q + i;
                if (mvalid[m] && row < n && col < B) y[(size_t)row * ldy + col] = yst[((m * NCW + c) * 2 + i) * nthr + tid];
            }
}

// ------------------------------------------------------------------------------------------------
// sweep mode: per-column signal values
// ------------------------------------------------------------------------------------------------
template <int MR, int NCW>
__global__ void __launch_bounds__(256, 1)
rk4_sweep_kernel(Geometry geo, int K, int B, int S, const double2* __restrict__ stat /*packed or null*/,
                 const double2* __restrict__ ops /*[K] packed*/,
                 const double* __restrict__ coeff /*[2S+1][K][ldc]*/, int ldc, const double* __restrict__ mu,
                 const double* __restrict__ times /*[2S+1]*/, double h, double2* __restrict__ y, int ldy) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wr = warp % geo.WR, wc = warp / geo.WR;
    const int KT = geo.KT, NCT = geo.NCT, n = geo.n;
    const size_t npad2 = (size_t)geo.npad * geo.npad;
    const int yin_elems = KT * NCT * 32;
    const int nthr = blockDim.x;
    const int ncols = 8 * NCT;
    double2* yin[2] = {reinterpret_cast<double2*>(smem_raw), reinterpret_cast<double2*>(smem_raw) + yin_elems};
    double2* yst = reinterpret_cast<double2*>(smem_raw) + 2 * yin_elems;  // [MR*NCW*2][nthr]
    double* scoef = reinterpret_cast<double*>(yst + MR * NCW * 2 * nthr);  // [K][ncols]
    const int col0 = blockIdx.x * ncols;

    int rt[MR], rtl[MR];
    bool mvalid[MR];
    double mu_row[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        rt[m] = wr + geo.WR * m;
        mvalid[m] = rt[m] < geo.RT;
        rtl[m] = mvalid[m] ? rt[m] : geo.RT - 1;
        const int row = 8 * rt[m] + g;
        mu_row[m] = (mu != nullptr && mvalid[m] && row < n) ? mu[row] : 0.0;
    }
    const bool framed = (mu != nullptr);

    // phases of this thread's rows at the current stage time: p = exp(-i mu t)
    double2 ph[MR];
    {
        const double t0 = framed ? times[0] : 0.0;
#pragma unroll
        for (int m = 0; m < MR; ++m) ph[m] = framed ? frame_phase(mu_row[m], t0) : make_double2(1.0, 0.0);
    }

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
                yst[((m * NCW + c) * 2 + i) * nthr + tid] = v;
                if (mvalid[m]) yin[0][yin_pos(NCT, rt[m], g, ct, 2 * q + i)] = cmul(ph[m], v);  // pre-phase
            }

    Accum<MR, NCW> acc;
    acc.zero();
    double kr[MR][NCW][2], ki[MR][NCW][2];
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) kr[m][c][0] = kr[m][c][1] = ki[m][c][0] = ki[m][c][1] = 0.0;
    int cur = 0;
    const double h2 = 0.5 * h;
    const double h6 = (1.0 / 6) * h;
    const int has_static = stat != nullptr ? 1 : 0;
    const int J = K + has_static;  // operator passes per k4 tile; pass 0 = static if present
    const int total = KT * J;

    for (int step = 0; step < S; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            const int entry = 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
            // signal values of this CTA's columns at this stage time
            for (int idx = tid; idx < K * ncols; idx += nthr) {
                const int j = idx / ncols, cc = idx - j * ncols;
                const int col = col0 + cc;
                scoef[idx] = col < B ? coeff[((size_t)entry * K + j) * ldc + col] : 0.0;
            }
            __syncthreads();  // scoef + previous epilogue's yin writes visible

            const double2* ysrc = yin[cur] + (wc * NCW) * 32;
            const double* csrc = scoef + (wc * NCW) * 8 + g;  // B-fragment lane holds column 8 ct + g

            // flattened (kt, j) loop with PF-deep A prefetch
            auto a_ptr = [&](int kt, int j, int m) {  // pass 0 is the static operator when present
                const double2* base = (has_static && j == 0) ? stat : ops + (size_t)(j - has_static) * npad2;
                return base + ((size_t)rtl[m] * KT + kt) * 32 + lane;
            };
            double2 abuf[PF][MR];
            int pk = 0, pj = 0;  // (kt, j) of the next fragment to prefetch
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                if (pk < KT) {
#pragma unroll
                    for (int m = 0; m < MR; ++m) abuf[u][m] = __ldg(a_ptr(pk, pj, m));
                    if (++pj == J) { pj = 0; ++pk; }
                }
            }
            int kt = 0, j = 0;
            double2 b[NCW];
            for (int it0 = 0; it0 < total; it0 += PF) {
#pragma unroll
                for (int u = 0; u < PF; ++u) {
                    if (it0 + u < total) {
                        double2 a[MR];
#pragma unroll
                        for (int m = 0; m < MR; ++m) a[m] = abuf[u][m];
                        if (pk < KT) {
#pragma unroll
                            for (int m = 0; m < MR; ++m) abuf[u][m] = __ldg(a_ptr(pk, pj, m));
                            if (++pj == J) { pj = 0; ++pk; }
                        }
                        if (j == 0) {
                            const int sw = lane ^ ((kt & 1) << 2);
#pragma unroll
                            for (int c = 0; c < NCW; ++c) b[c] = ysrc[(kt * NCT + c) * 32 + sw];
                        }
                        const int sig = has_static ? j - 1 : j;  // -1 -> static operator, coefficient 1
                        if (sig < 0) {
                            mma_block<MR, NCW>(acc, a, b);
                        } else {
                            double2 bs[NCW];
#pragma unroll
                            for (int c = 0; c < NCW; ++c) {
                                const double s = csrc[sig * ncols + c * 8];
                                bs[c] = make_double2(b[c].x * s, b[c].y * s);
                            }
                            mma_block<MR, NCW>(acc, a, bs);
                        }
                        if (++j == J) { j = 0; ++kt; }
                    }
                }
            }

            // ---- epilogue ----
            // post-phase conj(p(t_stage)) on k; pre-phase p(t_next) on the next stage input
            double2 ph_next[MR];
            {
                const int next_entry = stage == 3 ? 2 * step + 2 : (stage == 0 ? 2 * step + 1 : (stage == 1 ? 2 * step + 1 : 2 * step + 2));
                const double tn = framed ? times[next_entry] : 0.0;
#pragma unroll
                for (int m = 0; m < MR; ++m)
                    ph_next[m] = (framed && next_entry != entry) ? frame_phase(mu_row[m], tn) : ph[m];
            }
            double2* ydst = yin[cur ^ 1];
            // k-sum weights 1,2,2,1; next-input step h/2, h/2, h; final update (1/6) h * ksum
            const bool last = (stage == 3);
            const double keep = stage == 0 ? 0.0 : 1.0;
            const double wk = (stage == 1 || stage == 2) ? 2.0 : 1.0;
            const double astep = stage < 2 ? h2 : (stage == 2 ? h : h6);
#pragma unroll
            for (int m = 0; m < MR; ++m) {
#pragma unroll
                for (int c = 0; c < NCW; ++c) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double2 k = cmul_conj_a(ph[m], make_double2(acc.cr[m][c][i], acc.ci[m][c][i]));
                        const int slab = ((m * NCW + c) * 2 + i) * nthr + tid;
                        const double2 yv = yst[slab];
                        kr[m][c][i] = keep * kr[m][c][i] + wk * k.x;
                        ki[m][c][i] = keep * ki[m][c][i] + wk * k.y;
                        const double v_r = last ? kr[m][c][i] : k.x, v_i = last ? ki[m][c][i] : k.y;
                        const double2 nxt = make_double2(yv.x + astep * v_r, yv.y + astep * v_i);
                        if (last) yst[slab] = nxt;
                        if (mvalid[m]) ydst[yin_pos(NCT, rt[m], g, wc * NCW + c, 2 * q + i)] = cmul(ph_next[m], nxt);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < MR; ++m) ph[m] = ph_next[m];
            acc.zero();
            cur ^= 1;
            __syncthreads();  // all warps done reading scoef / yin[old cur] before they are rewritten
        }
    }

#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = col0 + 8 * (wc * NCW + c) + 2 * q + i;
                if (mvalid[m] && row < n && col < B) y[(size_t)row * ldy + col] = yst[((m * NCW + c) * 2 + i) * nthr + tid];
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
};

constexpr int kMaxFusedNpad = 256;
constexpr size_t kSmemLimit = 227 * 1024;

int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            sms = v;
        else
            sms = 148;  // B200
    }
    return sms;
}

bool pick_config(int n, int B, int K_sweep /*0 for shared*/, Config& cfg) {
    const int SMS = sm_count();
    const int npad = round_up8(n);
    if (npad > kMaxFusedNpad || n < 1) return false;
    Geometry geo;
    geo.n = n;
    geo.npad = npad;
    geo.KT = npad / 4;
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
    const int NCWmax = MR <= 2 ? 4 : 2;
    // The busiest SM runs ceil(ctas / #SMs) CTAs of NCW column tiles each: minimise that product,
    // ties go to the wider tile (fewer A-fragment reloads per column).
    long best_cost = -1;
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
        const long cost = (long)((ctas + SMS - 1) / SMS) * NCW;
        if (!found || cost < best_cost) {
            found = true;
            best_cost = cost;
            cfg.geo = g2;
            cfg.MR = MR;
            cfg.NCW = NCW;
            cfg.threads = threads;
            cfg.smem = smem;
            cfg.grid = ctas;
        }
    }
    return found;
}

template <int MR, int NCW>
int launch_shared_t(const Config& cfg, int B, int S, const double2* gen, double h, double2* y, int ldy, cudaStream_t st) {
    QDB_CUDA(cudaFuncSetAttribute(rk4_shared_kernel<MR, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    rk4_shared_kernel<MR, NCW><<<cfg.grid, cfg.threads, cfg.smem, st>>>(cfg.geo, B, S, gen, h, y, ldy);
    QDB_LAUNCH_CHECK("rk4_shared_kernel");
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

#define QDB_DISPATCH(MRv, NCWv, CALL)                \
    if (cfg.MR == MRv && cfg.NCW == NCWv) return CALL

}  // namespace

bool rk4_fused_supported(int n) { return n >= 1 && round_up8(n) <= kMaxFusedNpad; }

int launch_rk4_fused_shared(int n, int B, int S, const double2* gen_table, double h, double2* y, int ldy, cudaStream_t st) {
    Config cfg;
    if (!pick_config(n, B, 0, cfg)) {
        set_error("rk4 fused: unsupported shape n=%d B=%d", n, B);
        return QDB_E_UNSUPPORTED;
    }
#define ARGS cfg, B, S, gen_table, h, y, ldy, st
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

int launch_rk4_fused_sweep(int n, int K, int B, int S, const double2* stat_packed, const double2* ops_packed,
                           const double* coeff, int ldc,
                           const double* mu, const double* times_dev, double h, double2* y, int ldy, cudaStream_t st) {
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
