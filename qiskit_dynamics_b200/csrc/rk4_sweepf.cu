// Fused RK4 in sweep mode (per-column signal values) with the generator FORMED per column on the tensor pipe
// (SURVEY.md 8(a) rows a1 + a2 + a3 + a7 per column; the reference's order of operations: G_b = G_d + sum_j c_jb G_j,
// then G_b y_b -- models/operator_collections.py:101-134).
//
// Why.  rk4_sweep_kernel computes sum_j G_j (c_jb y_b): K+1 complex operator passes, 4 (K+1) fp64 FMAs per generator
// element and column.  The fp64 tensor pipe and the fp64 FMA pipe of sm_100 are ONE pipe (profiles/r01_l_fp64_probe2.jsonl:
// any mix of DMMA and DFMA issues 16 FMA per clock and sub-partition), so what counts is the FMA total -- and forming
// G_b first needs only 2 K + 4: the coefficients are real, so a generator element costs 2 K FMAs, and its use 4.
//   * formation = a DMMA whose k dimension is the OPERATOR index: A = 8 generator elements (rows 8 rt + g of column c)
//     x 4 operators, B = 4 operators x 8 columns of the batch (the signal values of the stage, held in registers),
//     C = the static operator broadcast over the columns, D = Re (or Im) of G_b[8 rt + g][c] for 8 columns.
//     K <= 4: one DMMA per part, K <= 8: two, ... K <= 16: four (template parameter KS).
//   * D arrives in exactly the accumulator layout (lane 4 g + q: row g, columns 2 q, 2 q + 1), so the product with
//     y_b[c] is four DFMAs per column on registers; the state row c is a 32 B broadcast load from shared memory.
// Per generator element and column: 2 Kpad + 4 FMAs (20 at K = 8) against 36 -- the algorithmic count of SURVEY 8(d).
//   * K <= 2 (KS = -K): a DMMA would waste three quarters (half) of its k dimension, so the formation runs on DFMAs as
//     well -- 2 K + 4 FMAs against the 4 (K + 1) of the operator-pass kernels.  Operators are then stored like the static
//     operator (8 entries per row tile and matrix column), the signal values per lane are those of its two columns.
//
// A CTA owns 8 NCT whole columns for the entire launch (as in rk4_fused.cu).  Operators stream L2 -> registers in
// fragment order (layout below, built per call by pack_sweepf_kernel) through a register ring of 2 matrix columns --
// 4 for small warp tiles, which have too few independent DMMA -> DFMA chains per column -- each slot refilled as soon
// as its column is consumed.  The stage vector lives in shared memory as [row][column] (row stride 8 NCT + 1 complex
// numbers: the epilogue's two-row stores fall into disjoint banks), single buffered between two barriers per stage; y
// and the RK4 k-sum sit in thread-private shared-memory slabs.  Frame phases: on the stage-vector rows when written,
// on the result rows when read (as rk4_sweep_kernel).  When the batch cannot fill the chip, a second warp group
// (KSPLIT) takes half of the matrix columns and hands its partial sums over through shared memory.
//
// Operator layout ("formed-sweep"): opsf[((rt * C2 + c) * KS + ks) * 32 + 4 g + q] = ops[4 ks + q][8 rt + g][c],
// (K <= 2: opsf[((rt * C2 + c) * K + j) * 8 + g] = ops[j][8 rt + g][c]);
// statf[((rt * C2 + c) * 8 + g) * 2 + {0, 1}] = (re, re), (im, im) of stat[8 rt + g][c] -- stored duplicated because the
// DMMA C operand is a register PAIR (both columns of the lane get the same static element): loaded this way it is
// usable as it arrives, instead of four register moves in front of every formation DMMA (ncu, first version: 80 moves per
// two matrix columns, on the issue path of the DMMAs).  Zero outside the matrix; C2 = n rounded up to the matrix columns
// in flight (times the warp groups that share them).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "qdb_common.cuh"
#include "rk4_device.cuh"

namespace qdb {

namespace {

struct FGeo {
    int n, npad, C2, RT;
    int WR, WC, NCT;
    int WK;  // warp groups that share the matrix columns (1, or 2 when the batch cannot fill the chip)
};

// D = A B + C with C kept intact (the static operator pair is reused for every column tile)
__device__ __forceinline__ void dmma_from(double& d0, double& d1, double a, double b, double2 c) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d0), "=d"(d1)
        : "d"(a), "d"(b), "d"(c.x), "d"(c.y));
}

__global__ void pack_sweepf_kernel(int n, int K, int KS, int C2, int RT, const double2* __restrict__ ops_packed,
                                   const double2* __restrict__ stat_packed, double2* __restrict__ opsf,
                                   double2* __restrict__ statf) {
    const int kpad = round_up16(n);
    const size_t per_op = (size_t)round_up8(n) * kpad;
    const size_t nops = KS > 0 ? (size_t)RT * C2 * KS * 32 : (size_t)RT * C2 * K * 8;
    const size_t nstat = (size_t)RT * C2 * 8;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nops && KS < 0) {  // FMA formation: [rt][c][j][g]
        const int gg = (int)(e & 7);
        size_t t = e >> 3;
        const int j = (int)(t % K);
        t /= K;
        const int c = (int)(t % C2), rt = (int)(t / C2);
        const int r = 8 * rt + gg;
        double2 v = make_double2(0.0, 0.0);
        if (r < n && c < n) v = ops_packed[(size_t)j * per_op + packed_index(kpad, r, c)];
        opsf[e] = v;
    } else if (e < nops) {
        const int lane = (int)(e & 31);
        size_t t = e >> 5;
        const int ks = (int)(t % KS);
        t /= KS;
        const int c = (int)(t % C2), rt = (int)(t / C2);
        const int j = 4 * ks + (lane & 3), r = 8 * rt + (lane >> 2);
        double2 v = make_double2(0.0, 0.0);
        if (j < K && r < n && c < n) v = ops_packed[(size_t)j * per_op + packed_index(kpad, r, c)];
        opsf[e] = v;
    } else if (e < nops + nstat && statf != nullptr) {
        const size_t s = e - nops;
        const int gg = (int)(s & 7);
        const size_t t = s >> 3;
        const int c = (int)(t % C2), rt = (int)(t / C2);
        const int r = 8 * rt + gg;
        double2 v = make_double2(0.0, 0.0);
        if (stat_packed != nullptr && r < n && c < n) v = stat_packed[packed_index(kpad, r, c)];
        statf[2 * s] = make_double2(v.x, v.x);
        statf[2 * s + 1] = make_double2(v.y, v.y);
    }
}

// columns of the generator in flight per warp: small warp tiles (one or two DMMA tiles) have too few independent
// DMMA -> DFMA chains per matrix column to cover their latency with the one or two warps a sub-partition holds
template <int MR, int NCW>
struct ColumnsInFlight {
    static constexpr int value = (MR * NCW <= 2) ? 4 : 2;  // cfg2 (one tile per warp): 15.8 / 10.4 / 11.6 us per step with 2 / 4 / 8
};

template <int MR, int NCW, int KS, bool KSPLIT, int NTHR>
__global__ void __launch_bounds__(NTHR, 1)
rk4_sweepf_kernel(FGeo geo, int K, int B, int S, const double2* __restrict__ statf /*or null*/,
                  const double2* __restrict__ opsf, const double* __restrict__ coeff /*[2S+1][K][ldc]*/, int ldc,
                  const double* __restrict__ mu, const double* __restrict__ times /*[2S+1]*/, double h,
                  double2* __restrict__ y, int ldy) {
    extern __shared__ __align__(16) double2 sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    // warp = wr + WR (wc + WC wk): group wk works on its share of the matrix columns for the same rows and batch columns;
    // group 0 owns the state (slabs, epilogue), the others hand their partial sums over through shared memory
    const int wgrp = geo.WR * geo.WC;
    const int wk = KSPLIT ? warp / wgrp : 0, w0 = warp - wk * wgrp;
    const int wr = w0 % geo.WR, wc = w0 / geo.WR;
    const int n = geo.n, C2 = geo.C2, NCT = geo.NCT;
    const int nthr_all = blockDim.x;
    const int nthr = 32 * wgrp;          // threads of one group: the stride of the thread-private slabs
    const int tid0 = tid - wk * nthr;    // index within the group
    const bool owner = KSPLIT ? (wk == 0) : true;
    const int ncols = 8 * NCT, LD = ncols + 1;
    double2* ys = sm;                                   // [npad][LD] stage vector, pre-phased
    double2* yslab = sm + (size_t)geo.npad * LD;        // [MR*NCW*2][nthr]
    double2* kslab = yslab + (size_t)MR * NCW * 2 * nthr;
    double2* red = kslab + (size_t)MR * NCW * 2 * nthr;  // [MR*NCW*2][nthr], only when WK == 2
    const int col0 = blockIdx.x * ncols;
    const int lc0 = 8 * wc * NCW;  // first local column of this warp
    const bool framed = (mu != nullptr);
    const bool has_stat = (statf != nullptr);

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
    double2 ph[MR];
    {
        const double t0 = framed ? times[0] : 0.0;
#pragma unroll
        for (int m = 0; m < MR; ++m) ph[m] = framed ? frame_phase(mu_row[m], t0) : make_double2(1.0, 0.0);
    }

    for (int i = tid; i < geo.npad * LD; i += nthr_all) ys[i] = make_double2(0.0, 0.0);
    __syncthreads();
    if (owner) {
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int lc = lc0 + 8 * c + 2 * q + i;
                double2 v = make_double2(0.0, 0.0);
                if (mvalid[m] && row < n && col0 + lc < B) v = y[(size_t)row * ldy + col0 + lc];
                yslab[((m * NCW + c) * 2 + i) * nthr + tid0] = v;
                if (mvalid[m]) ys[row * LD + lc] = cmul(ph[m], v);
            }
    }

    double ar[MR][NCW][2], ai[MR][NCW][2];  // G_b y accumulators (C-fragment layout)
#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c) ar[m][c][0] = ar[m][c][1] = ai[m][c][0] = ai[m][c][1] = 0.0;

    // operator stream: per matrix column c and row tile, KS fragments of 32 complex numbers (K <= 2: K x 8 entries)
    // + 8 static entries
    constexpr bool FMAF = (KS < 0);       // formation on the FMA pipe
    constexpr int NF = FMAF ? -KS : KS;   // fragments (DMMA k-steps), or operators, per matrix column and row tile
    constexpr int FSTRIDE = FMAF ? 8 : 32;
    using CoefT = typename std::conditional<FMAF, double2, double>::type;
    const double2* pa[MR];
    const double2* ps[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) {
        pa[m] = opsf + (size_t)rtl[m] * C2 * NF * FSTRIDE + (FMAF ? g : lane);
        ps[m] = has_stat ? statf + ((size_t)rtl[m] * C2 * 8 + g) * 2 : nullptr;
    }
    constexpr int CU = ColumnsInFlight<MR, NCW>::value;
    double2 ring[CU][MR][NF], rs[CU][MR][2];
    auto fetch = [&](int c, double2 (&dst)[MR][NF], double2 (&dsts)[MR][2]) {
#pragma unroll
        for (int m = 0; m < MR; ++m) {
#pragma unroll
            for (int ks = 0; ks < NF; ++ks) dst[m][ks] = __ldg(pa[m] + ((size_t)c * NF + ks) * FSTRIDE);
            dsts[m][0] = has_stat ? __ldg(ps[m] + (size_t)c * 16) : make_double2(0.0, 0.0);
            dsts[m][1] = has_stat ? __ldg(ps[m] + (size_t)c * 16 + 1) : make_double2(0.0, 0.0);
        }
    };
    const int cspan = KSPLIT ? C2 / 2 : C2, cb = wk * cspan, ce = cb + cspan;  // this group's matrix columns
#pragma unroll
    for (int i = 0; i < (CU == 2 ? 1 : CU); ++i) fetch(cb + i, ring[i], rs[i]);

    // signal values of a stage in DMMA B-fragment order: lane (g, q) holds c[4 ks + q][column 8 ct + g]
    // (FMA formation: c[j][columns 8 ct + 2 q, + 1], the two columns the lane accumulates)
    auto load_coef = [&](int entry, CoefT (&cf)[NCW][NF]) {
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int ks = 0; ks < NF; ++ks) {
                if constexpr (FMAF) {
                    const int col = col0 + lc0 + 8 * c + 2 * q;
                    const double* src = coeff + ((size_t)entry * K + ks) * ldc + col;
                    cf[c][ks] = make_double2(col < B ? src[0] : 0.0, col + 1 < B ? src[1] : 0.0);
                } else {
                    const int j = 4 * ks + q, col = col0 + lc0 + 8 * c + g;
                    cf[c][ks] = (j < K && col < B) ? coeff[((size_t)entry * K + j) * ldc + col] : 0.0;
                }
            }
    };
    CoefT cf[NCW][NF], cfn[NCW][NF];
    load_coef(0, cf);
    __syncthreads();

    // one matrix column: form G_b[rows of this warp][c] for the warp's columns, multiply with y_b[c]
    auto column = [&](int c, const double2 (&A)[MR][NF], const double2 (&As)[MR][2]) {
        double2 yv[NCW][2];
#pragma unroll
        for (int ct = 0; ct < NCW; ++ct)
#pragma unroll
            for (int i = 0; i < 2; ++i) yv[ct][i] = ys[c * LD + lc0 + 8 * ct + 2 * q + i];
        double gr[2][NCW][2], gi[2][NCW][2];
        auto form = [&](int m, int slot) {
            if constexpr (FMAF) {
#pragma unroll
                for (int ct = 0; ct < NCW; ++ct) {
                    gr[slot][ct][0] = fma(cf[ct][0].x, A[m][0].x, As[m][0].x);
                    gr[slot][ct][1] = fma(cf[ct][0].y, A[m][0].x, As[m][0].x);
                    gi[slot][ct][0] = fma(cf[ct][0].x, A[m][0].y, As[m][1].x);
                    gi[slot][ct][1] = fma(cf[ct][0].y, A[m][0].y, As[m][1].x);
                }
#pragma unroll
                for (int j = 1; j < NF; ++j)
#pragma unroll
                    for (int ct = 0; ct < NCW; ++ct) {
                        gr[slot][ct][0] = fma(cf[ct][j].x, A[m][j].x, gr[slot][ct][0]);
                        gr[slot][ct][1] = fma(cf[ct][j].y, A[m][j].x, gr[slot][ct][1]);
                        gi[slot][ct][0] = fma(cf[ct][j].x, A[m][j].y, gi[slot][ct][0]);
                        gi[slot][ct][1] = fma(cf[ct][j].y, A[m][j].y, gi[slot][ct][1]);
                    }
            } else {
#pragma unroll
                for (int ct = 0; ct < NCW; ++ct) {
                    dmma_from(gr[slot][ct][0], gr[slot][ct][1], A[m][0].x, cf[ct][0], As[m][0]);
                    dmma_from(gi[slot][ct][0], gi[slot][ct][1], A[m][0].y, cf[ct][0], As[m][1]);
                }
#pragma unroll
                for (int ks = 1; ks < NF; ++ks)
#pragma unroll
                    for (int ct = 0; ct < NCW; ++ct) {
                        dmma(gr[slot][ct][0], gr[slot][ct][1], A[m][ks].x, cf[ct][ks]);
                        dmma(gi[slot][ct][0], gi[slot][ct][1], A[m][ks].y, cf[ct][ks]);
                    }
            }
        };
        form(0, 0);
#pragma unroll
        for (int m = 0; m < MR; ++m) {
            if (m + 1 < MR) form(m + 1, (m + 1) & 1);  // next row tile's DMMAs are queued before this tile's DFMAs wait
            const int s = m & 1;
            // two passes over the tile's elements so that no DFMA follows the one it depends on
#pragma unroll
            for (int ct = 0; ct < NCW; ++ct)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    ar[m][ct][i] = fma(gr[s][ct][i], yv[ct][i].x, ar[m][ct][i]);
                    ai[m][ct][i] = fma(gr[s][ct][i], yv[ct][i].y, ai[m][ct][i]);
                }
#pragma unroll
            for (int ct = 0; ct < NCW; ++ct)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    ar[m][ct][i] = fma(-gi[s][ct][i], yv[ct][i].y, ar[m][ct][i]);
                    ai[m][ct][i] = fma(gi[s][ct][i], yv[ct][i].x, ai[m][ct][i]);
                }
        }
    };

    const int total_stages = 4 * S;
#pragma unroll 1
    for (int sidx = 0; sidx < total_stages; ++sidx) {
        const int step = sidx >> 2, stage = sidx & 3;
        const int entry = 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
        const int nstage = (stage + 1) & 3, nstep = step + (stage == 3 ? 1 : 0);
        const int next_entry = (sidx + 1 < total_stages) ? 2 * nstep + (nstage == 0 ? 0 : (nstage == 3 ? 2 : 1)) : entry;
        load_coef(next_entry, cfn);  // lands during the column loop
        // frame phases of the next stage time, computed in front of the epilogue (in-order warps cannot run them "under" the
        // column loop, and large tiles have no registers to carry them across it)
        constexpr bool kEarlyPhases = false;  // tried for small tiles: the sincos in front of the loop costs more than it hides (cfg2 10.4 -> 11.6 us/step)
        double2 ph_next[MR];
        auto next_phases = [&]() {
            const double tn = framed ? times[next_entry] : 0.0;
#pragma unroll
            for (int m = 0; m < MR; ++m) ph_next[m] = (framed && next_entry != entry) ? frame_phase(mu_row[m], tn) : ph[m];
        };
        if constexpr (kEarlyPhases) next_phases();

        if constexpr (CU == 2) {
#pragma unroll 1
            for (int c = cb; c < ce; c += 2) {
                fetch(c + 1, ring[1], rs[1]);
                column(c, ring[0], rs[0]);
                fetch(c + 2 < ce ? c + 2 : cb, ring[0], rs[0]);  // wraps: the operators are time independent
                column(c + 1, ring[1], rs[1]);
            }
        } else {
#pragma unroll 1
            for (int c = cb; c < ce; c += CU) {
#pragma unroll
                for (int i = 0; i < CU; ++i) {
                    column(c + i, ring[i], rs[i]);
                    const int nxt = c + CU + i;  // refill the slot CU - 1 columns ahead (wraps around)
                    fetch(nxt < ce ? nxt : nxt - cspan, ring[i], rs[i]);
                }
            }
        }

        // ---- partial sums of the other warp group (matrix columns split over two groups) ----
        if constexpr (KSPLIT) {
            if (!owner) {
#pragma unroll
                for (int m = 0; m < MR; ++m)
#pragma unroll
                    for (int ct = 0; ct < NCW; ++ct)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            red[((m * NCW + ct) * 2 + i) * nthr + tid0] = make_double2(ar[m][ct][i], ai[m][ct][i]);
                            ar[m][ct][i] = 0.0;
                            ai[m][ct][i] = 0.0;
                        }
            }
            __syncthreads();
            if (owner) {
#pragma unroll
                for (int m = 0; m < MR; ++m)
#pragma unroll
                    for (int ct = 0; ct < NCW; ++ct)
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const double2 p = red[((m * NCW + ct) * 2 + i) * nthr + tid0];
                            ar[m][ct][i] += p.x;
                            ai[m][ct][i] += p.y;
                        }
            }
        }

        // ---- epilogue: post-phase conj(p(t_stage)) on k, RK4 combine, pre-phase p(t_next) on the next stage input ----
        if constexpr (!kEarlyPhases) next_phases();
        const StageCoef sc(stage, h);
        if (owner) {
#pragma unroll
            for (int m = 0; m < MR; ++m)
#pragma unroll
                for (int ct = 0; ct < NCW; ++ct)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int slot = ((m * NCW + ct) * 2 + i) * nthr + tid0;
                        const double2 k = cmul_conj_a(ph[m], make_double2(ar[m][ct][i], ai[m][ct][i]));
                        double2 ks = stage == 0 ? make_double2(0.0, 0.0) : kslab[slot];
                        ks.x = sc.keep * ks.x + sc.wk * k.x;
                        ks.y = sc.keep * ks.y + sc.wk * k.y;
                        const double v_r = sc.last ? ks.x : k.x, v_i = sc.last ? ks.y : k.y;
                        const double2 yv = yslab[slot];
                        const double2 nxt = make_double2(yv.x + sc.astep * v_r, yv.y + sc.astep * v_i);
                        if (sc.last) yslab[slot] = nxt; else kslab[slot] = ks;
                        const double2 pn = cmul(ph_next[m], nxt);
                        ar[m][ct][i] = pn.x;  // parked in the accumulator registers until every warp has left the column loop
                        ai[m][ct][i] = pn.y;
                    }
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < MR; ++m)
#pragma unroll
            for (int ct = 0; ct < NCW; ++ct)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (owner && mvalid[m]) ys[(8 * rt[m] + g) * LD + lc0 + 8 * ct + 2 * q + i] = make_double2(ar[m][ct][i], ai[m][ct][i]);
                    ar[m][ct][i] = 0.0;
                    ai[m][ct][i] = 0.0;
                }
#pragma unroll
        for (int m = 0; m < MR; ++m) ph[m] = ph_next[m];
#pragma unroll
        for (int ct = 0; ct < NCW; ++ct)
#pragma unroll
            for (int ks = 0; ks < NF; ++ks) cf[ct][ks] = cfn[ct][ks];
        __syncthreads();
    }

#pragma unroll
    for (int m = 0; m < MR; ++m)
#pragma unroll
        for (int c = 0; c < NCW; ++c)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * rt[m] + g;
                const int col = col0 + lc0 + 8 * c + 2 * q + i;
                if (owner && mvalid[m] && row < n && col < B) y[(size_t)row * ldy + col] = yslab[((m * NCW + c) * 2 + i) * nthr + tid0];
            }
}

struct FConfig {
    FGeo geo;
    int MR, NCW, KS, threads, grid;
    double wave_cost;  // busiest SM of one wave, in column tiles of the reference tiling (see pick_sweepf)
    size_t smem;
};

constexpr size_t kSmemLimitF = 227 * 1024;

// max_ctas > 0: only tilings whose grid fits that many CTAs (one wave); need_C2 > 0: only tilings that read the operator copy
// packed for that column padding, without the warp-group split (the second launch of a wave-balanced pair, see
// launch_rk4_sweepf)
bool pick_sweepf(int n, int B, int K, FConfig& cfg, int max_ctas = 0, int need_C2 = 0) {
    if (n < 1 || round_up8(n) > 256 || K < 1 || K > 16 || B < 1) return false;
    const int SMS = sm_count();
    FGeo geo;
    geo.n = n;
    geo.npad = round_up8(n);
    geo.C2 = 0;  // set with the tiling: n rounded up to the columns in flight
    geo.RT = geo.npad / 8;
    const int CT = (B + 7) / 8;
    int WR, WC, MR;
    if (geo.RT >= 8) {
        WR = 8;
        WC = 1;
        MR = (geo.RT + 7) / 8;
        const int MR4 = (geo.RT + 3) / 4;
        if (MR4 <= 4 && MR4 * 4 < MR * 8) {  // fewer idle row slots with 4 row warps x 2 column warps
            WR = 4;
            WC = 2;
            MR = MR4;
        }
    } else {
        WR = 1;
        while (WR < geo.RT) WR *= 2;
        MR = 1;
        WC = 8 / WR;
    }
    if (const char* force = getenv("QDB_FORCE_SWEEPF")) {  // "WR,WC,MR" (profiling only)
        int fwr, fwc, fmr;
        if (sscanf(force, "%d,%d,%d", &fwr, &fwc, &fmr) == 3 && fwr * fmr >= geo.RT && fwr * fwc <= 12 && fmr >= 1 && fmr <= 4) {
            WR = fwr;
            WC = fwc;
            MR = fmr;
        }
    }
    static const int ncw_opts[5][4] = {{0, 0, 0, 0}, {4, 3, 2, 1}, {3, 2, 1, 0}, {2, 1, 0, 0}, {1, 0, 0, 0}};
    int force_ncw = 0;
    if (const char* fn = getenv("QDB_FORCE_SWEEPF_NCW")) force_ncw = atoi(fn);
    bool found = false;
    double best = 0;
    // fewer column warps when the batch is too small to give every SM a CTA; CTAs that end up with at most four warps
    // then split the matrix columns over two warp groups (WK = 2) so that every sub-partition still holds two warps
    const char* nok = getenv("QDB_SWEEPF_NO_KSPLIT");
    const bool ksplit_ok = !(nok && nok[0] == '1');
    const int MR0 = MR;
    auto consider = [&](int wr, int wc, int mr, int NCW) {
        if (NCW == 0 || (force_ncw > 0 && NCW != force_ncw)) return;
        const int threads0 = 32 * wr * wc;
        if (threads0 > 256 && (K <= 2 || mr * NCW > 4 || threads0 > 384)) return;  // 12-warp CTAs: 168 registers per thread
        const int NCT = NCW * wc;
        const int ctas = (CT + NCT - 1) / NCT;
        if (max_ctas > 0 && ctas > max_ctas) return;
        const int cu = (mr * NCW <= 2) ? 4 : 2;  // ColumnsInFlight<MR, NCW>
        const int WK = (ksplit_ok && need_C2 == 0 && mr * NCW <= 4 && threads0 <= 128 && ctas <= SMS && n >= 2 * cu) ? 2 : 1;
        const int C2 = (n + cu * WK - 1) / (cu * WK) * (cu * WK);
        if (need_C2 > 0 && C2 != need_C2) return;
        const int threads = threads0 * WK;
        const size_t smem = ((size_t)geo.npad * (8 * NCT + 1) + (size_t)(WK == 2 ? 3 : 2) * mr * NCW * 2 * threads0) * sizeof(double2);
        if (smem > kSmemLimitF) return;
        // busiest SM of a wave: fixed per-stage part + pipe time of the column tiles of its busiest sub-partition, in units
        // of the rule-based tiling's row tiles (so that those tilings keep the cost they were tuned with).  A sub-partition
        // with two or three warps keeps the fp64 pipe full; a lone warp (128-thread CTAs) reaches about 0.65 of it but leaves
        // twice as many CTAs for a batch that cannot fill the chip.  Measured at n = 81, K = 8: 12 units 208 us, 9 units 152 us.
        const double wq = threads0 <= 256 ? threads0 / 128.0 : (double)((threads0 / 32 + 3) / 4);  // warps of the busiest sub-partition
        const double wave = 0.5 + (double)NCW * mr * wq / MR0 / (threads >= 256 ? 1.0 : 0.65);
        const double cost = (double)((ctas + SMS - 1) / SMS) * wave;
        if (!found || cost < best - 1e-9) {
            found = true;
            best = cost;
            cfg.geo = geo;
            cfg.geo.WR = wr;
            cfg.geo.WC = wc;
            cfg.geo.NCT = NCT;
            cfg.geo.WK = WK;
            cfg.geo.C2 = C2;
            cfg.MR = mr;
            cfg.NCW = NCW;
            cfg.KS = K <= 2 ? -K : (K + 3) / 4;  // K <= 2: formation on the FMA pipe
            cfg.threads = threads;
            cfg.grid = ctas;
            cfg.smem = smem;
            cfg.wave_cost = wave;
        }
    };
    for (int wc = WC; wc >= 1; wc /= 2)
        for (int o = 0; o < 4; ++o) consider(WR, wc, MR, ncw_opts[MR][o]);
    // one row tile per warp on 9 .. 12 row warps (three warps per sub-partition): tilings with 3 octets per CTA, which the
    // 8-warp family cannot offer at these sizes -- the second launch of a wave-balanced pair usually needs exactly that
    if (geo.RT > 8 && geo.RT <= 12 && !getenv("QDB_FORCE_SWEEPF"))
        for (int o = 0; o < 4; ++o) consider(12, 1, 1, ncw_opts[1][o]);
    return found;
}

template <int MR, int NCW, int KS, bool KSPLIT>
int launch_sweepf_k(const FConfig& cfg, int K, int B, int S, const double2* statf, const double2* opsf, const double* coeff, int ldc,
                    const double* mu, const double* times, double h, double2* y, int ldy, cudaStream_t st) {
    if constexpr (!KSPLIT && KS > 0 && MR * NCW <= 4) {
        if (cfg.threads > 256) {  // 12-warp CTAs (three warps per sub-partition)
            QDB_CUDA(cudaFuncSetAttribute(rk4_sweepf_kernel<MR, NCW, KS, false, 384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
            rk4_sweepf_kernel<MR, NCW, KS, false, 384><<<cfg.grid, cfg.threads, cfg.smem, st>>>(cfg.geo, K, B, S, statf, opsf, coeff, ldc, mu, times, h, y, ldy);
            QDB_LAUNCH_CHECK("rk4_sweepf_kernel");
            return QDB_OK;
        }
    }
    QDB_CUDA(cudaFuncSetAttribute(rk4_sweepf_kernel<MR, NCW, KS, KSPLIT, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    rk4_sweepf_kernel<MR, NCW, KS, KSPLIT, 256><<<cfg.grid, cfg.threads, cfg.smem, st>>>(cfg.geo, K, B, S, statf, opsf, coeff, ldc, mu, times, h, y, ldy);
    QDB_LAUNCH_CHECK("rk4_sweepf_kernel");
    return QDB_OK;
}

template <int MR, int NCW, int KS>
int launch_sweepf_t(const FConfig& cfg, int K, int B, int S, const double2* statf, const double2* opsf, const double* coeff, int ldc,
                    const double* mu, const double* times, double h, double2* y, int ldy, cudaStream_t st) {
    // the matrix-column split exists for CTAs of at most four warps per group: tiles of up to MR x NCW = 4 x 1
    if constexpr (MR * NCW <= 4) {
        if (cfg.geo.WK == 2) return launch_sweepf_k<MR, NCW, KS, true>(cfg, K, B, S, statf, opsf, coeff, ldc, mu, times, h, y, ldy, st);
    }
    return launch_sweepf_k<MR, NCW, KS, false>(cfg, K, B, S, statf, opsf, coeff, ldc, mu, times, h, y, ldy, st);
}

}  // namespace

// 0 = automatic, 1 = legacy kernels only (QDB_SWEEP_KERNEL=legacy), 2 = formed-generator kernel wherever it exists
// (QDB_SWEEP_KERNEL=formed: also for K = 1, 2, where the automatic choice keeps the operator-pass kernels)
static int sweep_kernel_mode() {
    const char* e = getenv("QDB_SWEEP_KERNEL");
    if (e && strcmp(e, "legacy") == 0) return 1;
    if (e && strcmp(e, "formed") == 0) return 2;
    return 0;
}

bool rk4_sweepf_supported(int n, int K) { return n >= 1 && round_up8(n) <= 256 && K >= 1 && K <= 16; }

bool rk4_sweepf_selected(int n, int K, int B, bool small_kernel_available) {
    const int mode = sweep_kernel_mode();
    if (mode == 1 || !rk4_sweepf_supported(n, K)) return false;
    if (mode == 2) return true;
    (void)small_kernel_available;
    // 3 <= K <= 16: the formed-generator kernel is 1.02-4.9x faster at every size tried, shared-memory-resident small
    // systems included (profiles/r01_t_sweep_formed_vs_legacy_small.jsonl, ..._K.jsonl).
    if (K >= 3) return true;
    // K <= 2 (formation on the FMA pipe, 2 K + 4 against 4 (K + 1) FMAs): 1.07-1.85x faster once the batch is more than
    // one wave of column octets, except K = 1 at large n (n=128: 0.90x); below one wave the operator-pass kernels win
    // (n=32, B=1024: 0.72x) -- profiles/r01_z_sweep_formed_vs_legacy_K12.jsonl
    const int octets = (B + 7) / 8;
    return octets > sm_count() && (K == 2 || n < 100);
}

size_t rk4_sweepf_workspace_bytes(int n, int K) {
    const size_t RT = round_up8(n) / 8, C2 = (n + 7) & ~7 /* the largest of the paddings (4 columns in flight x 2 warp groups) */, KS = (K + 3) / 4;
    return (RT * C2 * KS * 32 + RT * C2 * 16) * sizeof(double2);
}

// Wave balancing.  Columns are independent and a CTA owns whole columns, so a batch whose CTA count is not a multiple of the
// SM count pays for a whole last wave (cfg5: 1024 column octets in CTAs of 4 = 256 CTAs on 148 SMs, 1.73 waves at the price
// of 2).  The batch is then cut in two: the full waves with the chosen tiling, and the remaining columns with the cheapest
// tiling that fits ONE wave (usually fewer octets per CTA: 148 x 4 + 144 x 3 octets at cfg5, 7 octets per SM instead of 8).
// QDB_SWEEPF_NO_BALANCE=1 switches it off.
static bool plan_sweepf(int n, int K, int B, FConfig& a, FConfig& b, int& colsA) {
    if (!pick_sweepf(n, B, K, a)) return false;
    colsA = B;
    b.grid = 0;
    const int SMS = sm_count();
    const char* nb = getenv("QDB_SWEEPF_NO_BALANCE");
    if ((nb && nb[0] == '1') || a.geo.WK != 1 || a.grid <= SMS || a.grid % SMS == 0) return true;
    const int full = a.grid / SMS * SMS;
    const int cols_full = full * 8 * a.geo.NCT;
    if (cols_full >= B) return true;
    FConfig c;
    if (!pick_sweepf(n, B - cols_full, K, c, SMS, a.geo.C2) || c.KS != a.KS) return true;
    const double whole = (double)((a.grid + SMS - 1) / SMS) * a.wave_cost;
    const double cut = (double)(a.grid / SMS) * a.wave_cost + c.wave_cost;
    if (cut < 0.97 * whole) {
        a.grid = full;
        b = c;
        colsA = cols_full;
    }
    return true;
}

bool rk4_sweepf_tiling(int n, int B, int K, int* out) {
    FConfig cfg, cfg2;
    int colsA = B;
    if (!plan_sweepf(n, K, B, cfg, cfg2, colsA)) return false;
    if (cfg2.grid > 0) cfg.grid += cfg2.grid;  // CTAs of both launches of a wave-balanced pair
    out[0] = cfg.geo.WR;
    out[1] = cfg.geo.WC * cfg.geo.WK;  // column warps x warp groups sharing the matrix columns
    out[2] = cfg.MR;
    out[3] = cfg.NCW;
    out[4] = 0;
    out[5] = cfg.grid;
    out[6] = cfg.threads;
    out[7] = (int)cfg.smem;
    out[8] = 2;  // 2 = formed-generator sweep kernel
    return true;
}

// One launch of the formed-generator kernel on columns [0, B) of y / coeff with tiling cfg (operator copy already packed).
static int launch_sweepf_cfg(const FConfig& cfg, int K, int B, int S, const double2* statf, const double2* opsf, const double* coeff,
                             int ldc, const double* mu, const double* times_dev, double h, double2* y, int ldy, cudaStream_t st) {
#define QDB_F(mr, ncw)                                                                                                     \
    if (cfg.MR == mr && cfg.NCW == ncw) {                                                                                  \
        if (cfg.KS == -1) return launch_sweepf_t<mr, ncw, -1>(cfg, K, B, S, statf, opsf, coeff, ldc, mu, times_dev, h, y, ldy, st); \
        if (cfg.KS == -2) return launch_sweepf_t<mr, ncw, -2>(cfg, K, B, S, statf, opsf, coeff, ldc, mu, times_dev, h, y, ldy, st); \
        if (cfg.KS == 1) return launch_sweepf_t<mr, ncw, 1>(cfg, K, B, S, statf, opsf, coeff, ldc, mu, times_dev, h, y, ldy, st); \
        if (cfg.KS == 2) return launch_sweepf_t<mr, ncw, 2>(cfg, K, B, S, statf, opsf, coeff, ldc, mu, times_dev, h, y, ldy, st); \
        if (cfg.KS == 3) return launch_sweepf_t<mr, ncw, 3>(cfg, K, B, S, statf, opsf, coeff, ldc, mu, times_dev, h, y, ldy, st); \
        return launch_sweepf_t<mr, ncw, 4>(cfg, K, B, S, statf, opsf, coeff, ldc, mu, times_dev, h, y, ldy, st);           \
    }
    QDB_F(1, 1)
    QDB_F(1, 2)
    QDB_F(1, 3)
    QDB_F(1, 4)
    QDB_F(2, 1)
    QDB_F(2, 2)
    QDB_F(2, 3)
    QDB_F(3, 1)
    QDB_F(3, 2)
    QDB_F(4, 1)
#undef QDB_F
    set_error("rk4 formed sweep: no kernel for MR=%d NCW=%d", cfg.MR, cfg.NCW);
    return QDB_E_UNSUPPORTED;
}

int launch_rk4_sweepf(int n, int K, int B, int S, const double2* stat_packed, const double2* ops_packed, const double* coeff,
                      int ldc, const double* mu, const double* times_dev, double h, double2* y, int ldy, void* ws, cudaStream_t st) {
    FConfig cfg, cfg2;
    int colsA = B;
    if (!plan_sweepf(n, K, B, cfg, cfg2, colsA)) {
        set_error("rk4 formed sweep: unsupported shape n=%d B=%d K=%d", n, B, K);
        return QDB_E_UNSUPPORTED;
    }
    const size_t nops = cfg.KS > 0 ? (size_t)cfg.geo.RT * cfg.geo.C2 * cfg.KS * 32 : (size_t)cfg.geo.RT * cfg.geo.C2 * K * 8, nstat = (size_t)cfg.geo.RT * cfg.geo.C2 * 8;  // static entries (2 double2 each)
    double2* opsf = (double2*)ws;
    double2* statf = stat_packed ? opsf + nops : nullptr;
    const size_t total = nops + (stat_packed ? nstat : 0);
    pack_sweepf_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(n, K, cfg.KS, cfg.geo.C2, cfg.geo.RT, ops_packed, stat_packed, opsf, statf);
    QDB_LAUNCH_CHECK("pack_sweepf_kernel");
    int rc = launch_sweepf_cfg(cfg, K, colsA, S, statf, opsf, coeff, ldc, mu, times_dev, h, y, ldy, st);
    if (rc != QDB_OK || cfg2.grid == 0) return rc;
    return launch_sweepf_cfg(cfg2, K, B - colsA, S, statf, opsf, coeff + colsA, ldc, mu, times_dev, h, y + colsA, ldy, st);
}

}  // namespace qdb
