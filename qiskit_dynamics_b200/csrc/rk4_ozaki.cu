// Fused shared-signal RK4 with the fp64 contraction EMULATED on the int8 tensor cores (tcgen05.mma kind::i8, accumulators
// and the generator operand in TMEM) -- an Ozaki-style error-free split.  n = 65..128 (rows and k padded to 128).
//
// Why.  tcgen05 has no fp64 kind; DMMA is the fp64 tensor pipe of sm_100a and rk4_shared3m_kernel already runs it at 85 %.
// The only way past that roof is to leave the fp64 pipe: every real operand is cut into NS = 5 signed BYTES against a
// per-row (generator) / per-column (stage vector) power-of-two scale,
//     x = 2^e (q_1 2^-8 + q_2 2^-16 + ... + q_NS 2^(-8 NS)) + O(2^(e - 8 NS - 1)),     q_p integers in [-128, 127],
// so that the products of slices are EXACT in int32 (K = 128, <= NS pairs per accumulator: |sum| < 2^24) and G y = sum over
// slice pairs.  Pairs of equal weight 2^(-8 (p + q)) share an accumulator ("group" g = p + q); pairs with g > NS + 1 lie
// below the truncation error of the operands and are dropped: 15 pairs x 4 real products (re = Ar Br + Ai (-Bi),
// im = Ar Bi + Ai Br).  Error per RHS evaluation: normwise 2^-40 per operand (measured against the DMMA kernel in
// tests/test_ozaki_gpu.py).  Balanced byte digits cost nothing to extract: with X = rint(x 2^(8 NS - e)) the bytes of
// (X + 0x80808080) ^ 0x80808080 ARE the signed digits, and a plane word (the same slice of four columns) is a 4 x 4 byte
// transpose -- eight PRMT.
//
// Hardware mapping (measured first: profiles/probe/umma_i8_probe.cu -> profiles/r02_m_umma_i8_probe.jsonl).  With both
// operands in shared memory an M128 x N x K32 int8 MMA costs (4096 + 32 N) / 128 cycles -- the operand READ, 41 cycles at
// N = 32 -- so the generator slices live in TMEM (A operand from TMEM: 21 cycles at N = 32, 33 at N = 64 = the peak of
// 8192 MAC/clk/SM), and one N = 64 MMA computes (re | im) of 32 columns at once from the operand planes (B_re | B_im) and
// (-B_im | B_re) (BImage in ozaki_device.cuh).  A CTA owns 32 whole columns for the launch (4096 columns = 128 CTAs; 16 per CTA for small batches):
//   TMEM: columns [0, 192) three int32 accumulator buffers (re | im), so that the five groups of a stage never wait for a
//     drain; [192, 512) the 2 NS generator slice planes (32 columns of packed int8 each), reloaded when the stage time
//     changes (every other stage);
//   shared memory: the stage-vector operand images of the NS slices in the no-swizzle MN-major core-matrix layout the MMA
//     reads ([-im | re | im] per k group: the planes (re | im) and (-im | re) overlap in re; 60 KB), y and the RK4 k-sum as
//     fp64 (128 KB);
//   warps 0-15: epilogue, thread = (row, 8 columns) -- drain a group (tcgen05.ld), combine the groups in int64, one
//     conversion to fp64, RK4 combine, column scales (warp REDUX + one named barrier), re-slice the next stage vector into
//     shared memory; warp 16: issues the 120 MMAs of a stage, LEAST significant group first; warps 17-20: load the next
//     generator entry's planes from L2 (coalesced: the planes are stored chunk-major) and tcgen05.st them into TMEM.  Slice
//     plane p is last read by group p + 1, so with the groups in descending order the planes are released one by one
//     DURING the stage (p_free mbarriers) and the reload hides behind the MMAs and the epilogue instead of following them.
//   Pipelines: full / empty mbarriers per accumulator buffer (MMA <-> epilogue), b_ready (stage vector sliced), a_ready
//     (generator planes in TMEM), p_free[p] (plane p no longer read).
// A stage is serial per column tile (MMA phase ~4.7K cycles, then last drain + combine + re-slice ~5.3K: measured with
// -DQDB_OZ_TIMELINE, profiles/r02_o_ozaki_variants.jsonl); the tensor pipe is busy 27 % of the time (ncu,
// profiles/r02_n_rk4_ozaki_ncu.json).
#include <cstdint>
#include <cstdlib>

#include "ozaki_device.cuh"

namespace qdb {

// -DQDB_OZ_TIMELINE: clock64() stamps of one stage of CTA 0 (profiles/probe/ozaki_probe.py prints them)
__device__ long long g_oz_dbg[64];
__device__ int g_oz_dbg_stage = 8;

namespace {

#ifdef QDB_OZ_TIMELINE
#define OZ_DBG(i) do { if (blockIdx.x == 0 && sidx == g_oz_dbg_stage && (threadIdx.x & 31) == 0) g_oz_dbg[i] = clock64(); } while (0)
#else
#define OZ_DBG(i) do { } while (0)
#endif


// shared-memory carve-up (bytes) of a CTA with SETS column sets of CS columns; the operand image of a slice (BImage<CS>) is
// 3 CS KD bytes
struct Smem {
    int b, y, k, red, ea, bar, tmem, total;
    __host__ __device__ constexpr Smem(int cs, int sets)
        : b(0),                                  // [SETS][NS][3 CS KD] int8: [-im | re | im] of every slice
          y(b + sets * NS * 3 * cs * KD),        // [SETS CS][KD] double2
          k(y + sets * cs * KD * 16),            // [SETS CS][KD] double2
          red(k + sets * cs * KD * 16),          // [SETS][4][CS] unsigned (high words of the column maxima)
          ea(red + sets * 4 * cs * 4),           // [4][KD] int: row exponents of generator entry e in slot e & 3 (the loaders run up to two entries ahead of the epilogue)
          bar(ea + 4 * KD * 4),                  // SETS (2 NACC + 1) + 1 + NS mbarriers
          tmem(bar + 32 * 8),
          total(tmem + 16) {}
};


// ---- generator table (row-major fp64, as generator_kernel writes it) -> int8 slice planes + row exponents ----
// planes[t][part][p][k / 16][row][k % 16] (a loader lane = a row reads 16 B next to its neighbours': coalesced; the eight
// chunks of a row are its TMEM image), expo[t][row]; one block per (row, t)
// the table comes row-major (generator_kernel's QDB_LAYOUT_ROWMAJOR) or in the packed DMMA-fragment layout (QDB_LAYOUT_PACKED)
template <bool PACKED>
__global__ void __launch_bounds__(128) ozaki_gslice_kernel(int n, const double2* __restrict__ gen, int8_t* __restrict__ planes,
                                                           int* __restrict__ expo) {
    const int row = blockIdx.x, t = blockIdx.y, k = threadIdx.x;
    __shared__ unsigned wmax[4];
    double2 v = make_double2(0.0, 0.0);
    if (row < n && k < n) {
        const int kpad = round_up16(n);
        v = PACKED ? gen[(size_t)t * round_up8(n) * kpad + packed_index(kpad, row, k)] : gen[((size_t)t * n + row) * n + k];
    }
    unsigned m = max(abs_hi(v.x), abs_hi(v.y));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((k & 31) == 0) wmax[k >> 5] = m;
    __syncthreads();
    m = max(max(wmax[0], wmax[1]), max(wmax[2], wmax[3]));
    const int e = slice_exponent_hi(m);
    if (k == 0) expo[(size_t)t * KD + row] = e;
    const double scale = pow2(8 * NS - e);
    const long long dr = digits_of(v.x, scale), di = digits_of(v.y, scale);
    int8_t* base = planes + (size_t)t * 2 * NS * KD * KD + (size_t)(k >> 4) * (KD * 16) + row * 16 + (k & 15);
#pragma unroll
    for (int p = 0; p < NS; ++p) {
        base[(size_t)(0 * NS + p) * KD * KD] = (int8_t)(dr >> (8 * (NS - 1 - p)));
        base[(size_t)(1 * NS + p) * KD * KD] = (int8_t)(di >> (8 * (NS - 1 - p)));
    }
}


__device__ __forceinline__ int stage_entry(int sidx) {
    const int step = sidx >> 2, stage = sidx & 3;
    return 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
}

// A CTA owns SETS independent column sets of CS columns (CS / 2 epilogue warps each) that share the generator in TMEM and the
// MMA warp.  Instantiated: <32, 1> (one wave of 32-column CTAs: batches above 16 columns per SM) and <16, 1> (smaller batches:
// twice the CTAs).  Two 16-column sets per CTA whose stages interleave on the MMA warp -- eight epilogue warps per set, or all
// sixteen alternating with four columns per thread -- were built and measured SLOWER than <32, 1> (27.0 / 30.2 against 23.8 us
// per step, profiles/r02_o_ozaki_variants.jsonl): the sets' epilogues collide and the N = 32 MMA costs 27.6 cycles.
// NKS: k chunks of 32 that hold data (4; 3 for n <= 96 -- a compile-time count: a run-time test per MMA costs 14 % of the step)
template <int CS, int SETS, int NKS>
__global__ void __launch_bounds__((SETS * (CS / 2) + 1 + LOADERS) * 32, 1)
rk4_ozaki_kernel(int n, int B, int S, const int8_t* __restrict__ planes, const int* __restrict__ expo, double h, double2* __restrict__ y,
                 int ldy) {
    constexpr Smem L(CS, SETS);
    constexpr int SET_WARPS = CS / 2, EPI_WARPS = SETS * SET_WARPS, MMA_WARP = EPI_WARPS;
    using BI = BImage<CS>;
    static_assert(SETS * NACC * 2 * CS <= (int)TMEM_A, "accumulators");
    extern __shared__ __align__(1024) uint8_t sm[];
    int8_t* bsl = reinterpret_cast<int8_t*>(sm + L.b);
    double2* ysm = reinterpret_cast<double2*>(sm + L.y);
    double2* ksm = reinterpret_cast<double2*>(sm + L.k);
    int* ea_s = reinterpret_cast<int*>(sm + L.ea);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bar);
    // per set: full[NACC], empty[NACC], b_ready; then a_ready, p_free[NS]
    constexpr int PER_SET = 2 * NACC + 1;
    uint64_t *a_ready = bars + SETS * PER_SET, *p_free = a_ready + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L.tmem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = blockIdx.x * (SETS * CS);
    const int total = 4 * S;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SETS; ++s) {
#pragma unroll
            for (int b = 0; b < NACC; ++b) {
                mbar_init(bars + s * PER_SET + b, 1);
                mbar_init(bars + s * PER_SET + NACC + b, SET_WARPS);
            }
            mbar_init(bars + s * PER_SET + 2 * NACC, SET_WARPS);
        }
        mbar_init(a_ready, LOADERS);
#pragma unroll
        for (int p = 0; p < NS; ++p) mbar_init(p_free + p, 1);
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
        // =========================== epilogue warps: thread = (row, 8 columns of the warp's set) ===========================
        const int set = warp / SET_WARPS, qd = warp & 3, oc = (warp % SET_WARPS) >> 2;
        const int row = 32 * qd + lane;
        const uint32_t lane_base = ((uint32_t)(32 * qd) << 16) + (uint32_t)(set * NACC * 2 * CS);
        uint64_t *full = bars + set * PER_SET, *empty = full + NACC, *b_ready = full + 2 * NACC;
        unsigned* red = reinterpret_cast<unsigned*>(sm + L.red) + set * 4 * CS;
        int8_t* bset = bsl + set * (NS * BI::SLICE);
        const int cbase = set * CS + 8 * oc;  // first of the thread's columns within the CTA
        int eb[8];  // column exponents of the current stage vector (this thread's 8 columns)

        // slices the stage vector x (this thread's 8 elements) into shared memory; returns through eb the column scales
        auto slice_stage = [&](const double2 (&x)[8], int sidx) {
            (void)sidx;
            unsigned m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = __reduce_max_sync(0xffffffffu, max(abs_hi(x[j].x), abs_hi(x[j].y)));  // over the warp's 32 rows
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) red[qd * CS + 8 * oc + j] = m[j];
            }
            if (warp == 0) OZ_DBG(22);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + set), "n"(SET_WARPS * 32) : "memory");  // the set's warps
            if (warp == 0) OZ_DBG(23);
            OZ_DBG(48 + warp);
            unsigned wl[3][NS];  // columns 0-3 of the octet: digits of re, im, -im, slice p at [p - 1]
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {  // four columns at a time: one 32-bit word of every plane
                unsigned lo[3][4], hi[3][4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * hh + jj, c = 8 * oc + j;
                    const unsigned mm = max(max(red[c], red[CS + c]), max(red[2 * CS + c], red[3 * CS + c]));
                    eb[j] = slice_exponent_hi(mm);
                    const double scale = pow2(8 * NS - eb[j]);
                    const long long d0 = digits_of(x[j].x, scale), d1 = digits_of(x[j].y, scale), d2 = digits_of_negated(x[j].y, scale);
                    lo[0][jj] = (unsigned)d0, hi[0][jj] = (unsigned)((unsigned long long)d0 >> 32);
                    lo[1][jj] = (unsigned)d1, hi[1][jj] = (unsigned)((unsigned long long)d1 >> 32);
                    lo[2][jj] = (unsigned)d2, hi[2][jj] = (unsigned)((unsigned long long)d2 >> 32);
                }
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    unsigned wlo[4], whi[4];  // byte j of the digits = slice NS - j
                    transpose4(lo[part][0], lo[part][1], lo[part][2], lo[part][3], wlo);
                    transpose4(hi[part][0], hi[part][1], hi[part][2], hi[part][3], whi);
#pragma unroll
                    for (int p = 1; p <= NS; ++p) {
                        const int byte = NS - p;
                        const unsigned w = byte < 4 ? wlo[byte] : whi[byte - 4];
                        if (hh == 0) {
                            wl[part][p - 1] = w;
                        } else {
                            // parts (re, im, -im) -> image positions (1, 2, 0)
                            *reinterpret_cast<uint2*>(bset + (p - 1) * BI::SLICE + BI::off8(oc, row, (part + 1) % 3)) = make_uint2(wl[part][p - 1], w);
                        }
                    }
                }
            }
            if (warp == 0) OZ_DBG(24);
            OZ_DBG(32 + warp);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(b_ready);
        };

        double2 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cbase + j, col = col0 + c;
            double2 v = make_double2(0.0, 0.0);
            if (row < n && col < B) v = y[(size_t)row * ldy + col];
            x[j] = v;
            ysm[c * KD + row] = v;
            ksm[c * KD + row] = make_double2(0.0, 0.0);
        }
        slice_stage(x, -1);

        unsigned pf[NACC] = {};
#pragma unroll 1
        for (int sidx = 0; sidx < total; ++sidx) {
            const int stage = sidx & 3, entry = stage_entry(sidx);
            // groups g = NS + 1 .. 2 arrive in order of increasing weight 2^(-8g): T = sum_g D_g 256^(NS + 1 - g) in int64
            // (|D_g| < 2^24, |T| < 2^(24 + 8 (NS - 1)): exact), ONE conversion to fp64 per value at the end
            long long tr[8], ti[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tr[j] = ti[j] = 0;
            if (warp == 0) OZ_DBG(8);
#pragma unroll
            for (int g = NS + 1; g >= 2; --g) {
                const int b = acc_of_group(g), sh = 8 * (NS + 1 - g);
                mbar_wait(full + b, pf[b]);
                pf[b] ^= 1u;
                tc_fence_after();
                int vr[8], vi[8];
                tmem_ld8(lane_base + (uint32_t)(2 * b * CS + 8 * oc), vr);
                tmem_ld8(lane_base + (uint32_t)((2 * b + 1) * CS + 8 * oc), vi);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                if (lane == 0) mbar_arrive(empty + b);
                if (warp == 0) OZ_DBG(8 + g);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    tr[j] += (long long)vr[j] << sh;
                    ti[j] += (long long)vi[j] << sh;
                }
            }
            // k = 2^(eA[row] + eB[col]) acc; RK4 combine; next stage input
            const int ea = ea_s[(entry & 3) * KD + row];
            const StageCoef sc(stage, h);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = cbase + j;
                const double s = pow2(ea + eb[j] - 8 * (NS + 1));
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
            if (warp == EPI_WARPS - 1) OZ_DBG(25);
            if (sidx + 1 < total) slice_stage(x, sidx);
            if (warp == 0) OZ_DBG(21);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cbase + j, col = col0 + c;
            if (row < n && col < B) y[(size_t)row * ldy + col] = ysm[c * KD + row];
        }
    } else {
        // ============ MMA issuer (the first warp after the epilogue); then four generator loaders, one per TMEM lane quarter ============
        const int last_entry = 2 * S;
        if (warp != MMA_WARP) {
            // thread = row of the generator: its 2 NS slice planes (128 B each) go global -> registers -> TMEM (tcgen05.st), one
            // plane ahead, least significant slice first: slice p of the previous entry is free once group p + 1 of the
            // previous entry's last stage has completed (p_free[p - 1], one completion per entry change)
            const int qd = warp & 3;
            const int row = 32 * qd + lane;
            const uint32_t a_lane_base = TMEM_A + ((uint32_t)(32 * qd) << 16);
#pragma unroll 1
            for (int e = 0; e <= last_entry; ++e) {
                const int8_t* src = planes + (size_t)e * 2 * NS * KD * KD + (size_t)row * 16;  // chunk i of the row: + i KD 16
                const unsigned par = (unsigned)(e - 1) & 1u;
                // the table is far larger than L2 in a real solve (1000 steps: 330 MB of planes): the CTAs share out an L2
                // prefetch of the entry after the next (1280 lines of 128 B), so that every reload finds its planes in L2
                if (e + 2 <= last_entry) {
                    const int line = (int)blockIdx.x + (int)gridDim.x * ((warp - MMA_WARP - 1) * 32 + lane);
                    if (line < 2 * NS * KD) asm volatile("prefetch.global.L2 [%0];" ::"l"(planes + (size_t)(e + 2) * 2 * NS * KD * KD + (size_t)line * 128));
                }
                uint4 w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(NS - 1) * KD * KD) + i * KD);
                ea_s[(e & 3) * KD + row] = expo[(size_t)e * KD + row];
#pragma unroll 1
                for (int it = 0; it < 2 * NS; ++it) {
                    const int p = NS - (it >> 1), part = it & 1;  // plane (part, p); next: (1, p) or (0, p - 1)
                    const int pn = part == 0 ? p : (p > 1 ? p - 1 : 1), partn = part ^ 1;
                    uint4 wn[8];
                    const int8_t* nxt = src + (size_t)(partn * NS + pn - 1) * KD * KD;
#pragma unroll
                    for (int i = 0; i < 8; ++i) wn[i] = __ldg(reinterpret_cast<const uint4*>(nxt) + i * KD);
                    if (e > 0 && part == 0) {
                        mbar_wait(p_free + (p - 1), par);
                        tc_fence_after();
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t v[8] = {w[2 * i].x, w[2 * i].y, w[2 * i].z, w[2 * i].w, w[2 * i + 1].x, w[2 * i + 1].y, w[2 * i + 1].z, w[2 * i + 1].w};
                        tmem_st8(a_lane_base + (uint32_t)((part * NS + p - 1) * 32 + 8 * i), v);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = wn[i];
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __threadfence_block();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
            }
        } else {
            const uint32_t bs_addr = (uint32_t)__cvta_generic_to_shared(bsl);
            constexpr uint32_t LBO = BI::KG, SBO = 128;  // between the 8-row k groups / between the 16-column cores
            const uint64_t bdesc0 = smem_desc(bs_addr, LBO, SBO);
            const uint32_t bd_hi = (uint32_t)(bdesc0 >> 32);
            uint32_t bd_lo = (uint32_t)bdesc0;
            unsigned pe[SETS][NACC], pb = 0u, pa = 0u;
#pragma unroll
            for (int s = 0; s < SETS; ++s)
#pragma unroll
                for (int b = 0; b < NACC; ++b) pe[s][b] = 1u;
            // ONE elected thread runs the whole issue loop (waits, MMAs, commits): no election and predicate shuffling per MMA
            if (elect_one())
#pragma unroll 1
            for (int sidx = 0; sidx < total; ++sidx) {
                const int entry = stage_entry(sidx);
                const bool release = sidx + 1 < total && stage_entry(sidx + 1) != entry;  // the loaders refill behind this stage
                asm volatile("" : "+r"(bd_lo));  // opaque per stage: the descriptor words are base + immediate, not hoisted registers
                OZ_DBG(0);
                if (sidx == 0 || stage_entry(sidx - 1) != entry) {
                    mbar_wait(a_ready, pa);
                    pa ^= 1u;
                }
                OZ_DBG(1);
                // fully unrolled: every TMEM / shared-memory operand is a constant or a base plus a compile-time offset (the
                // descriptor's address field is bits [0, 14) of its low word in 16 B units: an offset never carries out of it)
#pragma unroll
                for (int set = 0; set < SETS; ++set) {
                    uint64_t *full = bars + set * PER_SET, *empty = full + NACC, *b_ready = full + 2 * NACC;
                    mbar_wait(b_ready, pb);
                    tc_fence_after();
                    if (set == 0) OZ_DBG(2);
#pragma unroll
                    for (int g = NS + 1; g >= 2; --g) {
                        const int b = acc_of_group(g);
                        mbar_wait(empty + b, pe[set][b]);
                        pe[set][b] ^= 1u;
                        tc_fence_after();
                        const uint32_t d = (uint32_t)((set * NACC + b) * 2 * CS);  // (re | im): 2 CS accumulator columns
#pragma unroll
                        for (int p = 1; p < g; ++p) {
                            const int q = g - p;
                            const uint32_t a_re = TMEM_A + (uint32_t)((0 * NS + (p - 1)) * 32), a_im = TMEM_A + (uint32_t)((1 * NS + (p - 1)) * 32);
#pragma unroll
                            for (int ks = 0; ks < NKS; ++ks) {  // k chunks past n hold zeros
                                const uint32_t b1 = bd_lo + (uint32_t)(((set * NS + q - 1) * BI::SLICE + BI::RE_IM + ks * 4 * (int)LBO) >> 4);  // (re | im)
                                const uint32_t b2 = bd_lo + (uint32_t)(((set * NS + q - 1) * BI::SLICE + ks * 4 * (int)LBO) >> 4);               // (-im | re)
                                mma_ts1<idesc_for(2 * CS)>(d, a_re + 8 * ks, b1, bd_hi, (p == 1 && ks == 0) ? 0u : 1u);
                                mma_ts1<idesc_for(2 * CS)>(d, a_im + 8 * ks, b2, bd_hi, 1u);
                            }
                        }
                        umma_commit1(full + b);
                        if (release && set == SETS - 1) umma_commit1(p_free + (g - 2));  // slice plane g - 1 is read by no later group
                    }
                    if (set == 0) OZ_DBG(3);
                }
                pb ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace

bool rk4_ozaki_supported(int n) { return n >= OZ_MIN_N && n <= 128; }

// The emulated path is the faster one once its single wave of CTAs beats the DMMA kernels' time for the batch (measured at
// n = 128: 15.8 us per step for any B <= 2368 -- 16 columns per CTA -- and 21.8 us up to 4736, against 9.0 / 17.9 / 30.2 /
// 51.8 us of the DMMA kernels at B = 512 / 1024 / 2048 / 4096).  Smaller n pay the padding to 128 rows (n <= 96 runs three
// k chunks instead of four): at B = 4096 the emulation wins 1.93x at n = 100, 1.87x at 96, 1.60x at 80; at B = 2048 1.07x
// at n = 96 and 0.88x at 80 (profiles/r02_r_ozaki_small_n.jsonl).  QDB_RK4_INT8=0 keeps every batch on the fp64 DMMA kernels.
bool rk4_ozaki_preferred(int n, int B) {
    static const bool enabled = [] {
        const char* e = getenv("QDB_RK4_INT8");
        return !(e && e[0] == '0');
    }();
    if (!enabled || !rk4_ozaki_supported(n)) return false;
    if (n >= 121) return B >= 960;
    if (n >= 96) return B >= 1536;
    return B > sm_count() * 16;  // n >= 65
}

void rk4_ozaki_debug(long long* host64) { cudaMemcpyFromSymbol(host64, g_oz_dbg, sizeof(long long) * 64); }

// bytes of the int8 slice planes + row exponents of T table entries
size_t rk4_ozaki_table_bytes(int T) { return (size_t)T * (2 * NS * KD * KD + KD * sizeof(int)); }

// gen: T generator table entries, row-major n x n (QDB_LAYOUT_ROWMAJOR) or packed (QDB_LAYOUT_PACKED: 128 x 128 for these n)
// -> ws: int8 slice planes + row exponents (rk4_ozaki_table_bytes(T))
int launch_ozaki_slice(int n, int T, const double2* gen, int gen_layout, void* ws, cudaStream_t st) {
    int8_t* planes = reinterpret_cast<int8_t*>(ws);
    int* expo = reinterpret_cast<int*>(planes + (size_t)T * 2 * NS * KD * KD);
    for (int t0 = 0; t0 < T; t0 += kMaxGridY) {
        const int Tc = T - t0 < kMaxGridY ? T - t0 : kMaxGridY;
        if (gen_layout == QDB_LAYOUT_PACKED)
            ozaki_gslice_kernel<true><<<dim3(KD, Tc), 128, 0, st>>>(n, gen + (size_t)t0 * round_up8(n) * round_up16(n), planes + (size_t)t0 * 2 * NS * KD * KD,
                                                                    expo + (size_t)t0 * KD);
        else
            ozaki_gslice_kernel<false><<<dim3(KD, Tc), 128, 0, st>>>(n, gen + (size_t)t0 * n * n, planes + (size_t)t0 * 2 * NS * KD * KD,
                                                                     expo + (size_t)t0 * KD);
        QDB_LAUNCH_CHECK("ozaki_gslice_kernel");
    }
    return QDB_OK;
}

// S steps from the sliced table of 2S+1 entries in ws; gen == nullptr: ws was filled by launch_ozaki_slice
int launch_rk4_ozaki(int n, int B, int S, const double2* gen, int gen_layout, double h, double2* y, int ldy, void* ws, cudaStream_t st) {
    const int T = 2 * S + 1;
    if (gen) {
        const int rc = launch_ozaki_slice(n, T, gen, gen_layout, ws, st);
        if (rc != QDB_OK) return rc;
    }
    int8_t* planes = reinterpret_cast<int8_t*>(ws);
    int* expo = reinterpret_cast<int*>(planes + (size_t)T * 2 * NS * KD * KD);
#ifdef QDB_OZ_TIMELINE
    if (const char* ds = getenv("QDB_OZ_DBG_STAGE")) {
        const int v = atoi(ds);
        cudaMemcpyToSymbolAsync(g_oz_dbg_stage, &v, sizeof(int), 0, cudaMemcpyHostToDevice, st);
    }
#endif
    // 32 columns per CTA once 16 per CTA would no longer fit one wave of the SMs
    auto go = [&](auto kernel, int cs, int warps) -> int {
        const Smem L(cs, 1);
        QDB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        kernel<<<(B + cs - 1) / cs, warps * 32, L.total, st>>>(n, B, S, planes, expo, h, y, ldy);
        return QDB_OK;
    };
    const bool small = B <= sm_count() * 16, k3 = n <= 96;
    int rc;
    if (small) rc = k3 ? go(rk4_ozaki_kernel<16, 1, 3>, 16, 8 + 1 + LOADERS) : go(rk4_ozaki_kernel<16, 1, 4>, 16, 8 + 1 + LOADERS);
    else rc = k3 ? go(rk4_ozaki_kernel<32, 1, 3>, 32, 16 + 1 + LOADERS) : go(rk4_ozaki_kernel<32, 1, 4>, 32, 16 + 1 + LOADERS);
    if (rc != QDB_OK) return rc;
    QDB_LAUNCH_CHECK("rk4_ozaki_kernel");
    return QDB_OK;
}

}  // namespace qdb
