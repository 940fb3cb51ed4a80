// Fused sweep-mode RK4 for SMALL operators (n <= 32): S steps x 4 stages in one launch
// (SURVEY.md 8(a) rows a1 + a2 + a3 + a7 per column; BASELINE.json configs[1]: dim 32, 8 drive operators,
// batch-1024 amplitude sweep).
//
// At n = 32 the generic rk4_sweep_kernel is instruction bound: one 8x8 tile per warp leaves 4 DMMAs per
// (k-tile, operator) pass against ~40 address/bookkeeping instructions, and a batch of 1024 columns gives every
// SM a single column octet, i.e. four warps (ncu: tensor pipe 27 %, 'wait' + 'selected' 68 % of the stall samples).
// This kernel restructures the same arithmetic for that regime:
//   * all K+1 operators are staged ONCE per launch into shared memory in DMMA A-fragment order (147 KiB at
//     n = 32, K = 8) by TMA bulk copies (cp.async.bulk -> one mbarrier) -- the A operand is an LDS.128 at an
//     immediate offset;
//   * per-column signal values scale the STAGE VECTOR, not the fragment: the epilogue writes K+1 scaled copies
//     z_j = c_j[col] * u (one per operator; 36 KiB) so that the k loop is LDS + DMMA only -- the generic kernel
//     re-scales every B fragment once per row warp with DMULs that compete with the DMMAs for the fp64 pipe;
//   * the operator sum is split over JS warp groups (a CTA is RT row warps x JS operator groups); partial
//     accumulators meet in shared memory once per stage, so a single column octet keeps 8 warps busy;
//   * y and the RK4 k-sum of the two elements a thread owns stay in registers for the whole launch.
// Roofline: fp64 tensor pipe; executes 4 (K+1) 8 n^2 flops per column per step for the algorithmic
// 4 ((4K+8) n^2 + 12 n) + 28 n.
#include <cstdlib>

#include "qdb_common.cuh"
#include "rk4_device.cuh"

namespace qdb {

namespace {

template <int RT, int JS>
__global__ void __launch_bounds__(32 * RT * JS, 1)
rk4_sweep_small_kernel(int n, int K, int B, int S, const double2* __restrict__ stat /*packed or null*/,
                       const double2* __restrict__ ops /*[K] packed*/, const double* __restrict__ coeff /*[2S+1][K][ldc]*/,
                       int ldc, const double* __restrict__ mu, const double* __restrict__ times /*[2S+1]*/, double h,
                       double2* __restrict__ y, int ldy) {
    constexpr int KT = RT <= 2 ? 4 : 8;        // k-tiles of the packed layout (kpad = 16 or 32)
    constexpr int ENTRY = 32 * RT * KT;        // elements of one packed operator (npad * kpad)
    constexpr int NTHR = 32 * RT * JS;
    constexpr int KMAX = 16;                   // per-thread coefficient registers (host checks K <= KMAX)
    extern __shared__ __align__(16) double2 sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int rt = warp % RT, js = warp / RT;
    const int has_static = stat != nullptr ? 1 : 0;
    const int J = K + has_static;
    double2* ops_s = sm;                       // [J][ENTRY]
    double2* z = ops_s + (size_t)J * ENTRY;    // [J][KT][32] scaled stage vectors, B-fragment order, swizzled
    double2* part = z + (size_t)J * KT * 32;   // [JS-1][RT][2][32] partial accumulators
    const int col0 = blockIdx.x * 8;
    const bool framed = (mu != nullptr);
    const bool owner = (js == 0);              // owner warps hold y / k-sum and run the epilogue

    // ---- stage the operators (static first): one TMA bulk copy per operator, all counted on one mbarrier ----
    uint64_t* tma_bar = reinterpret_cast<uint64_t*>(part + (JS - 1) * RT * 64);
    if (tid == 0) {
        mbar_init(tma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(tma_bar, (unsigned)(J * ENTRY * sizeof(double2)));
        for (int j = 0; j < J; ++j)
            tma_bulk_g2s(ops_s + (size_t)j * ENTRY, (has_static && j == 0) ? stat : ops + (size_t)(j - has_static) * ENTRY,
                         (unsigned)(ENTRY * sizeof(double2)), tma_bar);
    }
    for (int idx = tid; idx < J * KT * 32; idx += NTHR) z[idx] = make_double2(0.0, 0.0);  // overlaps the copies
    mbar_wait(tma_bar, 0);
    __syncthreads();

    // ---- this thread's two state elements: row 8 rt + g, columns col0 + 2q + i ----
    const int row = 8 * rt + g;
    const double mu_row = (framed && row < n) ? mu[row] : 0.0;
    double2 yv[2], ks[2];
    int zpos[2];  // slot of the element inside one [KT][32] plane
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int col = col0 + 2 * q + i;
        yv[i] = (owner && row < n && col < B) ? y[(size_t)row * ldy + col] : make_double2(0.0, 0.0);
        ks[i] = make_double2(0.0, 0.0);
        zpos[i] = (2 * rt + (g >> 2)) * 32 + frag_swizzle((g & 3) + 4 * (2 * q + i));
    }
    double2 ph = framed ? frame_phase(mu_row, times[0]) : make_double2(1.0, 0.0);

    // per-column signal values of this thread's two columns at a table entry (issued early: the loads fly during
    // the k loop) and the K+1 scaled copies z_j = c_j * u of the next stage input
    auto load_coeffs = [&](int entry, double (&cj)[KMAX][2]) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int col = col0 + 2 * q + i;
                cj[j][i] = (j < K && col < B) ? __ldg(coeff + ((size_t)entry * K + j) * ldc + col) : 0.0;
            }
    };
    auto write_planes = [&](const double (&cj)[KMAX][2], const double2 (&u)[2]) {
        if (has_static) {
#pragma unroll
            for (int i = 0; i < 2; ++i) z[zpos[i]] = u[i];
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < K) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    z[(size_t)(j + has_static) * KT * 32 + zpos[i]] = make_double2(cj[j][i] * u[i].x, cj[j][i] * u[i].y);
            }
    };

    if (owner) {
        double2 u[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) u[i] = cmul(ph, yv[i]);  // pre-phase
        double c0[KMAX][2];
        load_coeffs(0, c0);
        write_planes(c0, u);
    }
    __syncthreads();

    const double2* a_base = ops_s + (size_t)rt * KT * 32 + lane;
    const double2* z_base = z + frag_swizzle(lane);
    const int total_stages = 4 * S;
#pragma unroll 1
    for (int sidx = 0; sidx < total_stages; ++sidx) {
        const int step = sidx >> 2, stage = sidx & 3;
        const int entry = 2 * step + (stage == 0 ? 0 : (stage == 3 ? 2 : 1));
        const int nstage = (stage + 1) & 3, nstep = step + (stage == 3 ? 1 : 0);
        const int next_entry = (sidx + 1 < total_stages) ? 2 * nstep + (nstage == 0 ? 0 : (nstage == 3 ? 2 : 1)) : entry;

        double cn[KMAX][2];
        if (owner) load_coeffs(next_entry, cn);

        // two accumulator sets (even / odd k-tiles): four independent DMMA chains per warp
        double cr[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, ci[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 1
        for (int j = js; j < J; j += JS) {
            const double2* A = a_base + (size_t)j * ENTRY;
            const double2* Z = z_base + (size_t)j * KT * 32;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const double2 a = A[kt * 32], b = Z[kt * 32];
                dmma(cr[kt & 1][0], cr[kt & 1][1], a.x, b.x);
                dmma(ci[kt & 1][0], ci[kt & 1][1], a.x, b.y);
                dmma(cr[kt & 1][0], cr[kt & 1][1], -a.y, b.y);
                dmma(ci[kt & 1][0], ci[kt & 1][1], a.y, b.x);
            }
        }
        double2 acc[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) acc[i] = make_double2(cr[0][i] + cr[1][i], ci[0][i] + ci[1][i]);
        if (JS > 1 && !owner) {
#pragma unroll
            for (int i = 0; i < 2; ++i) part[(((js - 1) * RT + rt) * 2 + i) * 32 + lane] = acc[i];
        }
        __syncthreads();  // partial sums visible; every warp is done reading z

        if (owner) {
#pragma unroll
            for (int o = 1; o < JS; ++o)
#pragma unroll
                for (int i = 0; i < 2; ++i) acc[i] = cadd(acc[i], part[(((o - 1) * RT + rt) * 2 + i) * 32 + lane]);
            const double2 ph_next = (framed && next_entry != entry) ? frame_phase(mu_row, times[next_entry]) : ph;
            const StageCoef sc(stage, h);
            double2 u[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double2 k = cmul_conj_a(ph, acc[i]);  // post-phase
                ks[i].x = sc.keep * ks[i].x + sc.wk * k.x;
                ks[i].y = sc.keep * ks[i].y + sc.wk * k.y;
                const double v_r = sc.last ? ks[i].x : k.x, v_i = sc.last ? ks[i].y : k.y;
                const double2 nxt = make_double2(yv[i].x + sc.astep * v_r, yv[i].y + sc.astep * v_i);
                if (sc.last) yv[i] = nxt;
                u[i] = cmul(ph_next, nxt);
            }
            ph = ph_next;
            write_planes(cn, u);
        }
        __syncthreads();  // next stage's planes visible
    }

    if (owner) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int col = col0 + 2 * q + i;
            if (row < n && col < B) y[(size_t)row * ldy + col] = yv[i];
        }
    }
}

template <int RT, int JS>
int launch_small_t(int n, int K, int B, int S, const double2* stat, const double2* ops, const double* coeff, int ldc,
                   const double* mu, const double* times, double h, double2* y, int ldy, size_t smem, cudaStream_t st) {
    auto kern = rk4_sweep_small_kernel<RT, JS>;
    QDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (B + 7) / 8;
    kern<<<grid, 32 * RT * JS, smem, st>>>(n, K, B, S, stat, ops, coeff, ldc, mu, times, h, y, ldy);
    QDB_LAUNCH_CHECK("rk4_sweep_small_kernel");
    return QDB_OK;
}

struct SmallCfg {
    int RT, JS;
    size_t smem;
};

bool pick_small(int n, int K, bool has_static, SmallCfg& c) {
    const char* off = getenv("QDB_NO_SMALL_SWEEP");
    if (off && off[0] == '1') return false;
    if (n < 1 || n > 32 || K < 1 || K > 16) return false;
    const int RT = (n + 7) / 8, KT = RT <= 2 ? 4 : 8;
    // the packed layout pads k to 16: row tiles 1-2 <-> kpad 16, row tiles 3-4 <-> kpad 32
    if (round_up16(n) / 4 != KT) return false;
    const int J = K + (has_static ? 1 : 0);
    int JS = RT <= 2 ? 4 : 2;
    while (JS > J) JS /= 2;
    const size_t smem = ((size_t)J * 32 * RT * KT + (size_t)J * KT * 32 + (size_t)(JS - 1) * RT * 64) * sizeof(double2) +
                        16 /*mbarrier*/;
    if (smem > 227 * 1024) return false;
    c.RT = RT;
    c.JS = JS;
    c.smem = smem;
    return true;
}

}  // namespace

bool rk4_sweep_small_supported(int n, int K, bool has_static) {
    SmallCfg c;
    return pick_small(n, K, has_static, c);
}

int launch_rk4_sweep_small(int n, int K, int B, int S, const double2* stat, const double2* ops, const double* coeff, int ldc,
                           const double* mu, const double* times, double h, double2* y, int ldy, cudaStream_t st) {
    SmallCfg c;
    if (!pick_small(n, K, stat != nullptr, c)) {
        set_error("rk4 small sweep: unsupported shape n=%d K=%d", n, K);
        return QDB_E_UNSUPPORTED;
    }
#define SMALL(RTv, JSv) \
    if (c.RT == RTv && c.JS == JSv) return launch_small_t<RTv, JSv>(n, K, B, S, stat, ops, coeff, ldc, mu, times, h, y, ldy, c.smem, st)
    SMALL(1, 1); SMALL(1, 2); SMALL(1, 4);
    SMALL(2, 1); SMALL(2, 2); SMALL(2, 4);
    SMALL(3, 1); SMALL(3, 2);
    SMALL(4, 1); SMALL(4, 2);
#undef SMALL
    set_error("rk4 small sweep: no kernel for RT=%d JS=%d", c.RT, c.JS);
    return QDB_E_UNSUPPORTED;
}

}  // namespace qdb
