// extern "C" entry points of libqdb.so (see include/qdb.h for the contract of each).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "qdb_common.cuh"

namespace qdb {

static thread_local char g_err[512] = "no error";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;  // B200
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

}  // namespace qdb

using namespace qdb;

static inline const double2* D2(const qdb_c128* p) { return reinterpret_cast<const double2*>(p); }
static inline double2* D2(qdb_c128* p) { return reinterpret_cast<double2*>(p); }

#define QDB_REQUIRE(cond, ...)      \
    do {                            \
        if (!(cond)) {              \
            set_error(__VA_ARGS__); \
            return QDB_E_ARG;       \
        }                           \
    } while (0)

extern "C" {

const char* qdb_last_error_string(void) { return g_err; }
int qdb_version(void) { return 100; }
int qdb_npad(int n) { return round_up8(n); }
size_t qdb_packed_elems(int n) { return (size_t)round_up8(n) * round_up16(n); }
size_t qdb_table_entry_bytes(int n, int layout) {
    if (layout == QDB_LAYOUT_ROWMAJOR) return (size_t)n * n * sizeof(double2);
    return qdb_packed_elems(n) * (layout == QDB_LAYOUT_PACKED3M ? 24 : 16);
}
int qdb_rk4_table_layout(int n, int B) {
    if (n < 1 || B < 1 || !rk4_fused_supported(n)) return QDB_LAYOUT_PACKED;
    return rk4_fused_table_layout(n, B);
}
unsigned long long qdb_launch_count(void) { return g_launches.load(); }

size_t qdb_workspace_bytes(int kind, int n, int K, int B, int S) {
    const size_t n2 = (size_t)n * n * sizeof(double2);
    const size_t np2 = qdb_packed_elems(n) * 24;  // a PACKED3M table entry (the larger of the two table layouts)
    const size_t yb = (size_t)n * B * sizeof(double2);
    if (S < 1) S = 1;
    switch (kind) {
        case QDB_WS_RHS:
            return align_up(n2) + 2 * align_up((size_t)n * sizeof(double2));
        case QDB_WS_RK4:
            if (rk4_fused_supported(n)) {
                // shared signals: generator table + stage times; per-column signals: stage times + the formed-sweep
                // operator copy (the caller does not say which mode it will ask for: the larger of the two)
                size_t entry = np2;  // ... or a PACKED entry plus its int8 slice planes (rk4_ozaki_kernel)
                if (rk4_ozaki_supported(n) && qdb_packed_elems(n) * 16 + rk4_ozaki_table_bytes(1) > entry)
                    entry = qdb_packed_elems(n) * 16 + rk4_ozaki_table_bytes(1);
                const size_t shared = align_up((size_t)(2 * S + 1) * entry) + align_up((size_t)(2 * S + 1) * sizeof(double));
                const size_t sweep = align_up((size_t)(2 * S + 1) * sizeof(double)) +
                                     (rk4_sweepf_supported(n, K) ? align_up(rk4_sweepf_workspace_bytes(n, K)) : 0);
                return shared > sweep ? shared : sweep;
            }
            // generic path: three generators + three state buffers (shared signals), or four state buffers + two phase
            // vectors (per-column signals: one GEMM per operator and stage with the column scale in its epilogue)
            return 3 * align_up(n2) + 4 * align_up(yb) + 2 * align_up((size_t)n * sizeof(double2)) + align_up(3 * sizeof(double));
        case QDB_WS_EXPM: {
            // S = 1: one step at a time.  S > 1: room to build the propagators of up to S steps side by side (batched
            // Taylor exponentials) before they are applied one after the other -- capped at 1 GiB of matrices
            const size_t per = 7 * align_up(n2);
            size_t chunk = (size_t)S;
            const size_t cap = ((size_t)1 << 30) / per;
            if (chunk > cap) chunk = cap;
            if (chunk < 1) chunk = 1;
            return chunk * per + align_up(yb) + align_up((size_t)S * sizeof(double));
        }
        case QDB_WS_PROP:
            return propagator_workspace_bytes(n, S);
        case QDB_WS_MAGNUS: {  // per step of a chunk: node generators (3), Magnus temporaries (7), A, Taylor (5), P
            const size_t per = 17 * align_up(n2);
            size_t chunk = (size_t)S;
            const size_t cap = ((size_t)1 << 30) / per;
            if (chunk > cap) chunk = cap;
            if (chunk < 1) chunk = 1;
            return chunk * per + align_up(yb) + align_up((size_t)3 * S * sizeof(double));
        }
        default:
            return 0;
    }
}

int qdb_pack_operators(int n, int count, const qdb_c128* src, qdb_c128* dst, void* stream) {
    QDB_REQUIRE(n >= 1 && count >= 0, "qdb_pack_operators: bad n=%d count=%d", n, count);
    if (count == 0) return QDB_OK;
    QDB_REQUIRE(src && dst, "qdb_pack_operators: null pointer");
    return launch_pack(n, count, D2(src), D2(dst), (cudaStream_t)stream);
}

int qdb_generator_c128(int n, int K, int T, int layout, const qdb_c128* ops, const qdb_c128* stat,
                       const double* coeff, int coeff_complex, const double* mu, const double* times,
                       double scale, qdb_c128* out, void* stream) {
    QDB_REQUIRE(n >= 1 && K >= 0 && T >= 0, "qdb_generator_c128: bad n=%d K=%d T=%d", n, K, T);
    QDB_REQUIRE(layout == QDB_LAYOUT_ROWMAJOR || layout == QDB_LAYOUT_PACKED || layout == QDB_LAYOUT_PACKED3M,
                "qdb_generator_c128: bad layout %d", layout);
    QDB_REQUIRE(out, "qdb_generator_c128: null output");
    QDB_REQUIRE(stat || (ops && K > 0), "qdb_generator_c128: neither static operator nor operators given");
    QDB_REQUIRE(K == 0 || (ops && coeff), "qdb_generator_c128: K=%d but ops/coeff missing", K);
    QDB_REQUIRE(!mu || times, "qdb_generator_c128: frame given without times");
    if (T == 0) return QDB_OK;
    return launch_generator(n, K, T, layout, D2(ops), D2(stat), coeff, coeff_complex, mu, times, 0.0, scale, D2(out),
                            (cudaStream_t)stream);
}

int qdb_frame_apply_c128(int n, int B, const double* mu, double t, int conj_phase, const qdb_c128* y_in, qdb_c128* y_out,
                         int ldy, void* stream) {
    QDB_REQUIRE(n >= 0 && B >= 0, "qdb_frame_apply_c128: bad n=%d B=%d", n, B);
    if (n == 0 || B == 0) return QDB_OK;
    QDB_REQUIRE(mu && y_in && y_out && ldy >= B, "qdb_frame_apply_c128: null pointer / bad ldy");
    return launch_frame_apply(n, B, mu, t, conj_phase, D2(y_in), D2(y_out), ldy, (cudaStream_t)stream);
}

int qdb_zgemm_c128(int M, int N, int Kd, const qdb_c128* A, int lda, const qdb_c128* Bm, int ldb, qdb_c128* C, int ldc,
                   qdb_c128 alpha, qdb_c128 beta, const double* colscale, const qdb_c128* pre, const qdb_c128* post,
                   void* stream) {
    QDB_REQUIRE(M >= 0 && N >= 0 && Kd >= 0, "qdb_zgemm_c128: negative dimension");
    if (M == 0 || N == 0) return QDB_OK;
    QDB_REQUIRE(A && Bm && C, "qdb_zgemm_c128: null pointer");
    QDB_REQUIRE(lda >= Kd && ldb >= N && ldc >= N, "qdb_zgemm_c128: leading dimension too small");
    return launch_zgemm(M, N, Kd, D2(A), lda, D2(Bm), ldb, D2(C), ldc, make_double2(alpha.re, alpha.im),
                        make_double2(beta.re, beta.im), colscale, D2(pre), D2(post), (cudaStream_t)stream);
}

int qdb_rhs_c128(int n, int K, int B, const qdb_c128* ops, const qdb_c128* stat, const double* coeff, int coeff_per_col,
                 int ldc, const double* mu, double t, const qdb_c128* y_in, qdb_c128* y_out, int ldy, void* workspace,
                 size_t ws_bytes, void* stream) {
    QDB_REQUIRE(n >= 1 && K >= 0 && B >= 0, "qdb_rhs_c128: bad n=%d K=%d B=%d", n, K, B);
    QDB_REQUIRE(stat || (ops && K > 0), "qdb_rhs_c128: neither static operator nor operators given");
    QDB_REQUIRE(K == 0 || (ops && coeff), "qdb_rhs_c128: K=%d but ops/coeff missing", K);
    if (B == 0) return QDB_OK;
    QDB_REQUIRE(y_in && y_out && ldy >= B, "qdb_rhs_c128: bad state pointers / ldy");
    QDB_REQUIRE(y_in != y_out, "qdb_rhs_c128: y_in and y_out must not alias");
    if (ws_bytes < qdb_workspace_bytes(QDB_WS_RHS, n, K, B, 1) || !workspace) {
        set_error("qdb_rhs_c128: workspace too small (%zu < %zu)", ws_bytes, qdb_workspace_bytes(QDB_WS_RHS, n, K, B, 1));
        return QDB_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    double2* G = (double2*)ws;
    double2* pre = (double2*)(ws + align_up((size_t)n * n * sizeof(double2)));
    double2* post = (double2*)((char*)pre + align_up((size_t)n * sizeof(double2)));
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
    int rc;
    if (!coeff_per_col) {
        // G_frame(t) once (a1 + a5), then one GEMM (a2); the frame phases ride inside G_frame
        rc = launch_generator(n, K, 1, QDB_LAYOUT_ROWMAJOR, D2(ops), D2(stat), coeff, 0, mu, nullptr, t, 1.0, G, st);
        if (rc != QDB_OK) return rc;
        return launch_zgemm(n, B, n, G, n, D2(y_in), ldy, D2(y_out), ldy, one, zero, nullptr, nullptr, nullptr, st);
    }
    QDB_REQUIRE(ldc >= B, "qdb_rhs_c128: ldc < B");
    const double2 *prep = nullptr, *postp = nullptr;
    if (mu) {
        rc = launch_phase_vectors(n, mu, t, pre, post, st);
        if (rc != QDB_OK) return rc;
        prep = pre;
        postp = post;
    }
    bool first = true;
    if (stat) {
        rc = launch_zgemm(n, B, n, D2(stat), n, D2(y_in), ldy, D2(y_out), ldy, one, zero, nullptr, prep, postp, st);
        if (rc != QDB_OK) return rc;
        first = false;
    }
    for (int j = 0; j < K; ++j) {
        rc = launch_zgemm(n, B, n, D2(ops) + (size_t)j * n * n, n, D2(y_in), ldy, D2(y_out), ldy, one, first ? zero : one,
                          coeff + (size_t)j * ldc, prep, postp, st);
        if (rc != QDB_OK) return rc;
        first = false;
    }
    return QDB_OK;
}

int qdb_rk4_steps_c128(int n, int K, int B, int S, const qdb_c128* ops_rm, const qdb_c128* stat_rm,
                       const qdb_c128* ops_packed, const qdb_c128* stat_packed, const double* coeff, int sig_mode, int ldc,
                       const double* mu, const double* times_host, double h, qdb_c128* y, int ldy, void* workspace,
                       size_t ws_bytes, void* stream) {
    QDB_REQUIRE(n >= 1 && K >= 0 && B >= 0 && S >= 0, "qdb_rk4_steps_c128: bad n=%d K=%d B=%d S=%d", n, K, B, S);
    QDB_REQUIRE(sig_mode == 0 || sig_mode == 1, "qdb_rk4_steps_c128: bad sig_mode %d", sig_mode);
    if (B == 0 || S == 0) return QDB_OK;
    QDB_REQUIRE(y && ldy >= B, "qdb_rk4_steps_c128: bad state pointer / ldy");
    QDB_REQUIRE(K == 0 || coeff, "qdb_rk4_steps_c128: K=%d but no signal table", K);
    QDB_REQUIRE(!mu || times_host, "qdb_rk4_steps_c128: frame given without stage times");
    QDB_REQUIRE(workspace, "qdb_rk4_steps_c128: null workspace");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const bool fused = rk4_fused_supported(n);
    int rc;

    if (fused) {
        QDB_REQUIRE(stat_packed || (ops_packed && K > 0), "qdb_rk4_steps_c128: packed operators missing");
        const size_t np2 = qdb_packed_elems(n);
        if (sig_mode == 1) {
            QDB_REQUIRE(ldc >= B, "qdb_rk4_steps_c128: ldc < B");
            // only the stage times go to the device; operators are used as stored
            const size_t need = align_up((size_t)(2 * S + 1) * sizeof(double));
            if (ws_bytes < need) {
                set_error("qdb_rk4_steps_c128: workspace too small (%zu < %zu)", ws_bytes, need);
                return QDB_E_WORKSPACE;
            }
            double* times_dev = (double*)ws;
            if (mu) QDB_CUDA(cudaMemcpyAsync(times_dev, times_host, (size_t)(2 * S + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
            // the formed-generator kernel needs room for its operator copy behind the stage times; without it the
            // operator-pass kernels run
            void* fws = nullptr;
            if (rk4_sweepf_supported(n, K) && ws_bytes >= need + align_up(rk4_sweepf_workspace_bytes(n, K))) fws = ws + need;
            return launch_rk4_fused_sweep(n, K, B, S, D2(stat_packed), D2(ops_packed), coeff, ldc, mu, times_dev, h, D2(y), ldy, fws, st);
        }
        // shared signals: chunk the step loop so that the generator table fits the workspace
        // (large batches at n = 80..128: the int8 tensor-core emulation, whose slice planes sit behind the packed table)
        const bool int8_path = rk4_ozaki_preferred(n, B);
        const int table_layout = int8_path ? QDB_LAYOUT_PACKED : rk4_fused_table_layout(n, B);
        const size_t per_entry = qdb_table_entry_bytes(n, table_layout) + (int8_path ? rk4_ozaki_table_bytes(1) : 0);
        // largest Sc whose 2 Sc + 1 table entries (+ their stage times) fit the workspace
        auto fits = [&](long long Sc) {
            const size_t E = (size_t)(2 * Sc + 1);
            return align_up(E * per_entry) + align_up(E * sizeof(double)) <= ws_bytes;
        };
        long long Sc_max = (long long)(ws_bytes / per_entry + 1) / 2;
        while (Sc_max >= 1 && !fits(Sc_max)) --Sc_max;
        if (Sc_max < 1) {
            set_error("qdb_rk4_steps_c128: workspace too small (%zu < %zu)", ws_bytes, qdb_workspace_bytes(QDB_WS_RK4, n, K, B, 1));
            return QDB_E_WORKSPACE;
        }
        if (Sc_max > S) Sc_max = S;
        double2* table = (double2*)ws;
        double* times_dev = (double*)(ws + align_up((size_t)(2 * Sc_max + 1) * per_entry));
        void* planes_ws = ws + (size_t)(2 * Sc_max + 1) * qdb_table_entry_bytes(n, table_layout);
        for (int s0 = 0; s0 < S; s0 += (int)Sc_max) {
            const int Sc = (S - s0 < Sc_max) ? S - s0 : (int)Sc_max;
            const int T = 2 * Sc + 1;
            if (mu) QDB_CUDA(cudaMemcpyAsync(times_dev, times_host + 2 * s0, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, st));
            rc = launch_generator(n, K, T, table_layout, D2(ops_packed), D2(stat_packed),
                                  coeff ? coeff + (size_t)2 * s0 * K : nullptr, 0, mu, times_dev, 0.0, 1.0, table, st);
            if (rc != QDB_OK) return rc;
            rc = int8_path ? launch_rk4_ozaki(n, B, Sc, table, table_layout, h, D2(y), ldy, planes_ws, st)
                           : launch_rk4_fused_shared(n, B, Sc, table, table_layout, h, D2(y), ldy, st);
            if (rc != QDB_OK) return rc;
        }
        return QDB_OK;
    }

    // ---- generic path for large n ----
    QDB_REQUIRE(stat_rm || (ops_rm && K > 0), "qdb_rk4_steps_c128: row-major operators missing");
    QDB_REQUIRE(ldy == B, "qdb_rk4_steps_c128: generic path needs ldy == B");
    if (ws_bytes < qdb_workspace_bytes(QDB_WS_RK4, n, K, B, 1)) {
        set_error("qdb_rk4_steps_c128: workspace too small (%zu < %zu)", ws_bytes, qdb_workspace_bytes(QDB_WS_RK4, n, K, B, 1));
        return QDB_E_WORKSPACE;
    }
    if (sig_mode == 1) {
        // per-column signals, n > 256 (vectorised Lindblad sweeps): sum_j G_j (c_jb y_b) -- one DMMA GEMM per operator and
        // stage, the column's signal value as the GEMM's column scale, the frame phases as its pre / post vectors, accumulated
        // in k; then the RK4 combine.  Same arithmetic as rk4_sweep_kernel, operands in HBM instead of on chip.
        QDB_REQUIRE(ldc >= B, "qdb_rk4_steps_c128: ldc < B");
        const size_t ybs = align_up((size_t)n * B * sizeof(double2)), pv = align_up((size_t)n * sizeof(double2));
        double2* kbuf = (double2*)ws;
        double2* ya = (double2*)(ws + ybs);
        double2* yb2 = (double2*)(ws + 2 * ybs);
        double2* acc = (double2*)(ws + 3 * ybs);
        double2* pre = (double2*)(ws + 4 * ybs);
        double2* post = (double2*)(ws + 4 * ybs + pv);
        const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
        const size_t nn = (size_t)n * n, cnt = (size_t)n * B;
        double2* Y = D2(y);
        auto stage = [&](int entry, const double2* yin, double2* yout, double a_next, double w, int first) -> int {
            const double2 *prep = nullptr, *postp = nullptr;
            if (mu) {
                if ((rc = launch_phase_vectors(n, mu, times_host[entry], pre, post, st)) != QDB_OK) return rc;
                prep = pre;
                postp = post;
            }
            bool started = false;
            if (stat_rm) {
                if ((rc = launch_zgemm(n, B, n, D2(stat_rm), n, yin, ldy, kbuf, ldy, one, zero, nullptr, prep, postp, st)) != QDB_OK) return rc;
                started = true;
            }
            for (int j = 0; j < K; ++j) {
                rc = launch_zgemm(n, B, n, D2(ops_rm) + (size_t)j * nn, n, yin, ldy, kbuf, ldy, one, started ? one : zero,
                                  coeff + ((size_t)entry * K + j) * ldc, prep, postp, st);
                if (rc != QDB_OK) return rc;
                started = true;
            }
            return launch_rk4_combine(cnt, Y, kbuf, yout, acc, a_next, w, first, st);
        };
        for (int s = 0; s < S; ++s) {
            if ((rc = stage(2 * s, Y, ya, 0.5 * h, 1.0, 1)) != QDB_OK) return rc;
            if ((rc = stage(2 * s + 1, ya, yb2, 0.5 * h, 2.0, 0)) != QDB_OK) return rc;
            if ((rc = stage(2 * s + 1, yb2, ya, h, 2.0, 0)) != QDB_OK) return rc;
            if ((rc = stage(2 * s + 2, ya, yb2, 0.0, 1.0, 0)) != QDB_OK) return rc;
            if ((rc = launch_axpby(cnt, Y, acc, (1.0 / 6) * h, Y, 1.0, st)) != QDB_OK) return rc;
        }
        return QDB_OK;
    }
    const size_t n2 = align_up((size_t)n * n * sizeof(double2));
    const size_t yb = align_up((size_t)n * B * sizeof(double2));
    double2* G3 = (double2*)ws;  // three generators, contiguous stride n*n (unaligned stride is fine)
    double2* ya = (double2*)(ws + 3 * n2);
    double2* yb_ = (double2*)(ws + 3 * n2 + yb);
    double2* acc = (double2*)(ws + 3 * n2 + 2 * yb);
    double* times_dev = (double*)(ws + 3 * n2 + 3 * yb);
    const size_t nn = (size_t)n * n;
    for (int s = 0; s < S; ++s) {
        if (mu) QDB_CUDA(cudaMemcpyAsync(times_dev, times_host + 2 * s, 3 * sizeof(double), cudaMemcpyHostToDevice, st));
        rc = launch_generator(n, K, 3, QDB_LAYOUT_ROWMAJOR, D2(ops_rm), D2(stat_rm), coeff ? coeff + (size_t)2 * s * K : nullptr,
                              0, mu, times_dev, 0.0, 1.0, G3, st);
        if (rc != QDB_OK) return rc;
        double2* Y = D2(y);
        // k1 = G0 y           ya = y + h/2 k1 ; acc  = k1
        if ((rc = launch_zgemm_rk4stage(n, B, G3, Y, ldy, Y, ya, acc, 0.5 * h, 1.0, 1, st)) != QDB_OK) return rc;
        // k2 = G1 ya          yb = y + h/2 k2 ; acc += 2 k2
        if ((rc = launch_zgemm_rk4stage(n, B, G3 + nn, ya, ldy, Y, yb_, acc, 0.5 * h, 2.0, 0, st)) != QDB_OK) return rc;
        // k3 = G1 yb          ya = y + h k3   ; acc += 2 k3
        if ((rc = launch_zgemm_rk4stage(n, B, G3 + nn, yb_, ldy, Y, ya, acc, h, 2.0, 0, st)) != QDB_OK) return rc;
        // k4 = G2 ya          yb = y (unused) ; acc += k4
        if ((rc = launch_zgemm_rk4stage(n, B, G3 + 2 * nn, ya, ldy, Y, yb_, acc, 0.0, 1.0, 0, st)) != QDB_OK) return rc;
        // y += (1/6) h acc
        if ((rc = launch_axpby((size_t)n * B, Y, acc, (1.0 / 6) * h, Y, 1.0, st)) != QDB_OK) return rc;
    }
    return QDB_OK;
}

int qdb_rk4_table_steps_c128(int n, int B, int S, const qdb_c128* gen_table_packed, int table_layout, double h, qdb_c128* y,
                             int ldy, void* stream) {
    QDB_REQUIRE(n >= 1 && B >= 0 && S >= 0, "qdb_rk4_table_steps_c128: bad n=%d B=%d S=%d", n, B, S);
    QDB_REQUIRE(table_layout == QDB_LAYOUT_PACKED || table_layout == QDB_LAYOUT_PACKED3M,
                "qdb_rk4_table_steps_c128: bad table layout %d", table_layout);
    if (B == 0 || S == 0) return QDB_OK;
    QDB_REQUIRE(gen_table_packed && y && ldy >= B, "qdb_rk4_table_steps_c128: null pointer / bad ldy");
    if (!rk4_fused_supported(n)) {
        set_error("qdb_rk4_table_steps_c128: on-chip path needs n <= 256 (got %d)", n);
        return QDB_E_UNSUPPORTED;
    }
    return launch_rk4_fused_shared(n, B, S, D2(gen_table_packed), table_layout, h, D2(y), ldy, (cudaStream_t)stream);
}

void qdb_ozaki_debug(long long* host64) { qdb::rk4_ozaki_debug(host64); }

size_t qdb_rk4_ozaki_workspace_bytes(int S) { return rk4_ozaki_table_bytes(2 * (S < 1 ? 1 : S) + 1); }
int qdb_rk4_int8_preferred(int n, int B) { return n >= 1 && B >= 1 && rk4_ozaki_preferred(n, B) ? 1 : 0; }

int qdb_rk4_ozaki_slice_c128(int n, int T, const qdb_c128* gen_table, int table_layout, void* workspace, size_t ws_bytes,
                             void* stream) {
    QDB_REQUIRE(n >= 1 && T >= 0, "qdb_rk4_ozaki_slice_c128: bad n=%d T=%d", n, T);
    QDB_REQUIRE(table_layout == QDB_LAYOUT_ROWMAJOR || table_layout == QDB_LAYOUT_PACKED, "qdb_rk4_ozaki_slice_c128: bad table layout %d",
                table_layout);
    if (T == 0) return QDB_OK;
    QDB_REQUIRE(gen_table && workspace, "qdb_rk4_ozaki_slice_c128: null pointer");
    if (!rk4_ozaki_supported(n)) {
        set_error("qdb_rk4_ozaki_slice_c128: the int8 tensor-core emulation exists for n = 65..128 (got %d)", n);
        return QDB_E_UNSUPPORTED;
    }
    if (ws_bytes < rk4_ozaki_table_bytes(T)) {
        set_error("qdb_rk4_ozaki_slice_c128: workspace too small (%zu < %zu)", ws_bytes, rk4_ozaki_table_bytes(T));
        return QDB_E_WORKSPACE;
    }
    return launch_ozaki_slice(n, T, D2(gen_table), table_layout, workspace, (cudaStream_t)stream);
}

int qdb_rk4_ozaki_steps_c128(int n, int B, int S, const qdb_c128* gen_table_rowmajor, double h, qdb_c128* y, int ldy, void* workspace,
                             size_t ws_bytes, void* stream) {
    QDB_REQUIRE(n >= 1 && B >= 0 && S >= 0, "qdb_rk4_ozaki_steps_c128: bad n=%d B=%d S=%d", n, B, S);
    if (B == 0 || S == 0) return QDB_OK;
    QDB_REQUIRE(y && ldy >= B && workspace, "qdb_rk4_ozaki_steps_c128: null pointer / bad ldy");
    if (!rk4_ozaki_supported(n)) {
        set_error("qdb_rk4_ozaki_steps_c128: the int8 tensor-core emulation exists for n = 65..128 (got %d)", n);
        return QDB_E_UNSUPPORTED;
    }
    if (ws_bytes < rk4_ozaki_table_bytes(2 * S + 1)) {
        set_error("qdb_rk4_ozaki_steps_c128: workspace too small (%zu < %zu)", ws_bytes, rk4_ozaki_table_bytes(2 * S + 1));
        return QDB_E_WORKSPACE;
    }
    return launch_rk4_ozaki(n, B, S, D2(gen_table_rowmajor), QDB_LAYOUT_ROWMAJOR, h, D2(y), ldy, workspace, (cudaStream_t)stream);
}

int qdb_signal_table_f64(int T, int K, int B, int nterms, const int* chan, const long long* samp_off, const int* samp_len,
                         const double* dt, const double* t0, const double* freq, const double* phase, int params_per_col,
                         const qdb_c128* samples, long long samp_col_stride, const qdb_c128* scale, const double* times,
                         double t_scalar, double* out, void* stream) {
    QDB_REQUIRE(T >= 0 && K >= 0 && B >= 0 && nterms >= 0, "qdb_signal_table_f64: bad T=%d K=%d B=%d nterms=%d", T, K, B, nterms);
    if (T == 0 || K == 0) return QDB_OK;
    QDB_REQUIRE(out, "qdb_signal_table_f64: null out");
    QDB_REQUIRE(times || T == 1, "qdb_signal_table_f64: times == NULL needs T == 1 (the time is t_scalar)");
    QDB_REQUIRE(nterms == 0 || (chan && samp_off && samp_len && dt && t0 && freq && phase && samples),
                "qdb_signal_table_f64: null term array");
    QDB_REQUIRE(samp_col_stride >= 0, "qdb_signal_table_f64: negative column stride");
    QDB_REQUIRE(B > 0 || (!params_per_col && samp_col_stride == 0), "qdb_signal_table_f64: per-column inputs need B > 0");
    return launch_signal_table(T, K, B, nterms, chan, samp_off, samp_len, dt, t0, freq, phase, params_per_col, D2(samples), samp_col_stride,
                               D2(scale), times, t_scalar, out, (cudaStream_t)stream);
}

int qdb_outcome_probabilities_f64(int n, int B, int n_out, const qdb_c128* y, int ldy, const int* outcome_of, int normalize,
                                  double* out, void* stream) {
    QDB_REQUIRE(n >= 1 && B >= 0 && n_out >= 1, "qdb_outcome_probabilities_f64: bad n=%d B=%d n_out=%d", n, B, n_out);
    if (B == 0) return QDB_OK;
    QDB_REQUIRE(y && outcome_of && out && ldy >= B, "qdb_outcome_probabilities_f64: null pointer / bad ldy");
    return launch_outcome_probabilities(n, B, n_out, D2(y), ldy, outcome_of, normalize, out, (cudaStream_t)stream);
}

int qdb_lindblad_supported(int n) { return lindblad_fused_supported(n) ? 1 : 0; }

int qdb_lindblad_rhs_c128(int n, int J, int B, const qdb_c128* m1_packed, const qdb_c128* m2t_packed, const qdb_c128* diss_packed,
                          const double* gamma, const double* mu, double t, const qdb_c128* rho_in, qdb_c128* rho_out, void* stream) {
    QDB_REQUIRE(n >= 1 && J >= 0 && B >= 0, "qdb_lindblad_rhs_c128: bad n=%d J=%d B=%d", n, J, B);
    if (B == 0) return QDB_OK;
    QDB_REQUIRE(m1_packed && m2t_packed && rho_in && rho_out, "qdb_lindblad_rhs_c128: null pointer");
    QDB_REQUIRE(J == 0 || diss_packed, "qdb_lindblad_rhs_c128: J=%d but no dissipators", J);
    QDB_REQUIRE(rho_in != rho_out, "qdb_lindblad_rhs_c128: rho_in and rho_out must not alias");
    if (!lindblad_fused_supported(n)) {
        set_error("qdb_lindblad_rhs_c128: on-chip path needs n <= 32 (got %d)", n);
        return QDB_E_UNSUPPORTED;
    }
    return launch_lindblad_rhs(n, J, B, D2(m1_packed), D2(m2t_packed), D2(diss_packed), gamma, mu, t, D2(rho_in), D2(rho_out),
                               (cudaStream_t)stream);
}

int qdb_lindblad_rk4_steps_c128(int n, int J, int B, int S, const qdb_c128* m1_table, const qdb_c128* m2t_table,
                                const qdb_c128* diss_packed, const double* gamma_table, const double* mu, const double* times_dev,
                                double h, qdb_c128* rho, void* stream) {
    QDB_REQUIRE(n >= 1 && J >= 0 && B >= 0 && S >= 0, "qdb_lindblad_rk4_steps_c128: bad n=%d J=%d B=%d S=%d", n, J, B, S);
    if (B == 0 || S == 0) return QDB_OK;
    QDB_REQUIRE(m1_table && m2t_table && rho, "qdb_lindblad_rk4_steps_c128: null pointer");
    QDB_REQUIRE(J == 0 || diss_packed, "qdb_lindblad_rk4_steps_c128: J=%d but no dissipators", J);
    QDB_REQUIRE(!mu || times_dev, "qdb_lindblad_rk4_steps_c128: frame given without stage times");
    if (!lindblad_fused_supported(n)) {
        set_error("qdb_lindblad_rk4_steps_c128: on-chip path needs n <= 32 (got %d)", n);
        return QDB_E_UNSUPPORTED;
    }
    return launch_lindblad_rk4(n, J, B, S, D2(m1_table), D2(m2t_table), D2(diss_packed), gamma_table, mu, times_dev, h, D2(rho),
                               (cudaStream_t)stream);
}

int qdb_rk4_tiling(int n, int B, int sweep_K, int* out) {
    QDB_REQUIRE(out && n >= 1 && B >= 1, "qdb_rk4_tiling: bad arguments");
    if (!rk4_fused_tiling(n, B, sweep_K, out)) {
        set_error("qdb_rk4_tiling: on-chip path needs n <= 256 (got %d)", n);
        return QDB_E_UNSUPPORTED;
    }
    return QDB_OK;
}

int qdb_dmma_probe(double* sink, int iters, double* flops_out, void* stream) {
    QDB_REQUIRE(sink && iters > 0, "qdb_dmma_probe: bad arguments");
    int grid = 0;
    const int rc = launch_dmma_probe(sink, iters, &grid, (cudaStream_t)stream);
    if (rc == QDB_OK && flops_out) *flops_out = (double)grid * 8.0 /*warps*/ * iters * 16.0 * 512.0;
    return rc;
}

int qdb_expm_c128(int n, const qdb_c128* A, int squarings, qdb_c128* out, void* workspace, size_t ws_bytes, void* stream) {
    QDB_REQUIRE(n >= 1 && squarings >= 0 && squarings < 64, "qdb_expm_c128: bad n=%d squarings=%d", n, squarings);
    QDB_REQUIRE(A && out && workspace, "qdb_expm_c128: null pointer");
    const size_t e = (size_t)n * n;
    if (ws_bytes < 6 * e * sizeof(double2)) {
        set_error("qdb_expm_c128: workspace too small (%zu < %zu)", ws_bytes, 6 * e * sizeof(double2));
        return QDB_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double2* As = (double2*)workspace;
    int rc = launch_axpby(e, As, D2(A), ldexp(1.0, -squarings), nullptr, 0.0, st);
    if (rc != QDB_OK) return rc;
    return expm_core(n, As, squarings, D2(out), As + e, st);
}

int qdb_expm_steps_c128(int n, int K, int B, int S, const qdb_c128* ops_rm, const qdb_c128* stat_rm, const double* coeff,
                        const double* mu, const double* times_mid_host, const int* squarings_host, double h, qdb_c128* y,
                        int ldy, void* workspace, size_t ws_bytes, void* stream) {
    QDB_REQUIRE(n >= 1 && K >= 0 && B >= 0 && S >= 0, "qdb_expm_steps_c128: bad n=%d K=%d B=%d S=%d", n, K, B, S);
    if (S == 0) return QDB_OK;
    QDB_REQUIRE(stat_rm || (ops_rm && K > 0), "qdb_expm_steps_c128: neither static operator nor operators given");
    QDB_REQUIRE(K == 0 || coeff, "qdb_expm_steps_c128: K=%d but no signal table", K);
    QDB_REQUIRE(squarings_host, "qdb_expm_steps_c128: squarings missing");
    QDB_REQUIRE(!mu || times_mid_host, "qdb_expm_steps_c128: frame given without times");
    QDB_REQUIRE(B == 0 || (y && ldy == B), "qdb_expm_steps_c128: need y with ldy == B");
    if (ws_bytes < qdb_workspace_bytes(QDB_WS_EXPM, n, K, B, 1) || !workspace) {
        set_error("qdb_expm_steps_c128: workspace too small (%zu < %zu)", ws_bytes, qdb_workspace_bytes(QDB_WS_EXPM, n, K, B, 1));
        return QDB_E_WORKSPACE;
    }
    if (B == 0) return QDB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const size_t n2 = align_up((size_t)n * n * sizeof(double2));
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
    int rc;
    // The propagators do not depend on the state: with room for several steps they are built side by side -- one
    // generator launch and one batched Taylor exponential per chunk (a lone n = 128 product occupies four CTAs) -- and
    // then applied one after the other.  All steps of a chunk share the largest number of squarings any of them needs.
    {
        const size_t fixed = align_up((size_t)n * B * sizeof(double2)) + align_up((size_t)S * sizeof(double));
        const size_t room = ws_bytes > fixed ? (ws_bytes - fixed) / (7 * n2) : 0;
        const int Sc = room > (size_t)S ? S : (int)room;
        if (Sc >= 2) {
            const size_t nn = (size_t)n * n;
            double2* As = (double2*)ws;                       // [Sc]
            double2* bws = As + (size_t)Sc * nn;              // [5 Sc]
            double2* P = bws + (size_t)5 * Sc * nn;           // [Sc]
            double2* ytmp = (double2*)(ws + (size_t)Sc * 7 * n2);
            double* times_dev = (double*)((char*)ytmp + align_up((size_t)n * B * sizeof(double2)));
            double2 *ycur = D2(y), *ynext = ytmp;
            for (int s0 = 0; s0 < S; s0 += Sc) {
                const int Sn = S - s0 < Sc ? S - s0 : Sc;
                int sq = 0;
                for (int s = 0; s < Sn; ++s) {
                    const int v = squarings_host[s0 + s];
                    QDB_REQUIRE(v >= 0 && v < 64, "qdb_expm_steps_c128: bad squarings[%d]=%d", s0 + s, v);
                    sq = v > sq ? v : sq;
                }
                if (mu) QDB_CUDA(cudaMemcpyAsync(times_dev, times_mid_host + s0, (size_t)Sn * sizeof(double), cudaMemcpyHostToDevice, st));
                rc = launch_generator(n, K, Sn, QDB_LAYOUT_ROWMAJOR, D2(ops_rm), D2(stat_rm), coeff ? coeff + (size_t)s0 * K : nullptr, 0,
                                      mu, mu ? times_dev : nullptr, 0.0, ldexp(h, -sq), As, st);
                if (rc != QDB_OK) return rc;
                if ((rc = expm_core_batched(n, Sn, As, sq, P, bws, st)) != QDB_OK) return rc;
                for (int s = 0; s < Sn; ++s) {
                    rc = launch_zgemm(n, B, n, P + (size_t)s * nn, n, ycur, ldy, ynext, ldy, one, zero, nullptr, nullptr, nullptr, st);
                    if (rc != QDB_OK) return rc;
                    double2* t = ycur;
                    ycur = ynext;
                    ynext = t;
                }
            }
            if (ycur != D2(y)) QDB_CUDA(cudaMemcpyAsync(y, ycur, (size_t)n * B * sizeof(double2), cudaMemcpyDeviceToDevice, st));
            return QDB_OK;
        }
    }
    double2* As = (double2*)ws;
    double2* P = (double2*)(ws + n2);
    double2* core_ws = (double2*)(ws + 2 * n2);  // 5 n^2 (contiguous, unaligned stride)
    double2* ytmp = (double2*)(ws + 7 * n2);
    double2* ycur = D2(y);
    double2* ynext = ytmp;
    for (int s = 0; s < S; ++s) {
        const int sq = squarings_host[s];
        QDB_REQUIRE(sq >= 0 && sq < 64, "qdb_expm_steps_c128: bad squarings[%d]=%d", s, sq);
        // As = (h / 2^sq) * G_frame(t_s + h/2)
        rc = launch_generator(n, K, 1, QDB_LAYOUT_ROWMAJOR, D2(ops_rm), D2(stat_rm), coeff ? coeff + (size_t)s * K : nullptr, 0,
                              mu, nullptr, mu ? times_mid_host[s] : 0.0, ldexp(h, -sq), As, st);
        if (rc != QDB_OK) return rc;
        if ((rc = expm_core(n, As, sq, P, core_ws, st)) != QDB_OK) return rc;
        if ((rc = launch_zgemm(n, B, n, P, n, ycur, ldy, ynext, ldy, one, zero, nullptr, nullptr, nullptr, st)) != QDB_OK) return rc;
        double2* t = ycur;
        ycur = ynext;
        ynext = t;
    }
    if (ycur != D2(y)) QDB_CUDA(cudaMemcpyAsync(y, ycur, (size_t)n * B * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    return QDB_OK;
}

int qdb_magnus_terms_c128(int n, int magnus_order, const qdb_c128* g, double h, double scale, qdb_c128* out, void* workspace,
                          size_t ws_bytes, void* stream) {
    QDB_REQUIRE(n >= 1, "qdb_magnus_terms_c128: bad n=%d", n);
    QDB_REQUIRE(magnus_order >= 1 && magnus_order <= 3, "qdb_magnus_terms_c128: Only magnus_order 1, 2, and 3 are supported (got %d)",
                magnus_order);
    QDB_REQUIRE(g && out, "qdb_magnus_terms_c128: null pointer");
    const size_t need = (size_t)(magnus_order == 1 ? 0 : magnus_order == 2 ? 1 : 7) * n * n * sizeof(double2);
    if (need > 0 && (!workspace || ws_bytes < need)) {
        set_error("qdb_magnus_terms_c128: workspace too small (%zu < %zu)", ws_bytes, need);
        return QDB_E_WORKSPACE;
    }
    return magnus_terms(n, magnus_order, D2(g), h, scale, D2(out), (double2*)workspace, (cudaStream_t)stream);
}

int qdb_magnus_steps_c128(int n, int K, int B, int S, int magnus_order, const qdb_c128* ops_rm, const qdb_c128* stat_rm,
                          const double* coeff, const double* mu, const double* times_host, const int* squarings_host, double h,
                          qdb_c128* y, int ldy, void* workspace, size_t ws_bytes, void* stream) {
    QDB_REQUIRE(n >= 1 && K >= 0 && B >= 0 && S >= 0, "qdb_magnus_steps_c128: bad n=%d K=%d B=%d S=%d", n, K, B, S);
    QDB_REQUIRE(magnus_order >= 1 && magnus_order <= 3, "qdb_magnus_steps_c128: Only magnus_order 1, 2, and 3 are supported (got %d)",
                magnus_order);
    if (S == 0) return QDB_OK;
    QDB_REQUIRE(stat_rm || (ops_rm && K > 0), "qdb_magnus_steps_c128: neither static operator nor operators given");
    QDB_REQUIRE(K == 0 || coeff, "qdb_magnus_steps_c128: K=%d but no signal table", K);
    QDB_REQUIRE(squarings_host, "qdb_magnus_steps_c128: squarings missing");
    QDB_REQUIRE(!mu || times_host, "qdb_magnus_steps_c128: frame given without times");
    QDB_REQUIRE(B == 0 || (y && ldy == B), "qdb_magnus_steps_c128: need y with ldy == B");
    // smallest usable workspace: one step at a time (qdb_workspace_bytes gives room for a chunk of steps side by side)
    const size_t min_need = 17 * align_up((size_t)n * n * sizeof(double2)) + align_up((size_t)n * B * sizeof(double2)) +
                            align_up((size_t)3 * S * sizeof(double));
    if (ws_bytes < min_need || !workspace) {
        set_error("qdb_magnus_steps_c128: workspace too small (%zu < %zu)", ws_bytes, min_need);
        return QDB_E_WORKSPACE;
    }
    if (B == 0) return QDB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const size_t n2 = align_up((size_t)n * n * sizeof(double2));
    const size_t nn = (size_t)n * n;
    double2* As = (double2*)ws;
    double2* P = (double2*)(ws + n2);
    double2* core_ws = (double2*)(ws + 2 * n2);   // 5 n^2 (contiguous, unaligned stride)
    double2* gnodes = (double2*)(ws + 7 * n2);    // 3 n^2: generator at the nodes of the step
    double2* mag_ws = (double2*)(ws + 10 * n2);   // 7 n^2: commutator temporaries
    double2* ytmp = (double2*)(ws + 17 * n2);
    double* times_dev = (double*)(ws + 17 * n2 + align_up((size_t)n * B * sizeof(double2)));
    const int Q = magnus_order;
    const double2 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
    int rc;
    // With room for several steps, their exponents (node generators, commutators) and exponentials are built side by
    // side in batched launches and then applied in order (see qdb_expm_steps_c128).
    if (Q >= 2) {
        const size_t fixed = align_up((size_t)n * B * sizeof(double2)) + align_up((size_t)3 * S * sizeof(double));
        const size_t room = ws_bytes > fixed ? (ws_bytes - fixed) / (17 * n2) : 0;
        const int Sc = room > (size_t)S ? S : (int)room;
        if (Sc >= 2) {
            double2* gn = (double2*)ws;                                   // [Sc][Q]
            double2* mws = gn + (size_t)Sc * Q * nn;                      // [Sc] or [7 Sc]
            double2* Asb = mws + (size_t)Sc * (Q == 2 ? 1 : 7) * nn;      // [Sc]
            double2* bws = Asb + (size_t)Sc * nn;                         // [5 Sc]
            double2* Pb = bws + (size_t)5 * Sc * nn;                      // [Sc]
            double2* yt = (double2*)(ws + (size_t)Sc * 17 * n2);
            double* tdev = (double*)((char*)yt + align_up((size_t)n * B * sizeof(double2)));
            double2 *ycur = D2(y), *ynext = yt;
            for (int s0 = 0; s0 < S; s0 += Sc) {
                const int Sn = S - s0 < Sc ? S - s0 : Sc;
                int sq = 0;
                for (int s = 0; s < Sn; ++s) {
                    const int v = squarings_host[s0 + s];
                    QDB_REQUIRE(v >= 0 && v < 64, "qdb_magnus_steps_c128: bad squarings[%d]=%d", s0 + s, v);
                    sq = v > sq ? v : sq;
                }
                if (mu) QDB_CUDA(cudaMemcpyAsync(tdev, times_host + (size_t)s0 * Q, (size_t)Sn * Q * sizeof(double), cudaMemcpyHostToDevice, st));
                rc = launch_generator(n, K, Sn * Q, QDB_LAYOUT_ROWMAJOR, D2(ops_rm), D2(stat_rm), coeff ? coeff + (size_t)s0 * Q * K : nullptr,
                                      0, mu, mu ? tdev : nullptr, 0.0, 1.0, gn, st);
                if (rc != QDB_OK) return rc;
                if ((rc = magnus_terms_batched(n, Q, Sn, gn, h, ldexp(1.0, -sq), Asb, mws, st)) != QDB_OK) return rc;
                if ((rc = expm_core_batched(n, Sn, Asb, sq, Pb, bws, st)) != QDB_OK) return rc;
                for (int s = 0; s < Sn; ++s) {
                    rc = launch_zgemm(n, B, n, Pb + (size_t)s * nn, n, ycur, ldy, ynext, ldy, one, zero, nullptr, nullptr, nullptr, st);
                    if (rc != QDB_OK) return rc;
                    double2* t = ycur;
                    ycur = ynext;
                    ynext = t;
                }
            }
            if (ycur != D2(y)) QDB_CUDA(cudaMemcpyAsync(y, ycur, (size_t)n * B * sizeof(double2), cudaMemcpyDeviceToDevice, st));
            return QDB_OK;
        }
    }
    if (mu) QDB_CUDA(cudaMemcpyAsync(times_dev, times_host, (size_t)S * Q * sizeof(double), cudaMemcpyHostToDevice, st));
    double2* ycur = D2(y);
    double2* ynext = ytmp;
    for (int s = 0; s < S; ++s) {
        const int sq = squarings_host[s];
        QDB_REQUIRE(sq >= 0 && sq < 64, "qdb_magnus_steps_c128: bad squarings[%d]=%d", s, sq);
        const double sc = ldexp(1.0, -sq);
        const double* cs = coeff ? coeff + (size_t)s * Q * K : nullptr;
        if (Q == 1) {  // As = (h / 2^sq) G_frame(t_s + h/2) straight from the generator kernel
            rc = launch_generator(n, K, 1, QDB_LAYOUT_ROWMAJOR, D2(ops_rm), D2(stat_rm), cs, 0, mu, mu ? times_dev + s : nullptr, 0.0,
                                  sc * h, As, st);
            if (rc != QDB_OK) return rc;
        } else {
            rc = launch_generator(n, K, Q, QDB_LAYOUT_ROWMAJOR, D2(ops_rm), D2(stat_rm), cs, 0, mu,
                                  mu ? times_dev + (size_t)s * Q : nullptr, 0.0, 1.0, gnodes, st);
            if (rc != QDB_OK) return rc;
            if ((rc = magnus_terms(n, Q, gnodes, h, sc, As, mag_ws, st)) != QDB_OK) return rc;
        }
        if ((rc = expm_core(n, As, sq, P, core_ws, st)) != QDB_OK) return rc;
        if ((rc = launch_zgemm(n, B, n, P, n, ycur, ldy, ynext, ldy, one, zero, nullptr, nullptr, nullptr, st)) != QDB_OK) return rc;
        double2* t = ycur;
        ycur = ynext;
        ynext = t;
    }
    if (ycur != D2(y)) QDB_CUDA(cudaMemcpyAsync(y, ycur, (size_t)n * B * sizeof(double2), cudaMemcpyDeviceToDevice, st));
    return QDB_OK;
}

int qdb_step_propagators_c128(int n, int K, int S, int kind, const qdb_c128* ops_rm, const qdb_c128* stat_rm, const double* coeff,
                              const double* mu, const double* times_host, const int* squarings_host, double h, qdb_c128* P_total,
                              void* workspace, size_t ws_bytes, void* stream) {
    QDB_REQUIRE(n >= 1 && K >= 0 && S >= 1, "qdb_step_propagators_c128: bad n=%d K=%d S=%d", n, K, S);
    QDB_REQUIRE(kind >= 0 && kind <= 3, "qdb_step_propagators_c128: kind %d not in {0 (RK4), 1, 2, 3 (Magnus order)}", kind);
    QDB_REQUIRE(stat_rm || (ops_rm && K > 0), "qdb_step_propagators_c128: neither static operator nor operators given");
    QDB_REQUIRE(K == 0 || coeff, "qdb_step_propagators_c128: K=%d but no signal table", K);
    QDB_REQUIRE(kind == 0 || squarings_host, "qdb_step_propagators_c128: squarings missing");
    QDB_REQUIRE(!mu || times_host, "qdb_step_propagators_c128: frame given without times");
    QDB_REQUIRE(P_total && workspace, "qdb_step_propagators_c128: null pointer");
    return step_propagator_product(n, K, S, kind, D2(ops_rm), D2(stat_rm), coeff, mu, times_host, squarings_host, h, D2(P_total),
                                   workspace, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
