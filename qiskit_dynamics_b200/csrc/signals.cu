// Signal tables on the device (the device form of row a6).
//
// A channel value is  s_j(t) = sum over the channel's terms of  Re[ scale * f(t) * exp(i (2 pi nu t + phi)) ]
// (signals/signals.py:144-155, 574-577, 801-803), where the envelope f is piecewise constant
// (DiscreteSignal, signals/signals.py:257-313) or constant (Signal with a numeric envelope).  One thread
// evaluates all K channels of one (time, column) pair and writes the coefficient table the fused RK4
// kernels consume: [T][K] (shared signals) or [T][K][B] (sweep mode: one column per simulation).
//
// Bin selection is integer work and must agree with the reference bit for bit: the reference indexes with
// NumPy's float floor division  (t - t0) // dt  (signals/signals.py:304-308), which is NOT floor((t-t0)/dt)
// (1.0 // 0.1 == 9.0).  npy_floor_divide below restates NumPy's npy_divmod (numpy/_core/src/npymath/
// npy_math_internal.h.src) operation for operation; fmod is exact in IEEE arithmetic, so host and device agree.
//
// Streaming kernel, HBM-write bound: 8 K bytes written per (time, column).
#include "qdb_common.cuh"

namespace qdb {

__device__ __forceinline__ double npy_floor_divide(double a, double b) {
    if (b == 0.0) return a / b;
    double mod = fmod(a, b);
    double div = (a - mod) / b;
    if (mod != 0.0) {
        if ((b < 0.0) != (mod < 0.0)) div -= 1.0;
    }
    double floordiv;
    if (div != 0.0) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        floordiv = copysign(0.0, a / b);
    }
    return floordiv;
}

struct SignalTerms {
    const int* chan;            // [nterms] channel the term adds into
    const long long* samp_off;  // [nterms] first sample of the term in `samples`
    const int* samp_len;        // [nterms] N >= 0 samples (piecewise constant), or -1: constant envelope samples[off]
    const double* dt;           // [nterms], or [nterms][B] when params_per_col
    const double* t0;
    const double* freq;
    const double* phase;
    int params_per_col;
};

__global__ void __launch_bounds__(256) signal_table_kernel(int T, int K, int B, int nterms, SignalTerms tm,
                                                            const double2* __restrict__ samples, long long col_stride,
                                                            const double2* __restrict__ scale /*[nterms][B] or null*/,
                                                            const double* __restrict__ times, double t_scalar,
                                                            double* __restrict__ out) {
    const int ncol = B > 0 ? B : 1;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int t_idx = blockIdx.y;
    if (b >= ncol) return;
    const double t = times ? times[t_idx] : t_scalar;
    double* o = out + ((size_t)t_idx * K) * ncol + b;
    for (int j = 0; j < K; ++j) o[(size_t)j * ncol] = 0.0;
    for (int i = 0; i < nterms; ++i) {
        const int N = tm.samp_len[i];
        const size_t pi = tm.params_per_col ? (size_t)i * ncol + b : (size_t)i;
        const double2* s = samples + tm.samp_off[i] + (size_t)b * col_stride;
        double2 f;
        if (N < 0) {
            f = s[0];
        } else {
            // idx = clip((t - t0) // dt, -1, N); -1 and N both read the zero pad (signals/signals.py:296-311)
            double q = npy_floor_divide(t - tm.t0[pi], tm.dt[pi]);
            // NumPy casts the float quotient to int64 before clipping; saturate the same way for huge values
            long long idx = q >= 9.2e18 ? (long long)9223372036854775807LL : (q <= -9.2e18 ? (-9223372036854775807LL - 1) : (long long)q);
            if (q != q) idx = (-9223372036854775807LL - 1);  // NaN -> INT64_MIN like the x86 cast
            f = (idx < 0 || idx >= N) ? make_double2(0.0, 0.0) : s[idx];
        }
        if (scale != nullptr) f = cmul(scale[(size_t)i * ncol + b], f);
        // exp(t * (2 pi i nu) + i phi): argument assembled like Signal.complex_value (signals/signals.py:148-150),
        // every product and sum rounded separately as NumPy does -- a fused multiply-add here would move theta by
        // half an ulp, i.e. the value by 1e-14 at theta ~ 100
        const double theta = __dadd_rn(__dmul_rn(t, __dmul_rn(6.283185307179586, tm.freq[pi])), tm.phase[pi]);
        double sn, cs;
        sincos(theta, &sn, &cs);
        o[(size_t)tm.chan[i] * ncol] += __dsub_rn(__dmul_rn(f.x, cs), __dmul_rn(f.y, sn));
    }
}

int launch_signal_table(int T, int K, int B, int nterms, const int* chan, const long long* samp_off, const int* samp_len,
                        const double* dt, const double* t0, const double* freq, const double* phase, int params_per_col,
                        const double2* samples, long long col_stride, const double2* scale, const double* times,
                        double t_scalar, double* out, cudaStream_t st) {
    SignalTerms tm{chan, samp_off, samp_len, dt, t0, freq, phase, params_per_col};
    const int ncol = B > 0 ? B : 1;
    const int threads = ncol >= 256 ? 256 : (ncol >= 64 ? 64 : 32);
    for (int ts = 0; ts < T; ts += kMaxGridY) {  // gridDim.y is capped at 65535: slices of the time axis
        const int Tc = T - ts < kMaxGridY ? T - ts : kMaxGridY;
        dim3 grid((unsigned)((ncol + threads - 1) / threads), (unsigned)Tc);
        signal_table_kernel<<<grid, threads, 0, st>>>(Tc, K, B, nterms, tm, samples, col_stride, scale, times ? times + ts : nullptr,
                                                      t_scalar, out + (size_t)ts * K * ncol);
    }
    QDB_LAUNCH_CHECK("signal_table_kernel");
    return QDB_OK;
}

}  // namespace qdb
