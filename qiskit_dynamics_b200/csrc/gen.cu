// Operator packing + generator-table formation.
//
// generator_kernel is a pure streaming kernel: it reads the K+1 stored operators once per time
// point (L2 resident: 2.4 MB at n=128, K=8) and writes one n x n generator per time point.  It is
// HBM/L2 bound: algorithmic bytes per output matrix = 16 (K + 2) n^2.
#include "qdb_common.cuh"

namespace qdb {

__global__ void pack_kernel(int n, int npad, size_t per_src, size_t per_dst, const double2* __restrict__ src,
                            double2* __restrict__ dst) {
    const int j = blockIdx.y;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per_dst) return;
    const int KT = round_up16(n) >> 2;
    const size_t tile = e >> 5;
    const int lane = (int)(e & 31);
    const int r = (int)(tile / KT) * 8 + (lane >> 2);
    const int c = (int)(tile % KT) * 4 + (lane & 3);
    double2 v = make_double2(0.0, 0.0);
    if (r < n && c < n) v = src[(size_t)j * per_src + (size_t)r * n + c];
    (void)npad;
    dst[(size_t)j * per_dst + e] = v;
}

int launch_pack(int n, int count, const double2* src, double2* dst, cudaStream_t st) {
    const int npad = round_up8(n);
    const size_t per_dst = (size_t)npad * round_up16(n);
    dim3 grid((unsigned)((per_dst + 255) / 256), count);
    pack_kernel<<<grid, 256, 0, st>>>(n, npad, (size_t)n * n, per_dst, src, dst);
    QDB_LAUNCH_CHECK("pack_kernel");
    return QDB_OK;
}

// out[t][e] = scale * (stat[e] + sum_j c[t][j] ops[j][e]) * conj(p_r(t)) * p_c(t)
template <bool kComplexCoeff>
__global__ void __launch_bounds__(256) generator_kernel(int n, int npad, int K, int layout, size_t elems,
                                                         const double2* __restrict__ ops,
                                                         const double2* __restrict__ stat,
                                                         const double* __restrict__ coeff,
                                                         const double* __restrict__ mu,
                                                         const double* __restrict__ times, double t_scalar,
                                                         double scale, double2* __restrict__ out) {
    extern __shared__ double2 s_phase[];  // [n] when mu != nullptr
    __shared__ double s_c[2 * 64];        // coefficients of this time point (K <= 64 per pass)
    const int t_idx = blockIdx.y;
    const bool framed = (mu != nullptr);
    if (framed) {
        const double t = times ? times[t_idx] : t_scalar;
        for (int a = threadIdx.x; a < n; a += blockDim.x) s_phase[a] = frame_phase(mu[a], t);
    }
    const size_t e0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;  // two elements / thread
    double2 acc[2];
    bool live[2];
    int rr[2], cc[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const size_t e = e0 + u;
        live[u] = e < elems;
        acc[u] = make_double2(0.0, 0.0);
        rr[u] = cc[u] = 0;
        if (live[u]) {
            if (layout != QDB_LAYOUT_ROWMAJOR) {
                const int KT = round_up16(n) >> 2;
                const size_t tile = e >> 5;
                const int lane = (int)(e & 31);
                rr[u] = (int)(tile / KT) * 8 + (lane >> 2);
                cc[u] = (int)(tile % KT) * 4 + (lane & 3);
            } else {
                rr[u] = (int)(e / n);
                cc[u] = (int)(e % n);
            }
            if (stat != nullptr) acc[u] = stat[e];
        }
    }
    for (int j0 = 0; j0 < K; j0 += 64) {
        const int kc = min(64, K - j0);
        __syncthreads();
        if (threadIdx.x < kc) {
            if (kComplexCoeff) {
                s_c[2 * threadIdx.x] = coeff[2 * ((size_t)t_idx * K + j0 + threadIdx.x)];
                s_c[2 * threadIdx.x + 1] = coeff[2 * ((size_t)t_idx * K + j0 + threadIdx.x) + 1];
            } else {
                s_c[threadIdx.x] = coeff[(size_t)t_idx * K + j0 + threadIdx.x];
            }
        }
        __syncthreads();
        for (int j = 0; j < kc; ++j) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (!live[u]) continue;
                const double2 g = ops[(size_t)(j0 + j) * elems + e0 + u];
                if (kComplexCoeff) {
                    const double2 c = make_double2(s_c[2 * j], s_c[2 * j + 1]);
                    acc[u] = cadd(acc[u], cmul(c, g));
                } else {
                    const double c = s_c[j];
                    acc[u].x = fma(c, g.x, acc[u].x);
                    acc[u].y = fma(c, g.y, acc[u].y);
                }
            }
        }
    }
    __syncthreads();  // s_phase visible (also covers the K == 0 case)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        if (!live[u]) continue;
        double2 v = acc[u];
        if (framed && rr[u] < n && cc[u] < n) {
            // frame_mat = conj(e_r) * e_c, e = exp(d t) = p   (rotating_frame.py:350-353)
            const double2 ph = cmul_conj_a(s_phase[rr[u]], s_phase[cc[u]]);
            v = cmul(v, ph);
        }
        v.x *= scale;
        v.y *= scale;
        if (layout == QDB_LAYOUT_PACKED3M) {
            // entry = complex plane [elems] followed by the (re + im) plane [elems] the 3M kernel multiplies with
            double2* entry = out + (size_t)t_idx * (elems + elems / 2);
            entry[e0 + u] = v;
            reinterpret_cast<double*>(entry + elems)[e0 + u] = v.x + v.y;
        } else {
            out[(size_t)t_idx * elems + e0 + u] = v;
        }
    }
}

int launch_generator(int n, int K, int T, int layout, const double2* ops, const double2* stat,
                     const double* coeff, int coeff_complex, const double* mu, const double* times,
                     double t_scalar, double scale, double2* out, cudaStream_t st) {
    const int npad = round_up8(n);
    const size_t elems = layout != QDB_LAYOUT_ROWMAJOR ? (size_t)npad * round_up16(n) : (size_t)n * n;
    const size_t smem = mu ? (size_t)n * sizeof(double2) : 0;
    if (smem > 48 * 1024) {
        if (coeff_complex)
            QDB_CUDA(cudaFuncSetAttribute(generator_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            QDB_CUDA(cudaFuncSetAttribute(generator_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    // gridDim.y is capped at 65535: long intervals (T = 2 S + 1 stage times) go up in slices of the time axis
    const size_t out_stride = layout == QDB_LAYOUT_PACKED3M ? elems + elems / 2 : elems;
    for (int ts = 0; ts < T; ts += kMaxGridY) {
        const int Tc = T - ts < kMaxGridY ? T - ts : kMaxGridY;
        dim3 grid((unsigned)((elems + 511) / 512), (unsigned)Tc);
        const double* cs = coeff ? coeff + (size_t)ts * K * (coeff_complex ? 2 : 1) : nullptr;
        const double* tms = times ? times + ts : nullptr;
        double2* os = out + (size_t)ts * out_stride;
        if (coeff_complex)
            generator_kernel<true><<<grid, 256, smem, st>>>(n, npad, K, layout, elems, ops, stat, cs, mu, tms, t_scalar, scale, os);
        else
            generator_kernel<false><<<grid, 256, smem, st>>>(n, npad, K, layout, elems, ops, stat, cs, mu, tms, t_scalar, scale, os);
    }
    QDB_LAUNCH_CHECK("generator_kernel");
    return QDB_OK;
}

// dst = a*x + b*y (complex arrays, real scalars); y may be nullptr
__global__ void axpby_kernel(size_t count, double2* dst, const double2* x, double a, const double2* y, double b) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double2 v = x[i];
    v.x *= a;
    v.y *= a;
    if (y) {
        const double2 w = y[i];
        v.x = fma(b, w.x, v.x);
        v.y = fma(b, w.y, v.y);
    }
    dst[i] = v;
}

// RK4 stage combine after a multi-launch stage product k (generic large-n sweep path):
//   yout = ybase + a_next * k ;  acc = (first ? 0 : acc) + w * k
__global__ void rk4_combine_kernel(size_t count, const double2* __restrict__ ybase, const double2* __restrict__ k,
                                   double2* __restrict__ yout, double2* __restrict__ acc, double a_next, double w, int first) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double2 kv = k[i], yb = ybase[i];
    yout[i] = make_double2(fma(a_next, kv.x, yb.x), fma(a_next, kv.y, yb.y));
    double2 a = first ? make_double2(0.0, 0.0) : acc[i];
    a.x = fma(w, kv.x, a.x);
    a.y = fma(w, kv.y, a.y);
    acc[i] = a;
}

int launch_rk4_combine(size_t count, const double2* ybase, const double2* k, double2* yout, double2* acc, double a_next,
                       double w, int first, cudaStream_t st) {
    if (count == 0) return QDB_OK;
    rk4_combine_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, ybase, k, yout, acc, a_next, w, first);
    QDB_LAUNCH_CHECK("rk4_combine_kernel");
    return QDB_OK;
}

int launch_axpby(size_t count, double2* dst, const double2* x, double a, const double2* y, double b, cudaStream_t st) {
    if (count == 0) return QDB_OK;
    axpby_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(count, dst, x, a, y, b);
    QDB_LAUNCH_CHECK("axpby_kernel");
    return QDB_OK;
}

}  // namespace qdb

namespace qdb {

// pre[a] = exp(-i mu_a t), post[a] = conj(pre[a])
__global__ void phase_kernel(int n, const double* __restrict__ mu, double t, double2* pre, double2* post) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const double2 p = frame_phase(mu[a], t);
    pre[a] = p;
    post[a] = make_double2(p.x, -p.y);
}

int launch_phase_vectors(int n, const double* mu, double t, double2* pre, double2* post, cudaStream_t st) {
    phase_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, mu, t, pre, post);
    QDB_LAUNCH_CHECK("phase_kernel");
    return QDB_OK;
}

// out = c0 I + c1 A1 + c2 A2 + c3 A3 + c4 A4   (n x n row-major; null pointers are skipped)
__global__ void poly_kernel(int n, double c0, double c1, const double2* __restrict__ A1, double c2,
                            const double2* __restrict__ A2, double c3, const double2* __restrict__ A3, double c4,
                            const double2* __restrict__ A4, double2* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * n) return;
    const int r = (int)(e / n), c = (int)(e % n);
    double2 v = make_double2(r == c ? c0 : 0.0, 0.0);
    if (A1) { const double2 a = A1[e]; v.x = fma(c1, a.x, v.x); v.y = fma(c1, a.y, v.y); }
    if (A2) { const double2 a = A2[e]; v.x = fma(c2, a.x, v.x); v.y = fma(c2, a.y, v.y); }
    if (A3) { const double2 a = A3[e]; v.x = fma(c3, a.x, v.x); v.y = fma(c3, a.y, v.y); }
    if (A4) { const double2 a = A4[e]; v.x = fma(c4, a.x, v.x); v.y = fma(c4, a.y, v.y); }
    out[e] = v;
}

int launch_poly(int n, double c0, double c1, const double2* A1, double c2, const double2* A2, double c3,
                const double2* A3, double c4, const double2* A4, double2* out, cudaStream_t st) {
    const size_t e = (size_t)n * n;
    poly_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(n, c0, c1, A1, c2, A2, c3, A3, c4, A4, out);
    QDB_LAUNCH_CHECK("poly_kernel");
    return QDB_OK;
}

}  // namespace qdb

namespace qdb {

// y_out[a][b] = q_a * y_in[a][b],  q = exp(-i mu t) or its conjugate (a3 as a stand-alone op)
__global__ void frame_apply_kernel(int n, int B, const double* __restrict__ mu, double t, int conj_phase,
                                   const double2* __restrict__ y_in, double2* __restrict__ y_out, int ldy) {
    const int a = blockIdx.y;
    double2 p = frame_phase(mu[a], t);
    if (conj_phase) p.y = -p.y;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x)
        y_out[(size_t)a * ldy + b] = cmul(p, y_in[(size_t)a * ldy + b]);
}

int launch_frame_apply(int n, int B, const double* mu, double t, int conj_phase, const double2* y_in, double2* y_out,
                       int ldy, cudaStream_t st) {
    if (n == 0 || B == 0) return QDB_OK;
    dim3 grid((unsigned)min((B + 255) / 256, 1024), n);
    frame_apply_kernel<<<grid, 256, 0, st>>>(n, B, mu, t, conj_phase, y_in, y_out, ldy);
    QDB_LAUNCH_CHECK("frame_apply_kernel");
    return QDB_OK;
}

}  // namespace qdb
