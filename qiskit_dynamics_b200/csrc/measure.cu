// Final-state measurement probabilities.
//
// After the last step the reference takes every final state out of the rotating frame, into the dressed basis,
// normalises it and reduces |amplitude|^2 to memory-slot outcome probabilities
// (backend/dynamics_backend.py:846-866, backend/backend_utils.py:106-147; Statevector.probabilities_dict of the
// absent qiskit package restated in oracle/numpy_oracle.py).  The basis changes are two qdb_zgemm_c128 calls
// (the frame phases ride in the second one's `pre` vector); this kernel is the reduction:
//     out[o][b] = sum over basis states i with outcome_of[i] == o of |y[i][b]|^2   ( / sum_i |y[i][b]|^2 )
// These per-column observables are what the multi-GPU path all-gathers.
//
// Streaming kernel, HBM bound: 16 n B bytes read, 8 n_out B written.  One thread owns one column (coalesced
// over b); outcome bins live in registers when n_out <= 16, else in the output array itself.
#include "qdb_common.cuh"

namespace qdb {

template <int NOUT_REG>
__global__ void __launch_bounds__(128) outcome_prob_kernel(int n, int B, int n_out, const double2* __restrict__ y, int ldy,
                                                            const int* __restrict__ outcome_of, int normalize,
                                                            double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double total = 0.0;
    if (NOUT_REG > 0) {
        double bins[NOUT_REG > 0 ? NOUT_REG : 1];
#pragma unroll
        for (int o = 0; o < NOUT_REG; ++o) bins[o] = 0.0;
        for (int i = 0; i < n; ++i) {
            const double2 v = y[(size_t)i * ldy + b];
            const double p = v.x * v.x + v.y * v.y;
            total += p;
            const int o = outcome_of[i];
#pragma unroll
            for (int k = 0; k < NOUT_REG; ++k) bins[k] += (k == o) ? p : 0.0;
        }
        const double inv = normalize ? 1.0 / total : 1.0;
#pragma unroll
        for (int o = 0; o < NOUT_REG; ++o)
            if (o < n_out) out[(size_t)o * B + b] = bins[o] * inv;
    } else {
        for (int o = 0; o < n_out; ++o) out[(size_t)o * B + b] = 0.0;
        for (int i = 0; i < n; ++i) {
            const double2 v = y[(size_t)i * ldy + b];
            const double p = v.x * v.x + v.y * v.y;
            total += p;
            out[(size_t)outcome_of[i] * B + b] += p;
        }
        if (normalize) {
            const double inv = 1.0 / total;
            for (int o = 0; o < n_out; ++o) out[(size_t)o * B + b] *= inv;
        }
    }
}

int launch_outcome_probabilities(int n, int B, int n_out, const double2* y, int ldy, const int* outcome_of, int normalize,
                                 double* out, cudaStream_t st) {
    const int threads = 128;
    const unsigned grid = (unsigned)((B + threads - 1) / threads);
    if (n_out <= 4)
        outcome_prob_kernel<4><<<grid, threads, 0, st>>>(n, B, n_out, y, ldy, outcome_of, normalize, out);
    else if (n_out <= 16)
        outcome_prob_kernel<16><<<grid, threads, 0, st>>>(n, B, n_out, y, ldy, outcome_of, normalize, out);
    else
        outcome_prob_kernel<0><<<grid, threads, 0, st>>>(n, B, n_out, y, ldy, outcome_of, normalize, out);
    QDB_LAUNCH_CHECK("outcome_prob_kernel");
    return QDB_OK;
}

}  // namespace qdb
