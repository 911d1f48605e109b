// Integer kernels of the codec path, bit-exact with
// oracle/entropy.py quantize_symbols / build_indexes / lookup_rows.  HBM-bound: 16 B/symbol
// (rounding) and 28 B/symbol (lookup); the 109 KB CDF table stays in L1/L2.
#include "kernels.cuh"

namespace cdc {

// q = rint(y - mu) (round-half-even), y_hat = q + mu.  mu is elementwise (mu_mod == 0) or a
// per-channel vector indexed by (i / mu_inner) % mu_mod (factorised-prior medians).
__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ y, const float* __restrict__ mu,
                                                       int32_t* __restrict__ q, float* __restrict__ yhat, long long n,
                                                       long long mu_inner, long long mu_mod) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const float m = mu_mod ? mu[(i / mu_inner) % mu_mod] : mu[i];
        const float r = rintf(__fsub_rn(y[i], m));
        q[i] = static_cast<int32_t>(r);
        if (yhat) yhat[i] = __fadd_rn(r, m);
    }
}

cudaError_t launch_quantize(const float* y, const float* mu, int32_t* q, float* yhat, int64_t n, int64_t mu_inner,
                            int64_t mu_mod, int num_sms, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    long long want = (n + 255) / 256;
    const long long cap = static_cast<long long>(num_sms) * 8;
    quantize_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, s>>>(y, mu, q, yhat, n, mu_inner, mu_mod);
    return cudaGetLastError();
}

// idx = 63 - #{j in [0,62] : max(sigma, table[0]) <= table[j]}  (or the channel when sigma == null);
// v = q - offset[idx]; outside [0, max_v) -> escape; (lo, hi) = cdf[idx][v], cdf[idx][v+1].
__global__ void __launch_bounds__(256) cdf_lookup_kernel(const int32_t* __restrict__ q, const float* __restrict__ sigma,
                                                         CdfTables t, long long inner, int32_t* __restrict__ idx_o,
                                                         int32_t* __restrict__ v_o, int32_t* __restrict__ lo_o,
                                                         int32_t* __restrict__ hi_o, int32_t* __restrict__ raw_o,
                                                         long long n) {
    __shared__ float tab[64];
    __shared__ int32_t s_start[256], s_len[256], s_off[256];
    if (t.scale_table && threadIdx.x < 64) tab[threadIdx.x] = t.scale_table[threadIdx.x];
    for (int r = threadIdx.x; r < t.rows && r < 256; r += 256) {
        s_start[r] = t.row_start[r];
        s_len[r] = t.cdf_length[r];
        s_off[r] = t.offset[r];
    }
    __syncthreads();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        int idx;
        if (sigma) {
            const float sg = fmaxf(sigma[i], tab[0]);
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < 63; ++j) cnt += (sg <= tab[j]) ? 1 : 0;
            idx = 63 - cnt;
        } else {
            idx = static_cast<int>((i / inner) % t.rows);
        }
        const int max_v = s_len[idx] - 2;
        int v = q[i] - s_off[idx];
        int raw = 0;
        if (v < 0) {
            raw = -2 * v - 1;
            v = max_v;
        } else if (v >= max_v) {
            raw = 2 * (v - max_v);
            v = max_v;
        }
        const int32_t* row = t.cdf + s_start[idx];
        idx_o[i] = idx;
        v_o[i] = v;
        lo_o[i] = __ldg(row + v);
        hi_o[i] = __ldg(row + v + 1);
        raw_o[i] = raw;
    }
}

cudaError_t launch_cdf_lookup(const int32_t* q, const float* sigma, CdfTables t, int64_t inner, int32_t* idx,
                              int32_t* v, int32_t* lo, int32_t* hi, int32_t* raw, int64_t n, int num_sms,
                              cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    if (t.rows > 256 || (sigma && !t.scale_table)) return cudaErrorInvalidValue;
    long long want = (n + 255) / 256;
    const long long cap = static_cast<long long>(num_sms) * 8;
    cdf_lookup_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, s>>>(q, sigma, t, inner, idx, v, lo, hi, raw,
                                                                                n);
    return cudaGetLastError();
}

}  // namespace cdc
