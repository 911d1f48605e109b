// Integer kernels of the codec path, bit-exact with
// oracle/entropy.py quantize_symbols / build_indexes / lookup_rows.  HBM-bound: 16 B/symbol
// (rounding) and 28 B/symbol (lookup); the 109 KB CDF table stays in L1/L2.
// Every global access of the streaming arrays is a 16-byte vector (four symbols per thread-iteration, grid-stride over
// one resident wave); a scalar tail covers n % 4 and unaligned callers.
#include "kernels.cuh"

namespace cdc {

// q = rint(y - mu) (round-half-even), y_hat = q + mu.  mu is elementwise (mu_mod == 0) or a
// per-channel vector indexed by (i / mu_inner) % mu_mod (factorised-prior medians).
__device__ __forceinline__ void quantize_one(float y, float m, int32_t& q, float& yh) {
    const float r = rintf(__fsub_rn(y, m));
    q = static_cast<int32_t>(r);
    yh = __fadd_rn(r, m);
}

__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ y, const float* __restrict__ mu,
                                                       int32_t* __restrict__ q, float* __restrict__ yhat, long long n,
                                                       long long mu_inner, long long mu_mod, int vec) {
    const long long tid = blockIdx.x * 256LL + threadIdx.x, nthr = gridDim.x * 256LL;
    const long long n4 = vec ? n / 4 : 0;
    for (long long v = tid; v < n4; v += nthr) {
        const float4 yy = reinterpret_cast<const float4*>(y)[v];
        float4 mm;
        if (mu_mod) {
            const long long i = v * 4;
            mm.x = mu[(i / mu_inner) % mu_mod];
            mm.y = mu[((i + 1) / mu_inner) % mu_mod];
            mm.z = mu[((i + 2) / mu_inner) % mu_mod];
            mm.w = mu[((i + 3) / mu_inner) % mu_mod];
        } else {
            mm = reinterpret_cast<const float4*>(mu)[v];
        }
        int4 qq;
        float4 hh;
        quantize_one(yy.x, mm.x, qq.x, hh.x);
        quantize_one(yy.y, mm.y, qq.y, hh.y);
        quantize_one(yy.z, mm.z, qq.z, hh.z);
        quantize_one(yy.w, mm.w, qq.w, hh.w);
        reinterpret_cast<int4*>(q)[v] = qq;
        if (yhat) reinterpret_cast<float4*>(yhat)[v] = hh;
    }
    for (long long i = n4 * 4 + tid; i < n; i += nthr) {  // tail (or everything, for unaligned buffers)
        const float m = mu_mod ? mu[(i / mu_inner) % mu_mod] : mu[i];
        int32_t qi;
        float hi;
        quantize_one(y[i], m, qi, hi);
        q[i] = qi;
        if (yhat) yhat[i] = hi;
    }
}

cudaError_t launch_quantize(const float* y, const float* mu, int32_t* q, float* yhat, int64_t n, int64_t mu_inner,
                            int64_t mu_mod, int num_sms, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const int vec = al(y) && al(q) && (!yhat || al(yhat)) && (mu_mod || al(mu)) ? 1 : 0;
    long long want = (n / 4 + 255) / 256 + 1;
    const long long cap = static_cast<long long>(num_sms) * 8;
    quantize_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, s>>>(y, mu, q, yhat, n, mu_inner, mu_mod, vec);
    return cudaGetLastError();
}

// idx = 63 - #{j in [0,62] : max(sigma, table[0]) <= table[j]}  (or the channel when sigma == null);
// v = q - offset[idx]; outside [0, max_v) -> escape; (lo, hi) = cdf[idx][v], cdf[idx][v+1].
struct CdfSmem {
    float tab[64];
    int32_t start[256], len[256], off[256];
};

__device__ __forceinline__ void cdf_one(const CdfSmem& sm, const int32_t* __restrict__ cdf, int rows, bool has_sigma, float sg_in,
                                        int qv, long long i, long long inner, int32_t& idx_o, int32_t& v_o, int32_t& lo_o,
                                        int32_t& hi_o, int32_t& raw_o) {
    int idx;
    if (has_sigma) {
        const float sg = fmaxf(sg_in, sm.tab[0]);
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < 63; ++j) cnt += (sg <= sm.tab[j]) ? 1 : 0;
        idx = 63 - cnt;
    } else {
        idx = static_cast<int>((i / inner) % rows);
    }
    const int max_v = sm.len[idx] - 2;
    int v = qv - sm.off[idx];
    int raw = 0;
    if (v < 0) {
        raw = -2 * v - 1;
        v = max_v;
    } else if (v >= max_v) {
        raw = 2 * (v - max_v);
        v = max_v;
    }
    const int32_t* row = cdf + sm.start[idx];
    idx_o = idx;
    v_o = v;
    lo_o = __ldg(row + v);
    hi_o = __ldg(row + v + 1);
    raw_o = raw;
}

__global__ void __launch_bounds__(256) cdf_lookup_kernel(const int32_t* __restrict__ q, const float* __restrict__ sigma,
                                                         CdfTables t, long long inner, int32_t* __restrict__ idx_o,
                                                         int32_t* __restrict__ v_o, int32_t* __restrict__ lo_o,
                                                         int32_t* __restrict__ hi_o, int32_t* __restrict__ raw_o,
                                                         long long n, int vec) {
    __shared__ CdfSmem sm;
    if (t.scale_table && threadIdx.x < 64) sm.tab[threadIdx.x] = t.scale_table[threadIdx.x];
    for (int r = threadIdx.x; r < t.rows && r < 256; r += 256) {
        sm.start[r] = t.row_start[r];
        sm.len[r] = t.cdf_length[r];
        sm.off[r] = t.offset[r];
    }
    __syncthreads();
    const long long tid = blockIdx.x * 256LL + threadIdx.x, nthr = gridDim.x * 256LL;
    const long long n4 = vec ? n / 4 : 0;
    const bool hs = sigma != nullptr;
    for (long long w = tid; w < n4; w += nthr) {
        const int4 qq = reinterpret_cast<const int4*>(q)[w];
        float4 ss = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hs) ss = reinterpret_cast<const float4*>(sigma)[w];
        int4 a, b, c, d, e;
        const long long i = w * 4;
        cdf_one(sm, t.cdf, t.rows, hs, ss.x, qq.x, i, inner, a.x, b.x, c.x, d.x, e.x);
        cdf_one(sm, t.cdf, t.rows, hs, ss.y, qq.y, i + 1, inner, a.y, b.y, c.y, d.y, e.y);
        cdf_one(sm, t.cdf, t.rows, hs, ss.z, qq.z, i + 2, inner, a.z, b.z, c.z, d.z, e.z);
        cdf_one(sm, t.cdf, t.rows, hs, ss.w, qq.w, i + 3, inner, a.w, b.w, c.w, d.w, e.w);
        reinterpret_cast<int4*>(idx_o)[w] = a;
        reinterpret_cast<int4*>(v_o)[w] = b;
        reinterpret_cast<int4*>(lo_o)[w] = c;
        reinterpret_cast<int4*>(hi_o)[w] = d;
        reinterpret_cast<int4*>(raw_o)[w] = e;
    }
    for (long long i = n4 * 4 + tid; i < n; i += nthr)
        cdf_one(sm, t.cdf, t.rows, hs, hs ? sigma[i] : 0.f, q[i], i, inner, idx_o[i], v_o[i], lo_o[i], hi_o[i], raw_o[i]);
}

cudaError_t launch_cdf_lookup(const int32_t* q, const float* sigma, CdfTables t, int64_t inner, int32_t* idx,
                              int32_t* v, int32_t* lo, int32_t* hi, int32_t* raw, int64_t n, int num_sms,
                              cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    if (t.rows > 256 || (sigma && !t.scale_table)) return cudaErrorInvalidValue;
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    const int vec = al(q) && (!sigma || al(sigma)) && al(idx) && al(v) && al(lo) && al(hi) && al(raw) ? 1 : 0;
    long long want = (n / 4 + 255) / 256 + 1;
    const long long cap = static_cast<long long>(num_sms) * 8;
    cdf_lookup_kernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, s>>>(q, sigma, t, inner, idx, v, lo, hi, raw,
                                                                                n, vec);
    return cudaGetLastError();
}

}  // namespace cdc
