// Micro-benchmark / semantics probe for tcgen05.mma.cta_group::2 in the form the kh-fused strip conv would use it:
//   * a CTA pair (cluster 2x1x1); each CTA holds its own 128-row A tile and HALF of the stacked B operand
//     (rows [0, N/2) in CTA 0, rows [N/2, N) in CTA 1, at the same shared-memory offset);
//   * one thread of CTA 0 issues M = 256 MMAs; tcgen05.commit multicasts the completion to both CTAs.
// Part 1 checks the numerics (which B row lands in which D column, which CTA's TMEM holds which rows).
// Part 2 measures cycles per MMA for N = 96 .. 256 against the one-CTA form.
// Part 3 repeats part 2 next to a controlled stream of ld.shared / st.shared from eight other warps (standing in for
// the TMA row loads, the input transform and the epilogue staging of the real kernel): how much of the shared-memory
// port does the pair form leave to the rest of the kernel?
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "ptx.cuh"
using namespace cdc;

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ part 1: numerics
// A: [256][64] fp16 (rows 0..127 in CTA 0), B: [N][64] fp16 (rows 0..N/2-1 in CTA 0); out: [256][N] fp32
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) check_kernel(const __half* A, const __half* B, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t rank = cluster_rank();
    const uint32_t sA = base, sB = base + 16 * 1024, bars = base + 48 * 1024;
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(gen + 48 * 1024 + 64);
    // 128-byte swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
    for (int i = threadIdx.x; i < 128 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(gen + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + (rank * 128 + r) * 64 + c * 8);
    }
    for (int i = threadIdx.x; i < (N / 2) * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(gen + 16 * 1024 + r * 128 + ((c ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(B + (rank * (N / 2) + r) * 64 + c * 8);
    }
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc2(smem_u32(const_cast<uint32_t*>(holder)), 256);
        tmem_relinquish2();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *holder;
    if (rank == 0 && threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_f16(256, N);
        const uint64_t desc_hi = make_sw128_desc(0) & 0xFFFFFFFF00000000ull;
        const uint32_t alo = (sA >> 4) & 0x3FFFu, blo = (sB >> 4) & 0x3FFFu;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma2_f16_ss(tmem, desc_hi | (alo + 2 * k), desc_hi | (blo + 2 * k), idesc, k != 0);
        umma2_commit_mc(bars, 3);
    }
    __syncwarp();
    mbar_wait(bars, 0);
    tc_fence_after();
    const int row = threadIdx.x;  // TMEM lane = warp * 32 + lane
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int c = 0; c < 16; ++c) out[(rank * 128 + row) * N + c0 + c] = __uint_as_float(v[c]);
    }
    if (threadIdx.x == 0) out[256 * N + rank] = __uint_as_float(tmem);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc2(tmem, 256);
}

template <int N>
bool run_check() {
    std::vector<__half> hA(256 * 64), hB(N * 64);
    srand(7 + N);
    for (auto& x : hA) x = __float2half((rand() % 17 - 8) / 8.0f);
    for (auto& x : hB) x = __float2half((rand() % 13 - 6) / 4.0f);
    __half *dA, *dB;
    float* dO;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dO, (256 * N + 2) * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0xFF, (256 * N + 2) * 4);
    const int smem = 1024 + 48 * 1024 + 256;
    cudaFuncSetAttribute(check_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    check_kernel<N><<<2, 128, smem>>>(dA, dB, dO);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> hO(256 * N + 2);
    cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    int bad = 0;
    for (int m = 0; m < 256; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < 64; ++k) ref += __half2float(hA[m * 64 + k]) * __half2float(hB[n * 64 + k]);
            const double d = fabs(ref - hO[m * N + n]);
            if (!(d <= 1e-3)) ++bad;
            if (d > worst || d != d) worst = d;
        }
    unsigned t0, t1;
    memcpy(&t0, &hO[256 * N], 4);
    memcpy(&t1, &hO[256 * N + 1], 4);
    printf("check cta_group::2 M=256 N=%3d: %s  max |err| %.3g, %d bad of %d; tmem base cta0 0x%x cta1 0x%x  %s\n", N, bad == 0 ? "OK" : "MISMATCH",
           worst, bad, 256 * N, t0, t1, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dO);
    return bad == 0 && e == cudaSuccess;
}

// ------------------------------------------------------------------------------------------------ parts 2 and 3: rate
struct Res {
    long long total, issue, bg_iters;
};

// PAIR: cta_group::2 (M = 256 per MMA, half of B per CTA) or the one-CTA form (M = 128, all of B).  Warps 4..11 of every
// CTA stream ld.shared.v4 + st.shared.v4 over a 32 KB scratch region, one 16-byte vector per thread per pass, then idle
// for `gap` cycles (gap < 0: no background traffic).
template <int N, bool PAIR>
__global__ void __launch_bounds__(384, 1) rate_kernel(Res* out, int rows, int gap) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t rank = PAIR ? cluster_rank() : 0u;
    constexpr int BROWS = PAIR ? N / 2 : N;
    const uint32_t sA = base, sB = base + 64 * 1024, scratch = sB + 3 * 256 * 128, bars = scratch + 32 * 1024;
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(gen + 64 * 1024 + 3 * 256 * 128 + 32 * 1024 + 64);
    volatile uint32_t* stop = holder + 1;
    for (int i = threadIdx.x; i < (64 * 1024 + 3 * 256 * 128 + 32 * 1024) / 4; i += 384) {
        uint32_t h = (i + 1) * 2654435761u;
        h ^= h >> 13;
        const uint32_t lo = (h & 0x83FFu) | 0x3400u, hi = ((h >> 16) & 0x83FFu) | 0x3800u;
        reinterpret_cast<uint32_t*>(gen)[i] = lo | (hi << 16);
    }
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);
        *stop = 0;
        fence_mbar_init();
    }
    if (warp == 1) {
        if (PAIR) {
            tmem_alloc2(smem_u32(const_cast<uint32_t*>(holder)), 512);
            tmem_relinquish2();
        } else {
            tmem_alloc(smem_u32(const_cast<uint32_t*>(holder)), 512);
            tmem_relinquish();
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *holder;
    (void)BROWS;
    if (warp == 0) {
        constexpr uint32_t idesc = make_idesc_f16(PAIR ? 256 : 128, N);
        const uint64_t desc_hi = make_sw128_desc(0) & 0xFFFFFFFF00000000ull;
        const bool leader = threadIdx.x == 0 && rank == 0;
        const long long t0 = clock64();
        uint32_t slot = 0;
        for (int r = 0; r < rows; ++r) {
            const uint32_t rowaddr = sA + slot * 17408u;
            const uint32_t dcol = (r & 1) * 256;
            if (leader) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const uint32_t alo = ((rowaddr + kw * 128) >> 4) & 0x3FFFu;
                    const uint32_t blo = ((sB + kw * (256 * 128)) >> 4) & 0x3FFFu;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (PAIR)
                            umma2_f16_ss(tmem + dcol, desc_hi | (alo + 2 * k), desc_hi | (blo + 2 * k), idesc, (kw | k) != 0);
                        else
                            umma_f16_ss(tmem + dcol, desc_hi | (alo + 2 * k), desc_hi | (blo + 2 * k), idesc, (kw | k) != 0);
                    }
                }
            }
            __syncwarp();
            slot = slot == 2 ? 0 : slot + 1;
        }
        const long long t1 = clock64();
        if (leader) {
            if (PAIR)
                umma2_commit_mc(bars + 16, 3);
            else
                umma_commit(bars + 16);
        }
        __syncwarp();
        mbar_wait(bars + 16, 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) {
            *stop = 1;
            out[blockIdx.x].total = t2 - t0;
            out[blockIdx.x].issue = t1 - t0;
        }
    } else if (warp >= 4) {
        long long iters = 0;
        if (gap >= 0) {
            const uint32_t a = scratch + (threadIdx.x - 128) * 16;
            while (*stop == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {  // 8 passes of 4 KB over the 32 KB region
                    uint32_t x, y, z, w;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a + j * 4096));
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + j * 4096), "r"(x ^ 1u), "r"(y), "r"(z), "r"(w) : "memory");
                }
                ++iters;
                const long long c0 = clock64();
                while (clock64() - c0 < gap) {
                }
            }
        }
        if (threadIdx.x == 128) out[blockIdx.x].bg_iters = iters;
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();
    if (warp == 1) {
        if (PAIR)
            tmem_dealloc2(tmem, 512);
        else
            tmem_dealloc(tmem, 512);
    }
}

template <int N, bool PAIR>
void run(int grid, int rows, int gap) {
    Res* d;
    cudaMalloc(&d, sizeof(Res) * grid);
    cudaMemset(d, 0, sizeof(Res) * grid);
    const int smem = 1024 + 64 * 1024 + 3 * 256 * 128 + 32 * 1024 + 256;
    cudaFuncSetAttribute(rate_kernel<N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = PAIR ? 2 : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<N, PAIR>, d, rows, gap);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    std::vector<Res> h(grid);
    cudaMemcpy(h.data(), d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
    double tot = 0, bg = 0;
    int nlead = 0;
    for (int i = 0; i < grid; ++i) {
        if (PAIR && (i & 1)) {
            bg += h[i].bg_iters;
            continue;
        }
        tot += h[i].total;
        bg += h[i].bg_iters;
        ++nlead;
    }
    const double n = 12.0 * rows, cyc = tot / nlead / n;
    // background bytes per cycle per CTA: iterations x (8 warps x 32 lanes x 16 B x 8 passes) x 2 (load + store)
    const double bgB = bg / grid * (256.0 * 16 * 8 * 2) / (tot / nlead);
    // MMA operand bytes per cycle per CTA: A 128 rows x 32 B + B rows x 32 B per MMA
    const double opB = (128 * 32 + (PAIR ? N / 2 : N) * 32) / cyc;
    printf("%s N=%3d gap %5d : %6.1f cyc/MMA (floor %3d = %5.1f %%)  operands %5.1f B/clk + background %5.1f B/clk per CTA %s\n",
           PAIR ? "pair" : "solo", N, gap, cyc, N / 2, 100.0 * (N / 2) / cyc, opB, bgB, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    bool ok = run_check<192>();
    ok = run_check<128>() && ok;
    ok = run_check<256>() && ok;
    ok = run_check<96>() && ok;
    if (!ok) {
        printf("numerics check failed: skipping the rate runs\n");
        return 1;
    }
    const int rows = 256;
    for (int gap : {-1, 1000, 400, 200, 100, 50, 0}) {
        run<192, false>(148, rows, gap);
        run<192, true>(148, rows, gap);
    }
    for (int gap : {-1, 200, 50}) {
        run<128, false>(148, rows, gap);
        run<128, true>(148, rows, gap);
        run<256, false>(148, rows, gap);
        run<256, true>(148, rows, gap);
        run<96, true>(148, rows, gap);
    }
    return 0;
}
