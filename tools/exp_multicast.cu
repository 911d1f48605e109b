// Does TMA multicast shorten the weight fetch at the start of a resident-weight conv launch?
// Every CTA of a kf conv launch ingests the SAME weight block (74 .. 166 KB) before its first MMA; 148 CTAs x 74 KB is
// 11 MB through the L2 -> SM fabric at once, ~2 us at the ~5.5 TB/s this chip delivers from L2.  Here every CTA fetches the
// same `kb` KB in 8 KB bulk copies, either on its own (cluster size 1) or as a cluster of c CTAs in which CTA r issues the
// blocks b = r (mod c) with .multicast::cluster to all c CTAs.  Reported: cycles from kernel start until a CTA's whole
// block has landed (mean / max over CTAs), with a cold and a warm L2.
#include <cuda.h>
#include <cuda_runtime.h>
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "h"(mask)
                 : "memory");
}

struct Res {
    long long total, sync;
};

__global__ void __launch_bounds__(128, 1) fetch_kernel(Res* out, const uint8_t* w, int nblk, int csz) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 200 * 1024;
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    long long t1 = t0;
    if (csz > 1) {
        cluster_sync_all();  // every CTA's barrier exists before a peer multicasts into it
        t1 = clock64();
    }
    if (threadIdx.x < 32) {
        const uint32_t rank = csz > 1 ? cluster_rank() : 0u;
        if (threadIdx.x == 0) mbar_expect_tx(bar, nblk * 8192u);
        __syncwarp();
        for (int b = threadIdx.x; b < nblk; b += 32) {
            if (csz == 1)
                bulk_g2s(base + b * 8192u, w + b * 8192u, 8192u, bar);
            else if (static_cast<uint32_t>(b % csz) == rank)
                bulk_g2s_mc(base + b * 8192u, w + b * 8192u, 8192u, bar, static_cast<uint16_t>((1u << csz) - 1u));
        }
        __syncwarp();
        mbar_wait(bar, 0);
        if (threadIdx.x == 0) {
            out[blockIdx.x].total = clock64() - t0;
            out[blockIdx.x].sync = t1 - t0;
        }
    }
    __syncthreads();
    if (csz > 1) cluster_sync_all();  // nobody exits while a peer may still write into it
}

int main() {
    const int grid = 144;
    Res* d;
    uint8_t *w, *flush;
    cudaMalloc(&d, sizeof(Res) * grid);
    cudaMalloc(&w, 1 << 20);
    cudaMalloc(&flush, 512u << 20);
    cudaMemset(w, 1, 1 << 20);
    const int smem = 1024 + 200 * 1024 + 64;
    cudaFuncSetAttribute(fetch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(fetch_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int kb : {72, 144}) {
        for (int csz : {1, 2, 4, 8}) {
            for (int warm = 0; warm < 2; ++warm) {
                if (!warm) cudaMemset(flush, warm, 512u << 20);  // evict the weights from L2
                cudaMemset(d, 0, sizeof(Res) * grid);
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(128);
                cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = csz;
                at[0].val.clusterDim.y = 1;
                at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                cudaError_t e = cudaLaunchKernelEx(&cfg, fetch_kernel, d, static_cast<const uint8_t*>(w), kb / 8, csz);
                if (e == cudaSuccess) e = cudaDeviceSynchronize();
                std::vector<Res> h(grid);
                cudaMemcpy(h.data(), d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
                double tot = 0, mx = 0, sy = 0;
                for (auto& r : h) {
                    tot += r.total;
                    sy += r.sync;
                    mx = r.total > mx ? r.total : mx;
                }
                printf("%3d KB per CTA, cluster %d, %s L2: %7.0f cycles mean, %7.0f max until the block is in (cluster sync %5.0f) %s\n", kb, csz,
                       warm ? "warm" : "cold", tot / grid, mx, sy / grid, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
