// Micro-benchmark for the kh-fused strip conv: cycles per tcgen05.mma (M=128, K=16, fp16 SS) by N, issued as
// straight-line code from an elected lane (the form conv_kf.cu uses), with an optional tcgen05.commit every
// CE MMAs.  Operands are resident in shared memory; A descriptors walk the (kw, k) offsets of a 3x3 strip row.
// Answers: (1) does one N=192 MMA (three kh taps stacked along N) cost ~96 cycles, i.e. half of three N=64
// MMAs; (2) what does a commit cost the issue stream.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
#include "ptx.cuh"
using namespace cdc;

struct Res { long long total, issue; };

template <int N, int CE>
__global__ void __launch_bounds__(128, 1) rate_kernel(Res* out, int rows) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t sA = base, sB = base + 64 * 1024, bars = sB + 3 * 256 * 128;
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(gen + 64 * 1024 + 3 * 256 * 128 + 64);
    for (int i = threadIdx.x; i < (64 * 1024 + 3 * 256 * 128) / 4; i += 128) {
        uint32_t h = (i + 1) * 2654435761u;
        h ^= h >> 13;
        const uint32_t lo = (h & 0x83FFu) | 0x3400u, hi = ((h >> 16) & 0x83FFu) | 0x3800u;
        reinterpret_cast<uint32_t*>(gen)[i] = lo | (hi << 16);
    }
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(holder)), 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;
    if (warp == 0) {
        constexpr uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t desc_hi = make_sw128_desc(0) & 0xFFFFFFFF00000000ull;
        const uint32_t leader = threadIdx.x == 0 ? 1u : 0u;
        const long long t0 = clock64();
        uint32_t slot = 0;
        for (int r = 0; r < rows; ++r) {
            const uint32_t rowaddr = sA + slot * 17408u;
            const uint32_t dcol = (r & 1) * 256;  // alternate accumulator windows
            if (leader) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const uint32_t alo = ((rowaddr + kw * 128) >> 4) & 0x3FFFu;
                    const uint32_t blo = ((sB + kw * (256 * 128)) >> 4) & 0x3FFFu;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_f16_ss(tmem + dcol, desc_hi | (alo + 2 * k), desc_hi | (blo + 2 * k), idesc, (kw | k) != 0);
                        if (CE > 0 && ((kw * 4 + k) % CE) == CE - 1) umma_commit(bars + 8);
                    }
                }
            }
            __syncwarp();
            slot = slot == 2 ? 0 : slot + 1;
        }
        const long long t1 = clock64();
        if (leader) umma_commit(bars + 16);
        __syncwarp();
        mbar_wait(bars + 16, 0);
        const long long t2 = clock64();
        if (threadIdx.x == 0) {
            out[blockIdx.x].total = t2 - t0;
            out[blockIdx.x].issue = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N, int CE>
void run(int grid, int rows) {
    Res* d;
    cudaMalloc(&d, sizeof(Res) * grid);
    const int smem = 1024 + 64 * 1024 + 3 * 256 * 128 + 256;
    cudaFuncSetAttribute(rate_kernel<N, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<N, CE><<<grid, 128, smem>>>(d, rows);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<Res> h(grid);
    cudaMemcpy(h.data(), d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
    double tot = 0, iss = 0;
    for (auto& r : h) { tot += r.total; iss += r.issue; }
    const double n = 12.0 * rows;
    printf("N=%3d commit_every=%2d grid %3d : %7.1f cyc/MMA total, %7.1f issue (floor %d) -> %.2f cyc per 64 columns %s\n", N, CE, grid,
           tot / grid / n, iss / grid / n, N / 2, tot / grid / n * 64.0 / N, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    const int rows = 256;
    run<32, 0>(148, rows);
    run<64, 0>(148, rows);
    run<96, 0>(148, rows);
    run<128, 0>(148, rows);
    run<192, 0>(148, rows);
    run<256, 0>(148, rows);
    run<192, 12>(148, rows);
    run<192, 4>(148, rows);
    run<192, 1>(148, rows);
    run<64, 12>(148, rows);
    run<64, 4>(148, rows);
    run<96, 12>(148, rows);
    run<192, 0>(1, rows);
    return 0;
}
