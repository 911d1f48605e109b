// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16/fp16 SS) issued by ONE thread with the
// operands already in shared memory -- by N, by A-descriptor alignment (start row % 8), by commit
// frequency.  Tells whether the conv kernels are bound by the tensor pipe, by shared-memory operand
// fetch, or by the issuing thread.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
#include "ptx.cuh"
using namespace cdc;

struct Res { long long total, issue; };

__device__ __forceinline__ void umma_pred(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc, uint32_t leader) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred;
}

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(Res* out, int nmma, int row_off, int commit_every, int vary_desc, int mode) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    // A: 3 rows x 136 x 128 B region (zero data is fine), B: N x 128 B
    const uint32_t sA = base, sB = base + 64 * 1024, bars = sB + 256 * 128;
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(gen + 64 * 1024 + 256 * 128 + 64);
    for (int i = threadIdx.x; i < (64 * 1024 + 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(holder)), 256);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;
    if (mode == 0 && threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t hi = make_sw128_desc(0) & 0xFFFFFFFF00000000ull;
        uint32_t ph = 0;
        // warm-up
        for (int i = 0; i < 8; ++i) umma_f16_ss(tmem, make_sw128_desc(sA), make_sw128_desc(sB), idesc, i);
        umma_commit(bars);
        mbar_wait(bars, 0);
        const long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            uint32_t a = sA + row_off * 128 + (i & 3) * 32;
            if (vary_desc) a += ((i >> 2) % 3) * 128 + (((i >> 2) / 3) % 3) * 17408;
            const uint64_t ad = hi | ((a >> 4) & 0x3FFF), bd = hi | (((sB + (i & 3) * 32) >> 4) & 0x3FFF);
            umma_f16_ss(tmem, ad, bd, idesc, 1);
            if (commit_every && (i % commit_every) == commit_every - 1) umma_commit(bars + 8);
        }
        const long long t1 = clock64();
        umma_commit(bars + 16);
        mbar_wait(bars + 16, ph);
        const long long t2 = clock64();
        out[blockIdx.x].total = t2 - t0;
        out[blockIdx.x].issue = t1 - t0;
    }
    if (mode >= 1 && warp == 0) {
        // warp-converged issue loop: descriptors are warp-uniform, only the MMA itself is predicated
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t hi = make_sw128_desc(0) & 0xFFFFFFFF00000000ull;
        const uint32_t leader = mode == 1 ? (threadIdx.x == 0 ? 1u : 0u) : elect_one();
        const long long t0 = clock64();
        if (mode == 3) {
            if (leader) {
#pragma unroll 1
                for (int i = 0; i < nmma; i += 4) {
                    const uint32_t a = sA + row_off * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem, hi | (((a + k * 32) >> 4) & 0x3FFF), hi | (((sB + k * 32) >> 4) & 0x3FFF), idesc, 1);
                }
            }
            __syncwarp();
        } else {
#pragma unroll 1
            for (int i = 0; i < nmma; i += 4) {
                const uint32_t a = sA + row_off * 128 + (vary_desc ? ((i >> 2) % 3) * 128 : 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_pred(tmem, hi | (((a + k * 32) >> 4) & 0x3FFF), hi | (((sB + k * 32) >> 4) & 0x3FFF), idesc, 1, leader);
            }
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
    if (warp == 1) tmem_dealloc(tmem, 256);
}

template <int N>
void run(const char* name, int grid, int nmma, int row_off, int commit_every, int vary, int mode = 0) {
    Res* d;
    cudaMalloc(&d, sizeof(Res) * grid);
    const int smem = 1024 + 64 * 1024 + 256 * 128 + 256;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<N><<<grid, 128, smem>>>(d, nmma, row_off, commit_every, vary, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<Res> h(grid);
    cudaMemcpy(h.data(), d, sizeof(Res) * grid, cudaMemcpyDeviceToHost);
    double tot = 0, iss = 0;
    for (auto& r : h) { tot += r.total; iss += r.issue; }
    printf("%-34s mode %d grid %3d N=%3d row_off=%d commit_every=%d vary=%d : %7.1f cyc/MMA total, %7.1f cyc/MMA issue (floor %d) %s\n", name, mode, grid, N,
           row_off, commit_every, vary, tot / grid / nmma, iss / grid / nmma, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    const int n = 2048;
    for (int grid : {148}) {
        run<64>("lane0 branch (baseline)", grid, n, 0, 0, 0, 0);
        run<64>("converged, @pred mma", grid, n, 0, 0, 0, 1);
        run<64>("converged, @pred mma, row+1", grid, n, 1, 0, 0, 1);
        run<64>("converged, elect pred", grid, n, 0, 0, 0, 2);
        run<64>("if(elect) unrolled x4", grid, n, 0, 0, 0, 3);
        run<64>("converged, @pred, vary taps", grid, n, 0, 0, 1, 1);
        run<128>("converged, @pred mma", grid, n, 0, 0, 0, 1);
        run<128>("if(elect) unrolled x4", grid, n, 0, 0, 0, 3);
        run<256>("converged, @pred mma", grid, n, 0, 0, 0, 1);
        run<16>("converged, @pred mma", grid, n, 0, 0, 0, 1);
    }
    return 0;
}
