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
__global__ void __launch_bounds__(128, 1) rate_kernel(Res* out, int nmma, int row_off, int commit_every, int vary_desc, int mode, int data_mode) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    // A: 3 rows x 136 x 128 B region (zero data is fine), B: N x 128 B
    const uint32_t sA = base, sB = base + 64 * 1024, bars = sB + 256 * 128;
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(gen + 64 * 1024 + 256 * 128 + 64);
    for (int i = threadIdx.x; i < (64 * 1024 + 256 * 128) / 4; i += 128) {
        uint32_t h = (i + 1) * 2654435761u;  // pseudo-random fp16 pairs with |x| in [0.25, 2): realistic operand toggling
        h ^= h >> 13;
        const uint32_t lo = (h & 0x83FFu) | 0x3400u, hi = ((h >> 16) & 0x83FFu) | 0x3800u;
        reinterpret_cast<uint32_t*>(gen)[i] = data_mode ? (lo | (hi << 16)) : 0u;
    }
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
    if (mode >= 4 && warp == 0) {
        // replica of the strip kernel's issuer loop (conv_strip.cu), one "output row" = 9 taps x CH x 4 MMAs
        constexpr int WB = N * 128;
        const int CH = commit_every ? commit_every : 1;  // reuse the argument as the runtime chunk count
        const int NR = 3;
        const bool resident = vary_desc == 0;
        const uint32_t idesc = make_idesc_f16(128, N);
        const uint64_t desc_hi = make_sw128_desc(0) & 0xFFFFFFFF00000000ull;
        const uint32_t leader = threadIdx.x == 0 ? 1u : 0u;
        const uint32_t slot_stride = CH * 17408u, ring = sA, wbase = sB;
        uint32_t aslot = 0;
        auto next_slot = [&](uint32_t s_) { return s_ + 1 == static_cast<uint32_t>(NR) ? 0u : s_ + 1; };
        const long long t0 = clock64();
        for (int r = 0; r < nmma / 36; ++r) {
            const uint32_t s1 = next_slot(aslot), s2 = next_slot(s1);
            const uint32_t rowaddr[3] = {ring + aslot * slot_stride, ring + s1 * slot_stride, ring + s2 * slot_stride};
            uint32_t acc = 0, wb = wbase;
            if (mode == 4) {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        uint32_t aaddr = rowaddr[kh] + kw * 128;
                        for (int ch = 0; ch < CH; ++ch, aaddr += 17408) {
                            const uint64_t adesc = desc_hi | static_cast<uint64_t>((aaddr >> 4) & 0x3FFFu);
                            const uint64_t bdesc = desc_hi | static_cast<uint64_t>((wb >> 4) & 0x3FFFu);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_pred(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, acc, leader);
                                acc = 1;
                            }
                            if (resident) wb += WB / 8;  // (small stride: stay inside the test buffer)
                        }
                    }
                }
            } else {  // mode 5: same work, issued from an elected-lane region with precomputed low words
                if (leader) {
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const uint32_t alo = ((rowaddr[kh] + kw * 128) >> 4) & 0x3FFFu;
                            const uint32_t blo = ((wb + (kh * 3 + kw) * (WB / 8)) >> 4) & 0x3FFFu;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16_ss(tmem, desc_hi | (alo + 2 * k), desc_hi | (blo + 2 * k), idesc, (kh | kw | k) != 0);
                        }
                    }
                }
                __syncwarp();
            }
            if (leader) umma_commit(bars + 8);
            __syncwarp();
            aslot = s1;
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
    } else if (mode >= 1 && warp == 0) {
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
void run(const char* name, int grid, int nmma, int row_off, int commit_every, int vary, int mode = 0, int data_mode = 0) {
    Res* d;
    cudaMalloc(&d, sizeof(Res) * grid);
    const int smem = 1024 + 64 * 1024 + 256 * 128 + 256;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<N><<<grid, 128, smem>>>(d, nmma, row_off, commit_every, vary, mode, data_mode);
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
    const int n = 36 * 64;
    // mode 0: plain single-thread issue, by N (the kh-fused strip kernel issues N = 192 / 96 MMAs)
    for (int grid : {148}) {
        run<32>("plain issue", grid, n, 0, 0, 1, 0, 1);
        run<64>("plain issue", grid, n, 0, 0, 1, 0, 1);
        run<96>("plain issue", grid, n, 0, 0, 1, 0, 1);
        run<128>("plain issue", grid, n, 0, 0, 1, 0, 1);
        run<192>("plain issue", grid, n, 0, 0, 1, 0, 1);
        run<256>("plain issue", grid, n, 0, 0, 1, 0, 1);
        run<192>("plain issue, commit/12", grid, n, 0, 12, 1, 0, 1);
    }
    for (int grid : {148}) {
        run<64>("micro baseline if(elect) x4", grid, n, 0, 0, 0, 3, 1);
        run<64>("strip issuer replica (pred asm)", grid, n, 0, 0, 0, 4, 1);
        run<64>("strip issuer, elected region", grid, n, 0, 0, 0, 5, 1);
        run<128>("strip issuer replica (pred asm)", grid, n, 0, 0, 0, 4, 1);
        run<128>("strip issuer, elected region", grid, n, 0, 0, 0, 5, 1);
    }
    return 0;
}
