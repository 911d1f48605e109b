// Experiment (round-1 design probe): can a SWIZZLE_128B K-major UMMA A-descriptor start at a row
// that is NOT a multiple of 8 (start address not 1024 B aligned), with or without the descriptor's
// base_offset field?  If yes, a 3x3 conv can load ONE halo tile (3 rows x 130 pixels) per 64-channel
// slice and run all 9 taps out of it by shifting the descriptor start, cutting L2->smem traffic ~3x.
// Build+run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/exp_halo tools/exp_halo.cu -I<csrc> && /tmp/exp_halo
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "ptx.cuh"
using namespace cdc;

constexpr int R = 192, NCFG = 24;
struct Cfg { int s, bo; };
struct Params {
    CUtensorMap amap, bmap;
    Cfg cfg[NCFG];
    int ncfg;
    float* out;  // [ncfg][128][64]
};

__global__ void __launch_bounds__(128, 1) exp_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    const uint32_t sA = base, sB = base + R * 128, bars = sB + 64 * 128;
    volatile uint32_t* holder = reinterpret_cast<volatile uint32_t*>(gen + R * 128 + 64 * 128 + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(holder)), 64);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *holder;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bars, R * 128 + 64 * 128);
        tma_load_2d(sA, &p.amap, bars, 0, 0);
        tma_load_2d(sB, &p.bmap, bars, 0, 0);
    }
    mbar_wait(bars, 0);
    tc_fence_after();
    uint32_t ph = 0;
    for (int c = 0; c < p.ncfg; ++c) {
        if (threadIdx.x == 0) {
            const uint64_t ad = make_sw128_desc(sA + p.cfg[c].s * 128, p.cfg[c].bo);
            const uint64_t bd = make_sw128_desc(sB);
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, ad + 2 * k, bd + 2 * k, ((1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24)), k != 0);
            umma_commit(bars + 8);
        }
        mbar_wait(bars + 8, ph);
        ph ^= 1;
        tc_fence_after();
        uint32_t v[32];
        for (int h = 0; h < 2; ++h) {
            tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + h * 32, v);
            tmem_ld_wait();
            float* o = p.out + (static_cast<size_t>(c) * 128 + warp * 32 + lane) * 64 + h * 32;
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 1) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
    std::vector<__nv_bfloat16> G(R * 64), Bm(64 * 64);
    auto gval = [](int r, int c) { return static_cast<float>((r * 7 + c * 3) % 97 - 48); };
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < 64; ++c) G[r * 64 + c] = __float2bfloat16(gval(r, c));
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) Bm[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dG, *dB;
    cudaMalloc(&dG, G.size() * 2);
    cudaMalloc(&dB, Bm.size() * 2);
    cudaMemcpy(dG, G.data(), G.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bm.data(), Bm.size() * 2, cudaMemcpyHostToDevice);
    Params p;
    cuuint64_t dims[2] = {64, R}, strides[1] = {128};
    cuuint32_t box[2] = {64, R}, es[2] = {1, 1};
    CUresult r1 = enc(&p.amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dG, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dimsb[2] = {64, 64};
    cuuint32_t boxb[2] = {64, 64};
    CUresult r2 = enc(&p.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d %d\n", (int)r1, (int)r2);
    const int ss[] = {0, 1, 2, 3, 5, 7, 8, 9, 16, 17, 33, 64};
    p.ncfg = 0;
    for (int s : ss) {
        p.cfg[p.ncfg++] = {s, 0};
        if (s & 7) p.cfg[p.ncfg++] = {s, s & 7};
    }
    cudaMalloc(&p.out, static_cast<size_t>(p.ncfg) * 128 * 64 * 4);
    const int smem = 1024 + R * 128 + 64 * 128 + 256;
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    exp_kernel<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> out(static_cast<size_t>(p.ncfg) * 128 * 64);
    cudaMemcpy(out.data(), p.out, out.size() * 4, cudaMemcpyDeviceToHost);
    for (int c = 0; c < p.ncfg; ++c) {
        const int s = p.cfg[c].s;
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n)
                if (out[(static_cast<size_t>(c) * 128 + m) * 64 + n] != gval(s + m, n)) ++bad;
        printf("start row %3d base_offset %d : %s (%d / 8192 mismatches)", s, p.cfg[c].bo, bad ? "WRONG" : "exact", bad);
        if (bad) {  // which source row did output rows 0..9 come from (matching the first 8 columns)?
            printf("  rows<-");
            for (int m = 0; m < 10; ++m) {
                int found = -1;
                for (int r = 0; r < R && found < 0; ++r) {
                    bool ok = true;
                    for (int n = 0; n < 8; ++n) ok &= out[(static_cast<size_t>(c) * 128 + m) * 64 + n] == gval(r, n);
                    if (ok) found = r;
                }
                printf(" %d", found);
            }
        }
        printf("\n");
    }
    return 0;
}
