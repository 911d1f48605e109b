// K-conv, strip variant: 3x3 stride-1 convolution on full-resolution levels, built to cut the
// L2 -> shared-memory traffic that bounds conv_tc.cu there (measured ~33 B/cycle/SM against a demand
// of >100 B/cycle for C_out = 64).
//
//   work unit : a vertical strip = 128 output columns x L output rows of one image
//   A operand : a ring of INPUT ROWS in shared memory.  Each input row (130 pixels = 128 + halo, 64
//               channels per chunk) is loaded ONCE by a 4-D TMA box {64, 130, 1, 1} (zero-filled outside
//               the image = conv padding) and serves the 3 output rows above/at/below it and all 3
//               horizontal taps: tap (kh, kw) of output row h is the UMMA descriptor that starts at
//               ring[row h+kh-1] + kw*128 B.  A SWIZZLE_128B K-major descriptor may start at any
//               128 B row (the swizzle is a function of the absolute smem address; tools/exp_halo.cu
//               verified this on B200), so no data is moved or duplicated for the 9 taps.
//   B operand : the whole [9 taps][C_in][C_out] weight block stays resident in shared memory when it
//               fits (C_in = C_out = 64: 72 KB), otherwise it streams through its own small ring.
//   Result    : A traffic per output row drops from 9 x 16 KB to 16.3 KB (x chunks).
//
//   warp 0 : row producer (TMA)      warp 1 : tcgen05.mma issuer     warp 2 : TMEM allocator
//   warp 3 : weight producer (TMA)   warps 4-7 : epilogue (shared with conv_tc.cu)
//
// Oracle counterpart: oracle/unet.py `conv(k=3)` inside RB / stem / final (the reference ships no code).
#include <stdio.h>

#include "conv_epilogue.cuh"
#include "conv_strip.cuh"
#include "ptx.cuh"

namespace cdc {

constexpr int kRowBytes = 17 * 1024;       // 130 pixels x 128 B = 16640, padded to a 1024 B multiple
constexpr int kRowTx = 130 * 128;          // bytes one row box delivers
constexpr int kStripBar = 512;             // barrier block
constexpr int kStripAux = kStripBar + 768 * 4 + 2 * 4 * 32 * 2 * 4;

template <int BN, int CPG, int EPI>
__global__ void __launch_bounds__(256, 1) conv_strip_kernel(const __grid_constant__ StripParams p) {
    constexpr int WB = BN * 128;  // one (tap, chunk) weight slice
    constexpr int ACC_STRIDE = BN < 32 ? 32 : BN;
    constexpr int TMEM_COLS = (2 * ACC_STRIDE <= 32) ? 32 : (2 * ACC_STRIDE <= 64) ? 64 : (2 * ACC_STRIDE <= 128) ? 128 : 256;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw_u32);
    const int CH = p.CH, NR = p.NR, NSW = p.NSW;
    const bool resident = NSW == 0;
    const uint32_t ring = base;
    const uint32_t wbase = ring + NR * CH * kRowBytes;
    const uint32_t wbytes = resident ? 9 * CH * WB : NSW * WB;
    const uint32_t aux = wbase + wbytes;
    uint8_t* aux_gen = gen + (aux - base);
    // barriers: row_full[8] row_empty[8] w_full[8] w_empty[8] wres[1] tfull[2] tempty[2]
    const uint32_t bar_rfull = aux, bar_rempty = aux + 64, bar_wfull = aux + 128, bar_wempty = aux + 192;
    const uint32_t bar_wres = aux + 256, bar_tfull = aux + 264, bar_tempty = aux + 280;
    volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(aux_gen + 304);
    float* bias_s = reinterpret_cast<float*>(aux_gen + kStripBar);
    float* red_s = reinterpret_cast<float*>(aux_gen + kStripBar + 768 * 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = p.batch * p.nseg * p.strips_per_col;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&p.amap[0]);
        prefetch_tensormap(&p.wmap);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(bar_rfull + 8 * s, 1);
            mbar_init(bar_rempty + 8 * s, 1);
            mbar_init(bar_wfull + 8 * s, 1);
            mbar_init(bar_wempty + 8 * s, 1);
        }
        mbar_init(bar_wres, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, 128);
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), TMEM_COLS);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < p.n_total; i += 256) bias_s[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    auto decode = [&](int u, int& b, int& seg, int& h0, int& h1) {
        const int si = u % p.strips_per_col;
        const int t = u / p.strips_per_col;
        seg = t % p.nseg;
        b = t / p.nseg;
        h0 = si * p.L;
        h1 = h0 + p.L < p.H ? h0 + p.L : p.H;
    };

    if (warp == 0) {
        if (lane == 0) {
            // -------------------------------------------------------- input-row producer
            uint32_t e = 0;  // ring entry counter (continues across strips)
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int b, seg, h0, h1;
                decode(u, b, seg, h0, h1);
                for (int h = h0 - 1; h <= h1; ++h, ++e) {
                    const uint32_t slot = e % NR, fill = e / NR;
                    mbar_wait(bar_rempty + 8 * slot, (fill & 1) ^ 1);
                    const uint32_t full = bar_rfull + 8 * slot;
                    mbar_expect_tx(full, CH * kRowTx);
                    for (int ch = 0; ch < CH; ++ch) {
                        const bool s1 = ch >= p.chunks0;
                        tma_load_4d(ring + (slot * CH + ch) * kRowBytes, s1 ? &p.amap[1] : &p.amap[0], full,
                                    (s1 ? ch - p.chunks0 : ch) * 64, seg * 128 - 1, h, b);
                    }
                }
            }
        }
    } else if (warp == 3) {
        if (lane == 0) {
            // -------------------------------------------------------- weight producer
            if (resident) {
                mbar_expect_tx(bar_wres, 9 * CH * WB);
                for (int i = 0; i < 9 * CH; ++i) tma_load_2d(wbase + i * WB, &p.wmap, bar_wres, i * 64, 0);
            } else {
                uint32_t ws = 0, wph = 0;
                for (int u = blockIdx.x; u < units; u += gridDim.x) {
                    int b, seg, h0, h1;
                    decode(u, b, seg, h0, h1);
                    for (int h = h0; h < h1; ++h)
                        for (int i = 0; i < 9 * CH; ++i) {  // K order = (tap, chunk), as the weight matrix
                            mbar_wait(bar_wempty + 8 * ws, wph ^ 1);
                            mbar_expect_tx(bar_wfull + 8 * ws, WB);
                            tma_load_2d(wbase + ws * WB, &p.wmap, bar_wfull + 8 * ws, i * 64, 0);
                            if (++ws == static_cast<uint32_t>(NSW)) {
                                ws = 0;
                                wph ^= 1;
                            }
                        }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // -------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc = make_idesc_f16(128, BN);
            uint32_t e = 0, it = 0, ws = 0, wph = 0;
            if (resident) {
                mbar_wait(bar_wres, 0);
                tc_fence_after();
            }
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int b, seg, h0, h1;
                decode(u, b, seg, h0, h1);
                const int rows = h1 - h0;
                for (int j = 0; j < rows; ++j, ++it) {
                    const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
                    if (j == 0) {
                        mbar_wait(bar_rfull + 8 * (e % NR), (e / NR) & 1);
                        mbar_wait(bar_rfull + 8 * ((e + 1) % NR), ((e + 1) / NR) & 1);
                    }
                    mbar_wait(bar_rfull + 8 * ((e + j + 2) % NR), ((e + j + 2) / NR) & 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
                    uint32_t first = 0;
                    for (int tap = 0; tap < 9; ++tap) {
                        const int kh = tap / 3, kw = tap - 3 * kh;
                        const uint32_t slot = (e + j + kh) % NR;
                        for (int ch = 0; ch < CH; ++ch) {
                            uint32_t wb;
                            if (resident) {
                                wb = wbase + (tap * CH + ch) * WB;
                            } else {
                                mbar_wait(bar_wfull + 8 * ws, wph);
                                tc_fence_after();
                                wb = wbase + ws * WB;
                            }
                            const uint64_t adesc = make_sw128_desc(ring + (slot * CH + ch) * kRowBytes + kw * 128);
                            const uint64_t bdesc = make_sw128_desc(wb);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_f16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, first);
                                first = 1;
                            }
                            if (!resident) {
                                umma_commit(bar_wempty + 8 * ws);
                                if (++ws == static_cast<uint32_t>(NSW)) {
                                    ws = 0;
                                    wph ^= 1;
                                }
                            }
                        }
                    }
                    umma_commit(bar_rempty + 8 * ((e + j) % NR));  // input row h-1 has served its last output row
                    if (j == rows - 1) {
                        umma_commit(bar_rempty + 8 * ((e + j + 1) % NR));
                        umma_commit(bar_rempty + 8 * ((e + j + 2) % NR));
                    }
                    umma_commit(bar_tfull + 8 * as);
                }
                e += rows + 2;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const EpiArgs ea{p.out, p.residual, p.ldc, p.x, p.xpad, p.x0_out, p.c0, p.c1};
        uint32_t it = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int b, seg, h0, h1;
            decode(u, b, seg, h0, h1);
            const int gx = seg * 128 + row;
            const bool valid = gx < p.W;
            for (int h = h0; h < h1; ++h, ++it) {
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + gx;
                mbar_wait(bar_tfull + 8 * as, aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * ACC_STRIDE;
                float* sdst = (EPI == EPI_STATS)
                                  ? p.stats + ((static_cast<size_t>(b) * p.H * p.nseg + static_cast<size_t>(h) * p.nseg + seg) * 32) * 2
                                  : nullptr;
                conv_epilogue_tile<BN, CPG, EPI>(ea, taddr, bar_tempty + 8 * as, bias_s, red_s + (it & 1) * (4 * 32 * 2),
                                                 q, lane, valid, pix, 0, sdst);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host
int strip_smem_bytes(int bn, int CH, int NR, int NSW) {
    const int wb = bn * 128;
    return 1024 + NR * CH * kRowBytes + (NSW == 0 ? 9 * CH * wb : NSW * wb) + kStripAux;
}

bool strip_plan(int bn, int CH, int* NR, int* NSW) {
    const int limit = 227 * 1024;
    for (int nr = 6; nr >= 4; --nr)  // resident weights first
        if (strip_smem_bytes(bn, CH, nr, 0) <= limit) {
            *NR = nr;
            *NSW = 0;
            return true;
        }
    for (int nr = 6; nr >= 4; --nr)
        for (int nsw = 6; nsw >= 3; --nsw)
            if (strip_smem_bytes(bn, CH, nr, nsw) <= limit) {
                *NR = nr;
                *NSW = nsw;
                return true;
            }
    return false;
}

bool strip_inst_ok(int bn, int cpg, int epi) {
    if (epi == EPI_STORE) return bn == 64 || bn == 128;
    if (epi == EPI_STATS) return (bn == 64 && cpg == 2) || (bn == 128 && cpg == 4);
    return epi == EPI_DDIM && bn == 16;
}

#define STRIP_ALL_CASES()        \
    STRIP_CASE(64, 1, EPI_STORE)  \
    STRIP_CASE(128, 1, EPI_STORE) \
    STRIP_CASE(64, 2, EPI_STATS)  \
    STRIP_CASE(128, 4, EPI_STATS) \
    STRIP_CASE(16, 1, EPI_DDIM)

cudaError_t configure_strip_kernels() {
    cudaError_t e;
#define STRIP_CASE(BN_, CPG_, EPI_)                                                                         \
    if ((e = cudaFuncSetAttribute(conv_strip_kernel<BN_, CPG_, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  227 * 1024)) != cudaSuccess)                                             \
        return e;
    STRIP_ALL_CASES()
#undef STRIP_CASE
    return cudaSuccess;
}

cudaError_t launch_conv_strip(const StripParams& p, int bn, int cpg, int epi, int num_sms, cudaStream_t stream) {
    if (epi != EPI_STATS) cpg = 1;
    const int units = p.batch * p.nseg * p.strips_per_col;
    const int grid = units < num_sms ? units : num_sms;
    const int smem = strip_smem_bytes(bn, p.CH, p.NR, p.NSW);
#define STRIP_CASE(BN_, CPG_, EPI_)                                        \
    if (bn == BN_ && cpg == CPG_ && epi == EPI_) {                         \
        conv_strip_kernel<BN_, CPG_, EPI_><<<grid, 256, smem, stream>>>(p); \
        return cudaGetLastError();                                         \
    }
    STRIP_ALL_CASES()
#undef STRIP_CASE
    return cudaErrorInvalidValue;
}

}  // namespace cdc
