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
//   warp 3 : weight producer (TMA)   warps 4-11 : epilogue (shared with conv_tc.cu)
//
// Oracle counterpart: oracle/unet.py `conv(k=3)` inside RB / stem / final (the reference ships no code).
#include <stdio.h>

#include "conv_epilogue.cuh"
#include "conv_strip.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace cdc {

constexpr int kRowBytes = 17 * 1024;       // 130 pixels x 128 B = 16640, padded to a 1024 B multiple
constexpr int kRowTx = 130 * 128;          // bytes one row box delivers
constexpr int kStripBar = 512;             // barrier block
constexpr int kStripAux = kStripBar + 768 * 4 + 2 * 4 * 32 * 2 * 4;

// CH (64-channel chunks of C_in) and RES (weights resident in shared memory) are compile-time so that the
// issuer's 36*CH tcgen05.mma per output row are straight-line code with immediate descriptor offsets.
template <int BN, int CPG, int EPI, int CH, bool RES>
__global__ void __launch_bounds__(128 + kEpiThreads, 1) conv_strip_kernel(const __grid_constant__ StripParams p) {
    constexpr int WB = BN * 128;  // one (tap, chunk) weight slice
    constexpr int ACC_STRIDE = BN < 32 ? 32 : BN;
    // Streamed weights: every weight stage feeds R = 2 output rows (two accumulators), which halves the weight
    // traffic and the issuer's waits / commits per MMA.  Resident weights: R = 1.
    constexpr int R = RES ? 1 : 2;
    constexpr int BUF_STRIDE = R * ACC_STRIDE;  // TMEM columns of one accumulator group
    constexpr int TMEM_COLS = (2 * BUF_STRIDE <= 32) ? 32 : (2 * BUF_STRIDE <= 64) ? 64 : (2 * BUF_STRIDE <= 128) ? 128 : (2 * BUF_STRIDE <= 256) ? 256 : 512;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - raw_u32);
    const int NR = p.NR, NSW = p.NSW;
    constexpr bool resident = RES;
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
            mbar_init(bar_tempty + 8 * s, EPI == EPI_DDIM ? 128 : kEpiThreads);
        }
        fence_mbar_init();
    }
    if (warp == 2) {  // (warp-collective)
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), TMEM_COLS);
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < p.n_total; i += 128 + kEpiThreads) bias_s[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_launch_dependents();
    pdl_wait();

    auto decode = [&](int u, int& b, int& seg, int& h0, int& h1) {
        const int si = u % p.strips_per_col;
        const int t = u / p.strips_per_col;
        seg = t % p.nseg;
        b = t / p.nseg;
        h0 = si * p.L;
        h1 = h0 + p.L < p.H ? h0 + p.L : p.H;
    };

    // Producers and issuer run warp-converged; only the async instruction is predicated on lane 0.
    const uint32_t leader = lane == 0 ? 1u : 0u;
    if (warp == 0) {
        // ------------------------------------------------------------ input-row producer
        uint32_t slot = 0, par = 0;  // ring position / fill parity (continues across strips)
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int b, seg, h0, h1;
            decode(u, b, seg, h0, h1);
            const int w0 = seg * 128 - 1;
            for (int h = h0 - 1; h <= h1; ++h) {
                mbar_wait(bar_rempty + 8 * slot, par ^ 1);
                const uint32_t full = bar_rfull + 8 * slot;
                mbar_expect_tx_p(full, CH * kRowTx, leader);
                uint32_t dst = ring + slot * CH * kRowBytes;
                for (int ch = 0; ch < CH; ++ch, dst += kRowBytes) {
                    const bool s1 = ch >= p.chunks0;
                    tma_load_4d_p(dst, s1 ? &p.amap[1] : &p.amap[0], full, (s1 ? ch - p.chunks0 : ch) * 64, w0, h, b, leader);
                }
                if (++slot == static_cast<uint32_t>(NR)) {
                    slot = 0;
                    par ^= 1;
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ weight producer
        if (resident) {
            mbar_expect_tx_p(bar_wres, 9 * CH * WB, leader);
            for (int i = 0; i < 9 * CH; ++i) tma_load_2d_p(wbase + i * WB, &p.wmap, bar_wres, i * 64, 0, leader);
        } else {
            uint32_t ws = 0, wph = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int b, seg, h0, h1;
                decode(u, b, seg, h0, h1);
                for (int h = h0; h < h1; h += R)            // one pass over the weights per group of R output rows
                    for (int i = 0; i < 9 * CH; ++i) {  // K order = (tap, chunk), as the weight matrix
                        mbar_wait(bar_wempty + 8 * ws, wph ^ 1);
                        mbar_expect_tx_p(bar_wfull + 8 * ws, WB, leader);
                        tma_load_2d_p(wbase + ws * WB, &p.wmap, bar_wfull + 8 * ws, i * 64, 0, leader);
                        if (++ws == static_cast<uint32_t>(NSW)) {
                            ws = 0;
                            wph ^= 1;
                        }
                    }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // The loop is kept free of divisions and 64-bit math: ring slots / parities advance
        // incrementally and descriptors differ only in their low word.
        constexpr uint32_t idesc = make_idesc_f16(128, BN);
        const uint64_t desc_hi = make_sw128_desc(0) & 0xFFFFFFFF00000000ull;
        const uint32_t slot_stride = static_cast<uint32_t>(CH) * kRowBytes;
        uint32_t it = 0, ws = 0, wph = 0;
        uint32_t wslot = 0, wpar = 0;  // next ring entry to wait for
        uint32_t aslot = 0;            // ring slot of input row (h - 1) of the current output row
        if (resident) {
            mbar_wait(bar_wres, 0);
            tc_fence_after();
        }
        auto wait_row = [&]() {
            mbar_wait(bar_rfull + 8 * wslot, wpar);
            if (++wslot == static_cast<uint32_t>(NR)) {
                wslot = 0;
                wpar ^= 1;
            }
        };
        auto next_slot = [&](uint32_t s_) { return s_ + 1 == static_cast<uint32_t>(NR) ? 0u : s_ + 1; };
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int b, seg, h0, h1;
            decode(u, b, seg, h0, h1);
            const int rows = h1 - h0;
            wait_row();
            wait_row();
            for (int j = 0; j < rows; j += R, ++it) {
                const int nr = (R == 2 && j + 1 < rows) ? 2 : 1;  // output rows in this group
                const bool last = j + R >= rows;
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && it < 64;
                if (dbg) p.dbg[it * 4 + 0] = clock64();
                mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
                if (dbg) p.dbg[it * 4 + 1] = clock64();
                for (int i = 0; i < nr; ++i) wait_row();
                tc_fence_after();
                if (dbg) p.dbg[it * 4 + 2] = clock64();
                const uint32_t d_tmem = tmem_base + as * BUF_STRIDE;
                const uint32_t s1 = next_slot(aslot), s2 = next_slot(s1), s3 = next_slot(s2);
                const uint32_t rowaddr[4] = {ring + aslot * slot_stride, ring + s1 * slot_stride, ring + s2 * slot_stride,
                                             ring + s3 * slot_stride};
                // Elected-lane region with 32-bit descriptor low words and compile-time accumulate flags:
                // 57-64 cycles per MMA (the M=128 operand-fetch floor) instead of 115 for a predicated-asm
                // formulation of the same loop (tools/exp_mma_rate.cu, modes 4 / 5).
                if (leader) {
                    uint32_t wb = wbase;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                            for (int ch = 0; ch < CH; ++ch) {
                                if (!resident) {
                                    mbar_wait(bar_wfull + 8 * ws, wph);
                                    tc_fence_after();
                                    wb = wbase + ws * WB;
                                }
                                const uint32_t blo = (wb >> 4) & 0x3FFFu;
                                const uint32_t alo0 = ((rowaddr[kh] + kw * 128 + ch * kRowBytes) >> 4) & 0x3FFFu;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_f16_ss(d_tmem, desc_hi | (alo0 + 2 * k), desc_hi | (blo + 2 * k), idesc,
                                                (kh | kw | k | ch) != 0 ? 1u : 0u);
                                if (R == 2 && nr == 2) {
                                    const uint32_t alo1 = ((rowaddr[kh + 1] + kw * 128 + ch * kRowBytes) >> 4) & 0x3FFFu;
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        umma_f16_ss(d_tmem + ACC_STRIDE, desc_hi | (alo1 + 2 * k), desc_hi | (blo + 2 * k),
                                                    idesc, (kh | kw | k | ch) != 0 ? 1u : 0u);
                                }
                                if (resident) {
                                    wb += WB;
                                } else {
                                    umma_commit(bar_wempty + 8 * ws);
                                    if (++ws == static_cast<uint32_t>(NSW)) {
                                        ws = 0;
                                        wph ^= 1;
                                    }
                                }
                            }
                        }
                        if (kh == 0) umma_commit(bar_rempty + 8 * aslot);  // input row h-1 is only read by the kh = 0 taps
                    }
                    if (nr == 2) umma_commit(bar_rempty + 8 * s1);
                    if (last) {  // the remaining rows of this strip are not needed by anyone
                        if (nr == 2) {
                            umma_commit(bar_rempty + 8 * s2);
                            umma_commit(bar_rempty + 8 * s3);
                        } else {
                            umma_commit(bar_rempty + 8 * s1);
                            umma_commit(bar_rempty + 8 * s2);
                        }
                    }
                    umma_commit(bar_tfull + 8 * as);
                }
                __syncwarp();
                aslot = last ? next_slot(nr == 2 ? s3 : s2) : (nr == 2 ? s2 : s1);
                if (dbg) p.dbg[it * 4 + 3] = clock64();
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const EpiArgs ea{p.out, p.residual, p.ldc, p.x, p.xpad, p.x0_out, p.c0, p.c1};
        uint32_t it = 0, tile_ctr = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int b, seg, h0, h1;
            decode(u, b, seg, h0, h1);
            const int gx = seg * 128 + row;
            const bool valid = gx < p.W;
            for (int h = h0; h < h1; h += R, ++it) {
                const int nr = (R == 2 && h + 1 < h1) ? 2 : 1;
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                long long* edbg = (p.dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0 && it < 32) ? p.dbg + 256 + it * 8 : nullptr;
                if (edbg) edbg[4] = clock64();
                mbar_wait(bar_tfull + 8 * as, aphase);
                tc_fence_after();
                if (edbg) edbg[0] = clock64();
                if (p.dbg != nullptr && p.dbg[511] == 1) {  // tools only: measure the MMA phase without epilogue work
                    tc_fence_before();
                    mbar_arrive(bar_tempty + 8 * as);
                    continue;
                }
                for (int r = 0; r < nr; ++r, ++tile_ctr) {
                    const size_t pix = (static_cast<size_t>(b) * p.H + (h + r)) * p.W + gx;
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BUF_STRIDE + r * ACC_STRIDE;
                    float* sdst = (EPI == EPI_STATS)
                                      ? p.stats + ((static_cast<size_t>(b) * p.H * p.nseg + static_cast<size_t>(h + r) * p.nseg + seg) * 32) * 2
                                      : nullptr;
                    conv_epilogue_tile<BN, CPG, EPI>(ea, taddr, bar_tempty + 8 * as, bias_s, red_s + (tile_ctr & 1) * (2 * 4 * 16 * 2),
                                                     q, half, lane, valid, pix, 0, sdst, edbg, r == nr - 1);
                }
                if (edbg) edbg[5] = clock64();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ host
// (BN, CPG, EPI, CH, RESIDENT) instantiations: exactly the layer shapes of the UNet / context net that
// the strip variant serves; anything else goes through conv_tc.cu.
#define STRIP_ALL_CASES()                  \
    STRIP_CASE(64, 2, EPI_STATS, 1, true)   \
    STRIP_CASE(64, 2, EPI_STATS, 2, false)  \
    STRIP_CASE(64, 1, EPI_STORE, 1, true)   \
    STRIP_CASE(64, 1, EPI_STORE, 2, false)  \
    STRIP_CASE(128, 4, EPI_STATS, 1, true)  \
    STRIP_CASE(128, 4, EPI_STATS, 2, false) \
    STRIP_CASE(128, 1, EPI_STORE, 2, false) \
    STRIP_CASE(16, 1, EPI_DDIM, 1, true)

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
    for (int nr = 6; nr >= 5; --nr)  // streamed weights run two output rows per group: 4 live rows + 1 in flight
        for (int nsw = 6; nsw >= 3; --nsw)
            if (strip_smem_bytes(bn, CH, nr, nsw) <= limit) {
                *NR = nr;
                *NSW = nsw;
                return true;
            }
    return false;
}

bool strip_inst_ok(int bn, int cpg, int epi, int CH, bool resident) {
#define STRIP_CASE(BN_, CPG_, EPI_, CH_, RES_) \
    if (bn == BN_ && (EPI_ != EPI_STATS || cpg == CPG_) && epi == EPI_ && CH == CH_ && resident == RES_) return true;
    STRIP_ALL_CASES()
#undef STRIP_CASE
    return false;
}

cudaError_t configure_strip_kernels() {
    cudaError_t e;
#define STRIP_CASE(BN_, CPG_, EPI_, CH_, RES_)                                                                   \
    if ((e = cudaFuncSetAttribute(conv_strip_kernel<BN_, CPG_, EPI_, CH_, RES_>,                                   \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess)      \
        return e;
    STRIP_ALL_CASES()
#undef STRIP_CASE
    return cudaSuccess;
}

cudaError_t launch_conv_strip(const StripParams& p, int bn, int cpg, int epi, int num_sms, cudaStream_t stream) {
    const int units = p.batch * p.nseg * p.strips_per_col;
    const int grid = units < num_sms ? units : num_sms;
    const int smem = strip_smem_bytes(bn, p.CH, p.NR, p.NSW);
    const bool resident = p.NSW == 0;
#define STRIP_CASE(BN_, CPG_, EPI_, CH_, RES_)                                                                \
    if (bn == BN_ && (EPI_ != EPI_STATS || cpg == CPG_) && epi == EPI_ && p.CH == CH_ && resident == RES_) {  \
        return launch_pdl(conv_strip_kernel<BN_, CPG_, EPI_, CH_, RES_>, dim3(grid), dim3(128 + kEpiThreads), smem, stream, p); \
    }
    STRIP_ALL_CASES()
#undef STRIP_CASE
    return cudaErrorInvalidValue;
}

}  // namespace cdc
