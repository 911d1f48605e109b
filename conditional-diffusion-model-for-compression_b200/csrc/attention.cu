// K-attn: fused flash-style softmax(Q K^T / 8) V for the 1/16-resolution mid block
// (4 heads x d = 64, N = HW/256 tokens).  Scores never touch HBM.  The product path is the tcgen05 kernel at the end
// of this file; the mma.sync kernel below is compiled only into the tools build (-DCDC_TOOLS) as its A/B reference.
// Oracle counterpart: oracle/unet.py Attn.forward.
#include "kernels.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace cdc {

constexpr int kAttD = 64;

#ifdef CDC_TOOLS  // the mma.sync kernel is the A/B reference of tools/ builds only (libcdc_b200_tools.so)
constexpr int kAttBQ = 64, kAttBK = 64, kAttPitch = 72;  // 16-bit elements per smem row
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    const int sz = pred ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32." CDC_MMA_SYNC_T "." CDC_MMA_SYNC_T ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// 8 warps: two groups of 4.  Group g takes the K/V tiles j = g, g+2, ... (its own double buffers and named barrier),
// so the serial chain of tiles is half as long (a 64-query CTA at N = 1536 walks 12 tiles instead of 24 -- the kernel
// is latency-bound at this size); the two partial (O, m, l) states are merged through shared memory at the end.
constexpr int kAttThreads = 256;
constexpr int kAttTile = kAttBK * kAttPitch;                                           // elements of one K or V tile
constexpr int kAttSmemBytes = (kAttBQ * kAttPitch + 2 * 2 * 2 * kAttTile) * 2;         // Q + [group][K|V][buffer]

__global__ void __launch_bounds__(kAttThreads) attention_kernel(const act_t* qkv, act_t* out, int N, int heads) {
    extern __shared__ __align__(16) uint8_t att_smem[];
    act_t(*Qs)[kAttPitch] = reinterpret_cast<act_t(*)[kAttPitch]>(att_smem);
    act_t* kv0 = reinterpret_cast<act_t*>(att_smem) + kAttBQ * kAttPitch;
    const int C = heads * kAttD, ld = 3 * C;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAttBQ;
    const int tid = threadIdx.x, grp = tid >> 7, gtid = tid & 127, warp = (tid >> 5) & 3, lane = tid & 31, g = lane >> 2, t = lane & 3;
    act_t(*Ks)[kAttBK][kAttPitch] = reinterpret_cast<act_t(*)[kAttBK][kAttPitch]>(kv0 + (grp * 4 + 0) * kAttTile);  // [buffer]
    act_t(*Vs)[kAttBK][kAttPitch] = reinterpret_cast<act_t(*)[kAttBK][kAttPitch]>(kv0 + (grp * 4 + 2) * kAttTile);
    const act_t* base = qkv + static_cast<size_t>(b) * N * ld + h * kAttD;
    pdl_wait();  // qkv is written by the preceding conv (it is only read through cp.async below)
    pdl_launch_dependents();  // (after the wait: see launch.cuh)

    auto load_tile = [&](act_t (*dst)[kAttPitch], const act_t* src, int row0, int first, int nthr) {
        for (int i = first; i < 64 * 8; i += nthr) {
            const int r = i >> 3, c = i & 7;
            const bool ok = row0 + r < N;
            const act_t* sp = src + static_cast<size_t>(ok ? row0 + r : 0) * ld + c * 8;
            cp_async16(smem_u32(&dst[r][c * 8]), sp, ok);
        }
    };
    const int ntiles = (N + kAttBK - 1) / kAttBK;
    load_tile(Qs, base, q0, tid, kAttThreads);
    if (grp < ntiles) {
        load_tile(Ks[0], base + C, grp * kAttBK, gtid, 128);
        load_tile(Vs[0], base + 2 * C, grp * kAttBK, gtid, 128);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();  // Q (loaded by both groups) and every group's first tile

    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    uint32_t qf[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        const int r = warp * 16 + g, c = kk * 16 + 2 * t;
        qf[kk][0] = *reinterpret_cast<const uint32_t*>(&Qs[r][c]);
        qf[kk][1] = *reinterpret_cast<const uint32_t*>(&Qs[r + 8][c]);
        qf[kk][2] = *reinterpret_cast<const uint32_t*>(&Qs[r][c + 8]);
        qf[kk][3] = *reinterpret_cast<const uint32_t*>(&Qs[r + 8][c + 8]);
    }

    int buf = 0;
    for (int j = grp; j < ntiles; j += 2, buf ^= 1) {
        if (j + 2 < ntiles) {
            load_tile(Ks[buf ^ 1], base + C, (j + 2) * kAttBK, gtid, 128);
            load_tile(Vs[buf ^ 1], base + 2 * C, (j + 2) * kAttBK, gtid, 128);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        named_bar_sync(1 + grp, 128);
        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[buf][n * 8 + g][kk * 16 + 2 * t]);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[buf][n * 8 + g][kk * 16 + 8 + 2 * t]);
                mma_16816(s[n], qf[kk], b0, b1);
            }
        }
        // scale to log2 domain, mask the key tail
        const int key0 = j * kAttBK;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = key0 + n * 8 + 2 * t + (e & 1);
                s[n][e] = key < N ? s[n][e] * sl2 : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
            mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float sc0 = exp2f(m0 - mn0), sc1 = exp2f(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[4][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p0 = exp2f(s[n][0] - mn0), p1 = exp2f(s[n][1] - mn0);
            const float p2 = exp2f(s[n][2] - mn1), p3 = exp2f(s[n][3] - mn1);
            rs0 += p0 + p1;
            rs1 += p2 + p3;
            pf[n >> 1][(n & 1) * 2 + 0] = pack_act2(p0, p1);
            pf[n >> 1][(n & 1) * 2 + 1] = pack_act2(p2, p3);
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1);
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * sc0 + rs0;
        l1 = l1 * sc1 + rs1;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            o[n][0] *= sc0;
            o[n][1] *= sc0;
            o[n][2] *= sc1;
            o[n][3] *= sc1;
        }
        // O += P V   (B fragments of V through ldmatrix.trans: rows = keys, cols = d)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                uint32_t b0, b1;
                const int r = kk * 16 + (lane & 15);
                ldmatrix_x2_trans(b0, b1, smem_u32(&Vs[buf][r][n * 8]));
                mma_16816(o[n], pf[kk], b0, b1);
            }
        }
        named_bar_sync(1 + grp, 128);
    }
    // merge the two groups' online-softmax states: group 1 -> shared memory -> group 0
    __syncthreads();  // all tiles consumed: the K/V buffers can be reused as scratch
    float* scratch = reinterpret_cast<float*>(kv0) + gtid * 36;
    if (grp == 1) {
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            scratch[n * 4 + 0] = o[n][0];
            scratch[n * 4 + 1] = o[n][1];
            scratch[n * 4 + 2] = o[n][2];
            scratch[n * 4 + 3] = o[n][3];
        }
        scratch[32] = m0;
        scratch[33] = m1;
        scratch[34] = l0;
        scratch[35] = l1;
    }
    __syncthreads();
    if (grp == 1) return;
    {
        const float mb0 = scratch[32], mb1 = scratch[33];
        const float mn0 = fmaxf(m0, mb0), mn1 = fmaxf(m1, mb1);  // group 0 always has a tile: finite
        const float sa0 = exp2f(m0 - mn0), sa1 = exp2f(m1 - mn1), sb0 = exp2f(mb0 - mn0), sb1 = exp2f(mb1 - mn1);
        l0 = l0 * sa0 + scratch[34] * sb0;
        l1 = l1 * sa1 + scratch[35] * sb1;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            o[n][0] = o[n][0] * sa0 + scratch[n * 4 + 0] * sb0;
            o[n][1] = o[n][1] * sa0 + scratch[n * 4 + 1] * sb0;
            o[n][2] = o[n][2] * sa1 + scratch[n * 4 + 2] * sb1;
            o[n][3] = o[n][3] * sa1 + scratch[n * 4 + 3] * sb1;
        }
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        const int d = n * 8 + 2 * t;
        if (r0 < N)
            *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * N + r0) * C + h * kAttD + d) =
                pack_act2(o[n][0] * inv0, o[n][1] * inv0);
        if (r1 < N)
            *reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(b) * N + r1) * C + h * kAttD + d) =
                pack_act2(o[n][2] * inv1, o[n][3] * inv1);
    }
}

cudaError_t configure_attention() {
    return cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmemBytes);
}

cudaError_t launch_attention(const act_t* qkv, act_t* o, int B, int N, int heads, cudaStream_t s) {
    static const cudaError_t cfg = configure_attention();  // > 48 KB of dynamic shared memory
    if (cfg != cudaSuccess) return cfg;
    dim3 grid((N + kAttBQ - 1) / kAttBQ, heads, B);
    return launch_pdl(attention_kernel, grid, dim3(kAttThreads), kAttSmemBytes, s, qkv, o, N, heads);
}

#endif  // CDC_TOOLS

// =====================================================================================================================
// tcgen05 version: S = Q K^T and O_tile = P V run on the 5th-generation tensor cores with TMEM accumulators.
//
//   CTA = 128 queries of one head; K/V tiles of 128 keys stream through a 2-deep TMA ring.
//   warp 0 : TMA producer (Q once, then K and V tiles; rows past N are zero-filled by the 3-D tensor map)
//   warp 1 : TMEM allocator + MMA issuer: S_b = Q K_b^T (M 128 x N 128 x K 64, both operands K-major), issued one tile
//            AHEAD of  O_b = P_b V_b  (M 128 x N 64 x K 128; A = P from shared memory, B = V as an MN-major operand:
//            the [key][d] tile exactly as TMA lands it, instruction-descriptor bit 16)
//   warps 2-5, 6-9 : two softmax warpgroups, thread = query row = TMEM lane; group g owns the tiles j = g (mod 2),
//            i.e. TMEM / P buffer g, and keeps its OWN online-softmax state (m, l, o) -- the two streams are merged
//            through shared memory at the end, so while one group does exp2 the tensor core serves the other.
//            Two passes over the S row in TMEM (row max, then exp2 / sum / fp16 P -> 128-byte-swizzled shared
//            memory) keep only 32 score registers live; the running output stays in REGISTERS
//            (o = o * scale + O_tile, O_tile read back with tcgen05.ld): no TMEM read-modify-write correction pass.
// Oracle counterpart: oracle/unet.py Attn.forward (softmax(q k^T / 8) v per head).
// =====================================================================================================================
constexpr int kTcThreads = 320;  // TMA warp, MMA warp, two softmax warpgroups
constexpr int kTcTile = 128 * 128;                               // bytes of one [128 rows][64 x 16-bit] tile
constexpr int kTcSmemBytes = 1024 + kTcTile * (1 + 2 + 2 + 4) + 256;  // Q, K[2], V[2], P[2][2 key blocks], barriers

__device__ __forceinline__ float ex2_approx(float x) {  // one MUFU op; ex2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* desc, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
    extern __shared__ uint8_t att_raw[];
    const uint32_t raw = smem_u32(att_raw), base = (raw + 1023u) & ~1023u;
    uint8_t* gen = att_raw + (base - raw);
    const uint32_t sQ = base, sK = sQ + kTcTile, sV = sK + 2 * kTcTile, sP = sV + 2 * kTcTile, bars = sP + 4 * kTcTile;
    // barriers: q_full | k_full[2] k_empty[2] v_full[2] v_empty[2] | s_full[2] s_empty[2] p_full[2] o_full[2] o_empty[2]
    const uint32_t b_qf = bars, b_kf = bars + 8, b_ke = bars + 24, b_vf = bars + 40, b_ve = bars + 56;
    const uint32_t b_sf = bars + 72, b_se = bars + 88, b_pf = bars + 104, b_of = bars + 120, b_oe = bars + 136;
    volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(gen + 9 * kTcTile + 160);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = p.N, C = p.heads * kAttD;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
    const int T = (N + 127) / 128;

    if (threadIdx.x == 0) {
        stamp_begin(p.stamp);
        prefetch_tensormap(&p.qkv_map);
        mbar_init(b_qf, 1);
    }
    if (warp == 0 && lane < 2) {  // (buffer i's barriers by lane i)
        const int i = lane;
        mbar_init(b_kf + 8 * i, 1);
        mbar_init(b_ke + 8 * i, 1);
        mbar_init(b_vf + 8 * i, 1);
        mbar_init(b_ve + 8 * i, 1);
        mbar_init(b_sf + 8 * i, 1);
        mbar_init(b_se + 8 * i, 128);
        mbar_init(b_pf + 8 * i, 128);
        mbar_init(b_of + 8 * i, 1);
        mbar_init(b_oe + 8 * i, 128);
    }
    if (warp == 0) {
        fence_mbar_init();
        __syncwarp();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_holder)), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    pdl_wait();
    pdl_launch_dependents();  // (after the wait: see launch.cuh)

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(b_qf, kTcTile);
            tma_load_3d(sQ, &p.qkv_map, b_qf, h * kAttD, q0, b);
            for (int j = 0; j < T; ++j) {
                const uint32_t buf = j & 1, ph = (j >> 1) & 1;
                mbar_wait(b_ke + 8 * buf, ph ^ 1);
                mbar_expect_tx(b_kf + 8 * buf, kTcTile);
                tma_load_3d(sK + buf * kTcTile, &p.qkv_map, b_kf + 8 * buf, C + h * kAttD, j * 128, b);
                mbar_wait(b_ve + 8 * buf, ph ^ 1);
                mbar_expect_tx(b_vf + 8 * buf, kTcTile);
                tma_load_3d(sV + buf * kTcTile, &p.qkv_map, b_vf + 8 * buf, 2 * C + h * kAttD, j * 128, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        constexpr uint32_t idesc_s = make_idesc_f16(128, 128);
        constexpr uint32_t idesc_o = make_idesc_f16(128, 64) | (1u << 16);  // B (= V) is MN-major
        const uint64_t desc_hi = make_sw128_desc(0) & 0xFFFFFFFF00000000ull;
        auto issue_s = [&](int t) {  // S_buf = Q K_t^T
            const uint32_t buf = t & 1, ph = (t >> 1) & 1;
            mbar_wait(b_kf + 8 * buf, ph);
            mbar_wait(b_se + 8 * buf, ph ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t alo = sQ >> 4, blo = (sK + buf * kTcTile) >> 4;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16_ss(tmem + buf * 128, desc_hi | (alo + 2 * k), desc_hi | (blo + 2 * k), idesc_s, k != 0 ? 1u : 0u);
                umma_commit(b_sf + 8 * buf);
                umma_commit(b_ke + 8 * buf);
            }
            __syncwarp();
        };
        mbar_wait(b_qf, 0);
        issue_s(0);
        for (int j = 0; j < T; ++j) {
            const uint32_t buf = j & 1, ph = (j >> 1) & 1;
            if (j + 1 < T) issue_s(j + 1);
            mbar_wait(b_pf + 8 * buf, ph);
            mbar_wait(b_vf + 8 * buf, ph);
            mbar_wait(b_oe + 8 * buf, ph ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {  // O_buf = P_buf V_buf: two 64-key blocks of P x four 16-key steps
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t alo = ((sP + (buf * 2 + kb) * kTcTile) >> 4) + 2 * k;
                        const uint32_t blo = (sV + buf * kTcTile + (kb * 64 + k * 16) * 128) >> 4;  // 16 key rows further down
                        umma_f16_ss(tmem + 256 + buf * 64, desc_hi | alo, desc_hi | blo, idesc_o, (kb | k) != 0 ? 1u : 0u);
                    }
                umma_commit(b_of + 8 * buf);
                umma_commit(b_ve + 8 * buf);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3, row = q * 32 + lane, grp = (warp - 2) >> 2;
        const uint32_t tq = tmem + (static_cast<uint32_t>(q * 32) << 16);
        const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
        float m = -INFINITY, l = 0.f;
        float o[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = 0.f;
        for (int j = grp; j < T; j += 2) {  // this group's tiles all use buffer `grp`
            const uint32_t buf = grp, ph = (j >> 1) & 1;
            mbar_wait(b_sf + 8 * buf, ph);
            tc_fence_after();
            const int key0 = j * 128;
            const int nkeys = N - key0 < 128 ? N - key0 : 128;  // valid keys of this tile (only the last tile is ragged)
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {  // pass 1: row maximum
                uint32_t v[32];
                tmem_ld32(tq + buf * 128 + c * 32, v);
                tmem_ld_wait();
                if (nkeys == 128) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c * 32 + i < nkeys) mx = fmaxf(mx, __uint_as_float(v[i]));
                }
            }
            const float m_new = fmaxf(m, mx * sl2);  // every tile holds at least one valid key: finite
            const float sc = ex2_approx(m - m_new);
            float rs = 0.f;
            const uint32_t prow = sP + buf * 2 * kTcTile + row * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {  // pass 2: p = exp2(s - m), row sum, fp16 P into the swizzled A-operand tile
                uint32_t v[32];
                tmem_ld32(tq + buf * 128 + c * 32, v);
                tmem_ld_wait();
                if (c == 3) {  // S_buf fully read: the issuer may overwrite it with tile j + 2
                    tc_fence_before();
                    mbar_arrive(b_se + 8 * buf);
                }
                float pv[32];
                if (nkeys == 128) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        pv[i] = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m_new));
                        rs += pv[i];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        pv[i] = c * 32 + i < nkeys ? ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m_new)) : 0.f;
                        rs += pv[i];
                    }
                }
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int key = c * 32 + s4 * 8, kb = key >> 6, ck = (key & 63) >> 3;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + kb * kTcTile + ((ck ^ (row & 7)) << 4)),
                                 "r"(pack_act2(pv[s4 * 8 + 0], pv[s4 * 8 + 1])), "r"(pack_act2(pv[s4 * 8 + 2], pv[s4 * 8 + 3])),
                                 "r"(pack_act2(pv[s4 * 8 + 4], pv[s4 * 8 + 5])), "r"(pack_act2(pv[s4 * 8 + 6], pv[s4 * 8 + 7]))
                                 : "memory");
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(b_pf + 8 * buf);
            l = fmaf(l, sc, rs);
            m = m_new;
            // o = o * scale + P_j V_j (the other group's softmax overlaps this wait)
            mbar_wait(b_of + 8 * buf, ph);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(tq + 256 + buf * 64 + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], sc, __uint_as_float(v[i]));
            }
            tc_fence_before();
            mbar_arrive(b_oe + 8 * buf);
        }
        // merge the two groups' streams: group 1 -> shared memory (the K/V/P tiles are dead by now) -> group 0
        named_bar_sync(1, 256);
        float* scr = reinterpret_cast<float*>(gen + kTcTile) + row * 66;  // 128 rows x 66 floats = 33 KB over the K / V tiles
        if (grp == 1) {
#pragma unroll
            for (int i = 0; i < 64; ++i) scr[i] = o[i];
            scr[64] = m;
            scr[65] = l;
        }
        named_bar_sync(1, 256);
        if (grp == 0 && q0 + row < N) {
            const float mb = scr[64], lb = scr[65];
            const float mn = fmaxf(m, mb);  // group 0 always has a tile: finite
            const float sa = ex2_approx(m - mn), sb = ex2_approx(mb - mn);
            const float inv = 1.0f / (l * sa + lb * sb);
            const float wa = sa * inv, wb = sb * inv;
            uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * N + q0 + row) * C + h * kAttD);
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    w[e] = pack_act2(o[s8 * 16 + 2 * e] * wa + scr[s8 * 16 + 2 * e] * wb, o[s8 * 16 + 2 * e + 1] * wa + scr[s8 * 16 + 2 * e + 1] * wb);
                st_global_v8(dst + 2 * s8, make_uint4(w[0], w[1], w[2], w[3]), make_uint4(w[4], w[5], w[6], w[7]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp_end(p.stamp);
    if (warp == 1) tmem_dealloc(tmem, 512);
}

cudaError_t configure_attention_tc() {
    return cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
}

cudaError_t launch_attention_tc(const AttnTcParams& p, int B, cudaStream_t s) {
    static const cudaError_t cfg = configure_attention_tc();
    if (cfg != cudaSuccess) return cfg;
    dim3 grid((p.N + 127) / 128, p.heads, B);
    return launch_pdl(attention_tc_kernel, grid, dim3(kTcThreads), kTcSmemBytes, s, p);
}

}  // namespace cdc
