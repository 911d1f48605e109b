// Parameter block of the strip variant of K-conv (conv_strip.cu): 3x3 / stride 1 / single N tile.
#pragma once
#include "conv_tc.cuh"

namespace cdc {

struct alignas(64) StripParams {
    CUtensorMap amap[2];  // activation sources, box {64 ch, 130 px, 1 row, 1 image}
    CUtensorMap wmap;     // weights [C_out][9 * CH * 64], box {64, BN}
    int chunks0;          // 64-channel chunks that come from source 0 (the rest from source 1)
    int CH;               // chunks in total (C_in / 64)
    int H, W, batch;
    int nseg;             // ceil(W / 128) column segments
    int L;                // output rows per strip
    int strips_per_col;   // ceil(H / L)
    int NR;               // input-row ring slots
    int NSW;              // weight ring stages; 0 = weights resident in shared memory
    int ldc, n_total;
    act_t* out;
    const float* bias;
    const act_t* residual;
    float* stats;         // EPI_STATS: [batch][H * nseg][32][2]
    float* x;             // EPI_DDIM (see ConvParams)
    act_t* xpad;
    float* x0_out;
    float c0, c1;
    long long* dbg;       // optional: issuer timeline of CTA 0 (clock64 stamps), tools only
};

bool strip_inst_ok(int bn, int cpg, int epi, int CH, bool resident);
bool strip_plan(int bn, int CH, int* NR, int* NSW);  // shared-memory plan; false if it does not fit
int strip_smem_bytes(int bn, int CH, int NR, int NSW);
cudaError_t configure_strip_kernels();
cudaError_t launch_conv_strip(const StripParams& p, int bn, int cpg, int epi, int num_sms, cudaStream_t stream);

}  // namespace cdc
