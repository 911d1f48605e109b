// Launch helper.  Every kernel of the decode path CAN be launched with programmatic stream serialization (PDL) so that
// its prologue (barrier init, TMEM allocation, weight TMA loads) overlaps the tail of its predecessor; each kernel
// executes griddepcontrol.wait (pdl_wait) before it touches memory written by earlier kernels, which keeps completion
// transitive along the chain.  Measured on B200 inside the captured 17-step graph (bench.py, 768x512): plain stream
// order 29.37 ms / image, PDL with early triggers in the convs 30.15 ms, early triggers everywhere 30.67 ms (a dependent
// grid that is triggered before its predecessor's CTAs are all resident takes their SMs).  The kernels now trigger right
// AFTER their own griddepcontrol.wait (every CTA of the kernel is resident by then): 25.52 ms with PDL against 25.47 ms
// without -- no loss, but no gain either, so PDL stays OFF (the tools build, -DCDC_TOOLS, enables it with CDC_PDL=1).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

namespace cdc {

inline bool pdl_enabled() {
#ifdef CDC_TOOLS
    static const bool on = getenv("CDC_PDL") != nullptr;
    return on;
#else
    return false;  // the product build reads no environment variables
#endif
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace cdc
