// Samples (globaltimer, clock64) from one resident warp while other work runs: the slope is the SM clock the
// chip actually sustains under that load (nvidia-smi's 100 ms samples do not resolve a 26 ms graph replay).
// Built as a tiny shared library for tools/clockmon.py.
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void clockmon_kernel(long long* buf, int n, long long period_ns) {
    if (threadIdx.x != 0) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int i = 0; i < n; ++i) {
        do {
            __nanosleep(1000);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        } while (static_cast<long long>(t - t0) < period_ns * (i + 1));
        buf[2 * i] = static_cast<long long>(t);
        buf[2 * i + 1] = clock64();
    }
}
extern "C" int clockmon_launch(void* stream, long long* buf, int n, long long period_ns) {
    clockmon_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(buf, n, period_ns);
    return static_cast<int>(cudaGetLastError());
}
