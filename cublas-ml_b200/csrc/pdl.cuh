// Programmatic dependent launch for the kernels of one training step.
//
// Every kernel on the step path is launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// pdl_enter(): `griddepcontrol.wait` blocks until the previous kernel of the stream has completed and its writes are
// visible, `griddepcontrol.launch_dependents` lets the NEXT kernel's CTAs become resident as soon as this grid's CTAs
// have all started and SMs free up.  The next kernel's launch latency and prologue (barrier init, TMEM allocation,
// tensor-map prefetch) then overlap this kernel's tail instead of following it.  Rule kept everywhere: a kernel launched
// with the attribute executes pdl_enter() before it touches global memory, so completion stays transitive along the
// stream.  B2NN_PDL=0 launches everything fully serialised (for A/B measurements).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <utility>

namespace b2 {

__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
}

// Opt a kernel into more than 48 KB of dynamic shared memory once per DEVICE (the attribute is per device; a process
// may hold contexts on several GPUs).  `done` is a function-local static of the launch site: one bit per device ordinal.
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, int bytes, std::atomic<unsigned long long>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

inline bool pdl_enabled() {
    static const bool on = [] { const char* e = std::getenv("B2NN_PDL"); return !(e && std::atoi(e) == 0); }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace b2
