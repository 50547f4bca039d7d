// Shared helpers for libpivp.so (sm_100a only).  See include/pivp.h for the C-ABI contract.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <utility>

#define PIVP_OK 0
#define PIVP_EINVAL (-1)     // bad shape / alignment / null pointer
#define PIVP_ECUDA (-2)      // CUDA runtime error at launch
#define PIVP_EUNSUPPORTED (-3)

namespace pivp {

void set_error(const char* fmt, ...);
void note_launch();          // bumps the process-wide kernel-launch counter (pivp_launch_count)

// 2-D strided view of an NHWC tensor slice: element (row m, channel ch) lives at p[m*cs + co + ch].
struct View {
    float* p;
    int cs;   // row (pixel) stride in elements
    int co;   // channel offset of the slice inside the row
};
struct CView {
    const float* p;
    int cs;
    int co;
};

// ---- Programmatic dependent launch (PDL).  Kernels launched through launch_k() carry the programmatic-stream-serialization
// attribute, and every kernel of this library calls pdl_wait() (= griddepcontrol.wait: returns once the PREVIOUS kernel of the stream
// has completed and its writes are visible) before its first global access, so stream-order semantics are kept while the launch of
// kernel N+1 (CTA scheduling, parameter fetch, and in the tcgen05 kernels barrier init / TMEM allocation) no longer waits for the
// full kernel boundary after kernel N.  Measured on the b32 training step (~810 dependent launches in one CUDA graph): 9.40 ms,
// stable, against 9.48 / 9.78 ms (two modes) with plain launches.  An explicit early trigger (griddepcontrol.launch_dependents at the
// top of every kernel) was measured too and is WORSE (10.03 ms): CTAs of the next kernel become resident on whichever SMs free up
// first and pack depth-first, so multi-CTA-per-SM kernels start unbalanced; a late trigger in the GEMM epilogues gave 9.56 ms.  The
// trigger is therefore left implicit (at kernel completion); an early trigger in the LayerNorm kernels alone (so that the one-CTA-per-SM
// GEMM that follows could run its prologue early) changed nothing (9.30 vs 9.29 ms).  PIVP_PDL=0 or pivp_set_pdl(0) launches plainly.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); }

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at;
    memset(&at, 0, sizeof(at));
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = &at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

static inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PIVP_ECUDA;
    }
    note_launch();
    return PIVP_OK;
}

// cudaFuncSetAttribute (the > 48 KB dynamic shared-memory opt-in) is PER DEVICE: a process that drives several GPUs must apply it on each
// one.  `need()` is true the first time it is asked on the calling thread's current device (and for a larger size than granted before).
struct PerDeviceOnce {
    size_t granted[64] = {};
    bool need(size_t bytes = 1) {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (bytes <= granted[d]) return false;
        granted[d] = bytes;
        return true;
    }
};

#define PIVP_REQUIRE(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            pivp::set_error(__VA_ARGS__);        \
            return PIVP_EINVAL;                  \
        }                                        \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32).  `red` must hold 32 floats.  Result valid in all threads.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect `red` from a previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// accurate variants used by the fp32 parity path
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace pivp
