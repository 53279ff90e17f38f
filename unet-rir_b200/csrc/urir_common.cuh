// Shared helpers for liburir (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/urir.h"

namespace urir {

// ---- error plumbing: thread-local message + monotonically increasing launch counter ------
void set_error(const char* fmt, ...);
int  fail(int code, const char* fmt, ...);
void count_launch(int kind);   // kind: 0 = SIMT/bandwidth kernel, 1 = tcgen05 kernel

#define URIR_CHECK_ARG(cond, ...) do { if (!(cond)) return ::urir::fail(URIR_ERR_ARG, __VA_ARGS__); } while (0)

#define URIR_CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) \
    return ::urir::fail(URIR_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); } while (0)

#define URIR_LAUNCH_OK(kind) do { cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) \
    return ::urir::fail(URIR_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__, __LINE__, \
                        cudaGetErrorString(_e)); ::urir::count_launch(kind); } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
// SMs of the current device (cudaDevAttrMultiProcessorCount, cached; 148 on B200): persistent grids and the
// dispatch gates ("enough tiles per SM") are sized from it, never from a literal
int sm_count();

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// Kernels launched through launch_pdl() may become resident while their predecessor in the stream is still
// running: their prologue (barrier init, TMEM allocation, tensor-map prefetch, CTA scheduling) overlaps the
// predecessor's tail. Such a kernel MUST execute pdl_wait() before it touches global memory a predecessor may
// write (and before it writes anything); pdl_trigger() lets ITS successor start early in turn. Completion
// order stays transitive because every kernel waits before it finishes. Captured into CUDA graphs as
// programmatic dependency edges. URIR_NO_PDL=1 turns the attribute off (plain stream order).
bool pdl_enabled();
template <typename... KA, typename... A>
static inline cudaError_t launch_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}

// ---- deterministic mode (URIR_DETERMINISTIC=1 / urir_set_deterministic) -----------------------
// Every cross-CTA floating-point reduction of the library ends in atomicAdd / red.global from each CTA, whose
// arrival order -- and therefore the rounding of the sum -- changes from run to run. In deterministic mode each
// such kernel gets a "gate": a device counter through which its CTAs commit their partial results in blockIdx
// order (CTA i spins until the counter reads i, adds, publishes i + 1; the last CTA resets it to 0 so CUDA-graph
// replays can reuse it). Lower-numbered CTAs are dispatched first and never wait on higher ones, so the chain
// cannot deadlock; a bounded spin traps instead of hanging if that assumption were ever violated.
// next_gate() hands out the slots round-robin (host side, nullptr when the mode is off): kernels that may run
// concurrently on two streams of one step get distinct slots.
unsigned int* next_gate();
bool deterministic();

// ---- device helpers ---------------------------------------------------------------------
// gate_enter / gate_leave bracket the global atomics of a CTA. `nthreads` threads (all of which call both, with the
// same arguments) synchronise on named barrier `bar_id` (0 = the whole CTA's __syncthreads barrier when nthreads ==
// blockDim.x); `leader` is true for exactly one of them.
__device__ __forceinline__ void gate_bar(int bar_id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void gate_enter(unsigned int* gate, unsigned int turn, bool leader, int bar_id, int nthreads) {
    if (gate == nullptr) return;
    if (leader) {
        unsigned int v;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gate) : "memory");
            if (v == turn) break;
            if (clock64() - t0 > (1ll << 31)) asm volatile("trap;");
        }
    }
    gate_bar(bar_id, nthreads);
}
__device__ __forceinline__ void gate_leave(unsigned int* gate, unsigned int turn, unsigned int total, bool leader, int bar_id,
                                           int nthreads) {
    if (gate == nullptr) return;
    __threadfence();
    gate_bar(bar_id, nthreads);
    if (leader) {
        const unsigned int nxt = (turn + 1 == total) ? 0u : turn + 1;
        asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(gate), "r"(nxt) : "memory");
    }
}
__device__ __forceinline__ unsigned int cta_linear() { return blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); }
__device__ __forceinline__ unsigned int cta_count() { return gridDim.x * gridDim.y * gridDim.z; }

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }
__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    float2 r;
    r.x = __uint_as_float(u << 16);
    r.y = __uint_as_float(u & 0xffff0000u);
    return r;
}

template <typename T> __device__ __forceinline__ float ld_as_f32(const T* p);
template <> __device__ __forceinline__ float ld_as_f32<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_f32<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p))) << 16);
}
template <typename T> __device__ __forceinline__ void st_from_f32(T* p, float v);
template <> __device__ __forceinline__ void st_from_f32<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_f32<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = f2bf(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit streaming loads / stores (guideline 13: bypass L1 for single-use data)
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_na_v4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace urir
