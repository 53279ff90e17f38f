// The embedding-conditioned bottleneck (dl_models/u_net.py:253-263):
// Embedding(2000,256) -> Flatten -> Dense(H5*W5*16) -> Dropout(.3) [-> Reshape -> Conv1x1 -> Add,
// which run through the conv entry points]. The Dense layer is a skinny GEMM (M = batch) whose
// cost is streaming the 8192x1440 kernel once: weight-read bound, CUDA cores with bf16 weights,
// fp32 accumulation, shared-memory staged activations.
#include "urir_common.cuh"
#include "../../include/urir.h"

namespace urir {

// ---- Embedding ---------------------------------------------------------------------------
__global__ void embedding_fwd_kernel(const int* __restrict__ idx, const float* __restrict__ table,
                                     __nv_bfloat16* __restrict__ out, long long n_tok, int D, int vocab) {
    const int D4 = D >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tok * D4; i += (long long)gridDim.x * blockDim.x) {
        const long long tok = i / D4;
        const int d = (int)(i - tok * D4) << 2;
        int id = idx[tok];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        const float4 v = __ldg(reinterpret_cast<const float4*>(table + (size_t)id * D + d));
        uint2 o; o.x = pack_bf16x2(v.x, v.y); o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(out + tok * D + d) = o;
    }
}

template <typename T>
__global__ void embedding_bwd_kernel(const int* __restrict__ idx, const T* __restrict__ dx,
                                     float* __restrict__ dtable, long long n_tok, int D, int vocab) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tok * D; i += (long long)gridDim.x * blockDim.x) {
        const long long tok = i / D;
        const int d = (int)(i - tok * D);
        int id = idx[tok];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        atomicAdd(dtable + (size_t)id * D + d, ld_as_f32(dx + i));
    }
}

// deterministic mode: one block per vocabulary row scans the token list in order (no atomics; rows nobody indexes
// just get their zero) -- n_tok is B * 32, so the scan is tiny
template <typename T>
__global__ void embedding_bwd_ordered_kernel(const int* __restrict__ idx, const T* __restrict__ dx,
                                             float* __restrict__ dtable, long long n_tok, int D, int vocab) {
    const int row = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (long long tok = 0; tok < n_tok; ++tok) {
            int id = idx[tok];
            id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
            if (id == row) acc += ld_as_f32(dx + tok * D + d);
        }
        dtable[(size_t)row * D + d] = acc;
    }
}

// ---- Dense + Dropout on the tensor cores -----------------------------------------------------
// Dense(8192 -> 1440) over a batch is a 1x1 convolution over B "pixels": forward = conv fprop (A = x K-major,
// B = w^T [N][Kd] K-major), dx = conv dgrad (B = w [Kd][N]), dw = conv wgrad (both operands MN-major, the
// reduction runs over the batch). All three go through the tcgen05 implicit-GEMM entry points, so the
// 23.6 MB bf16 kernel is streamed by TMA once per pass instead of through CUDA-core FMAs.
__global__ void mul_mask_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mask,
                                     __nv_bfloat16* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = f2bf(ld_as_f32(x + i) * mask[i]);
}

static urir_conv_desc dense_desc(int B, int Kd, int N) {
    urir_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.N = B; d.H = 1; d.W = 1; d.C = Kd; d.K = N; d.R = 1; d.S = 1; d.stride = 1; d.P = 1; d.Q = 1;
    d.x_ld = Kd; d.y_ld = N; d.x_dtype = URIR_BF16; d.y_dtype = URIR_BF16; d.impl = URIR_IMPL_AUTO;
    return d;
}

// ---- Dropout mask: counter-based (splitmix64 of seed, step, index) -----------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void dropout_mask_kernel(float* __restrict__ mask, long long n, float rate, uint64_t seed,
                                    const int* __restrict__ step_dev) {
    const uint64_t step = step_dev ? (uint64_t)(uint32_t)*step_dev : 0ull;
    const float keep_scale = 1.f / (1.f - rate);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint64_t h = splitmix64(splitmix64(seed ^ (step * 0xD1342543DE82EF95ull)) + (uint64_t)i);
        const float u = (float)(h >> 40) * (1.0f / 16777216.0f);
        mask[i] = (u >= rate) ? keep_scale : 0.f;
    }
}

// ---- host ------------------------------------------------------------------------------------
int embedding_fwd(const int* idx, const float* table, void* out, int B, int T, int D, int vocab, cudaStream_t st) {
    URIR_CHECK_ARG(D % 4 == 0, "embedding: D must be a multiple of 4");
    const long long n = (long long)B * T * (D / 4);
    embedding_fwd_kernel<<<cdiv(n, 256), 256, 0, st>>>(idx, table, (__nv_bfloat16*)out, (long long)B * T, D, vocab);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int embedding_bwd(const int* idx, const void* dx, int dx_dtype, float* dtable, int B, int T, int D, int vocab, cudaStream_t st) {
    if (deterministic()) {
        if (dx_dtype == URIR_BF16) embedding_bwd_ordered_kernel<__nv_bfloat16><<<vocab, 256, 0, st>>>(idx, (const __nv_bfloat16*)dx, dtable, (long long)B * T, D, vocab);
        else embedding_bwd_ordered_kernel<float><<<vocab, 256, 0, st>>>(idx, (const float*)dx, dtable, (long long)B * T, D, vocab);
        URIR_LAUNCH_OK(0);
        return URIR_OK;
    }
    URIR_CUDA_OK(cudaMemsetAsync(dtable, 0, sizeof(float) * (size_t)vocab * D, st));
    const long long n = (long long)B * T * D;
    if (dx_dtype == URIR_BF16) embedding_bwd_kernel<__nv_bfloat16><<<cdiv(n, 256), 256, 0, st>>>(idx, (const __nv_bfloat16*)dx, dtable, (long long)B * T, D, vocab);
    else embedding_bwd_kernel<float><<<cdiv(n, 256), 256, 0, st>>>(idx, (const float*)dx, dtable, (long long)B * T, D, vocab);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int dense_fwd(const void* x, const void* w_kn, const void* w_nk, const float* bias, const float* mask, void* out,
              int B, int Kd, int N, cudaStream_t st) {
    const urir_conv_desc d = dense_desc(B, Kd, N);
    int rc = urir_conv2d_fprop(&d, x, w_kn, w_nk, bias, out, nullptr, (void*)st);
    if (rc) return rc;
    if (mask) {
        const long long n = (long long)B * N;
        mul_mask_bf16_kernel<<<cdiv(n, 256), 256, 0, st>>>((const __nv_bfloat16*)out, mask, (__nv_bfloat16*)out, n);
        URIR_LAUNCH_OK(0);
    }
    return URIR_OK;
}
int dense_bwd(const void* x, const void* w_kn, const void* w_nk, const void* dy, const float* mask, void* dy_eff,
              float* dw, float* db, void* dx, int B, int Kd, int N, cudaStream_t st) {
    const urir_conv_desc d = dense_desc(B, Kd, N);
    const void* g = dy;
    if (mask) {
        URIR_CHECK_ARG(dy_eff != nullptr, "dense_bwd: a mask needs the dy_eff scratch buffer");
        const long long n = (long long)B * N;
        mul_mask_bf16_kernel<<<cdiv(n, 256), 256, 0, st>>>((const __nv_bfloat16*)dy, mask, (__nv_bfloat16*)dy_eff, n);
        URIR_LAUNCH_OK(0);
        g = dy_eff;
    }
    int rc;
    if (dw && (rc = urir_conv2d_wgrad(&d, x, g, dw, (void*)st))) return rc;
    if (db && (rc = urir_channel_sum(g, URIR_BF16, B, N, N, 0, db, (void*)st))) return rc;
    if (dx && (rc = urir_conv2d_dgrad(&d, g, w_kn, w_nk, nullptr, dx, nullptr, (void*)st))) return rc;
    return URIR_OK;
}
int dropout_mask(float* mask, long long n, float rate, uint64_t seed, const int* step_dev, cudaStream_t st) {
    URIR_CHECK_ARG(rate >= 0.f && rate < 1.f, "dropout: rate must be in [0,1)");
    dropout_mask_kernel<<<cdiv(n, 256), 256, 0, st>>>(mask, n, rate, seed, step_dev);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
