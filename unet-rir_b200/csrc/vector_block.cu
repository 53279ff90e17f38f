// The embedding-conditioned bottleneck (dl_models/u_net.py:253-263):
// Embedding(2000,256) -> Flatten -> Dense(H5*W5*16) -> Dropout(.3) [-> Reshape -> Conv1x1 -> Add,
// which run through the conv entry points]. The Dense layer is a skinny GEMM (M = batch) whose
// cost is streaming the 8192x1440 kernel once: weight-read bound, CUDA cores with bf16 weights,
// fp32 accumulation, shared-memory staged activations.
#include "urir_common.cuh"

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

__global__ void embedding_bwd_kernel(const int* __restrict__ idx, const float* __restrict__ dx,
                                     float* __restrict__ dtable, long long n_tok, int D, int vocab) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tok * D; i += (long long)gridDim.x * blockDim.x) {
        const long long tok = i / D;
        const int d = (int)(i - tok * D);
        int id = idx[tok];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        atomicAdd(dtable + (size_t)id * D + d, dx[i]);
    }
}

// ---- Dense forward: acc[b][n] += sum_k x[b][k] w[k][n] over a K chunk ---------------------
// block: 64 columns x 4 k-lanes; each thread keeps BT row accumulators; x chunk lives in smem.
constexpr int DENSE_BT = 32;     // batch rows per pass
constexpr int DENSE_KC = 512;    // k per block
__global__ void __launch_bounds__(256)
dense_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                 float* __restrict__ acc_out, int B, int Kd, int N) {
    __shared__ __align__(16) __nv_bfloat16 xs[DENSE_BT][DENSE_KC + 8];
    const int col = blockIdx.x * 64 + (threadIdx.x & 63);
    const int klane = threadIdx.x >> 6;                     // 0..3
    const int k0 = blockIdx.y * DENSE_KC;
    const int kc = (Kd - k0 < DENSE_KC) ? Kd - k0 : DENSE_KC;
    for (int b0 = 0; b0 < B; b0 += DENSE_BT) {
        const int bt = (B - b0 < DENSE_BT) ? B - b0 : DENSE_BT;
        __syncthreads();
        for (int i = threadIdx.x; i < DENSE_BT * (DENSE_KC / 8); i += 256) {
            const int r = i / (DENSE_KC / 8), c8 = (i % (DENSE_KC / 8)) * 8;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (r < bt && c8 < kc) v = __ldg(reinterpret_cast<const uint4*>(x + (size_t)(b0 + r) * Kd + k0 + c8));
            *reinterpret_cast<uint4*>(&xs[r][c8]) = v;
        }
        __syncthreads();
        float acc[DENSE_BT];
#pragma unroll
        for (int b = 0; b < DENSE_BT; ++b) acc[b] = 0.f;
        if (col < N) {
            for (int kk = klane * 8; kk < kc; kk += 32) {
                float wv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) wv[j] = ld_as_f32(w + (size_t)(k0 + kk + j) * N + col);
#pragma unroll
                for (int b = 0; b < DENSE_BT; ++b) {
                    const uint4 u = *reinterpret_cast<const uint4*>(&xs[b][kk]);
                    const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
                    acc[b] = fmaf(a0.x, wv[0], acc[b]); acc[b] = fmaf(a0.y, wv[1], acc[b]);
                    acc[b] = fmaf(a1.x, wv[2], acc[b]); acc[b] = fmaf(a1.y, wv[3], acc[b]);
                    acc[b] = fmaf(a2.x, wv[4], acc[b]); acc[b] = fmaf(a2.y, wv[5], acc[b]);
                    acc[b] = fmaf(a3.x, wv[6], acc[b]); acc[b] = fmaf(a3.y, wv[7], acc[b]);
                }
            }
#pragma unroll
            for (int b = 0; b < DENSE_BT; ++b)
                if (b < bt) atomicAdd(acc_out + (size_t)(b0 + b) * N + col, acc[b]);
        }
    }
}

__global__ void dense_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ bias,
                                      const float* __restrict__ mask, __nv_bfloat16* __restrict__ out, long long n, int N) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = acc[i] + (bias ? bias[i % N] : 0.f);
        if (mask) v *= mask[i];
        out[i] = f2bf(v);
    }
}

// ---- Dense backward -----------------------------------------------------------------------
// dw[k][n] = sum_b x[b][k] * g[b][n] ; block tile 64k x 64n, thread 4x4 ; g = dy*mask staged fp32
__global__ void __launch_bounds__(256)
dense_bwd_w_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                   const float* __restrict__ mask, float* __restrict__ dw, int B, int Kd, int N) {
    __shared__ float xs[32][64 + 4];
    __shared__ float gs[32][64 + 4];
    const int k0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tk = (threadIdx.x >> 4) * 4, tn = (threadIdx.x & 15) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int b0 = 0; b0 < B; b0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * 64; i += 256) {
            const int r = i >> 6, c = i & 63;
            const int b = b0 + r;
            xs[r][c] = (b < B && k0 + c < Kd) ? ld_as_f32(x + (size_t)b * Kd + k0 + c) : 0.f;
            float g = 0.f;
            if (b < B && n0 + c < N) {
                g = ld_as_f32(dy + (size_t)b * N + n0 + c);
                if (mask) g *= mask[(size_t)b * N + n0 + c];
            }
            gs[r][c] = g;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float4 xv = *reinterpret_cast<const float4*>(&xs[r][tk]);
            const float4 gv = *reinterpret_cast<const float4*>(&gs[r][tn]);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ga[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + tk + i;
        if (k >= Kd) continue;
        if (n0 + tn + 3 < N) {
            *reinterpret_cast<float4*>(dw + (size_t)k * N + n0 + tn) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n0 + tn + j < N) dw[(size_t)k * N + n0 + tn + j] = acc[i][j];
        }
    }
}

// db[n] = sum_b g[b][n]
__global__ void dense_bwd_b_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ mask,
                                   float* __restrict__ db, int B, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
        float g = ld_as_f32(dy + (size_t)b * N + n);
        if (mask) g *= mask[(size_t)b * N + n];
        s += g;
    }
    db[n] = s;
}

// dx[b][k] = sum_n g[b][n] * w[k][n] ; block tile: 64 k x 32 b, n chunks of 64 staged in smem
__global__ void __launch_bounds__(256)
dense_bwd_x_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ mask,
                   const __nv_bfloat16* __restrict__ w, float* __restrict__ dx, int B, int Kd, int N) {
    __shared__ float ws[64][64 + 1];
    __shared__ float gs[32][64 + 4];
    const int k0 = blockIdx.x * 64, b0 = blockIdx.y * 32;
    const int kl = threadIdx.x & 63, bg = (threadIdx.x >> 6) * 8;    // 8 batch rows per thread
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int n0 = 0; n0 < N; n0 += 64) {
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 64; i += 256) {
            const int r = i >> 6, c = i & 63;
            ws[r][c] = (k0 + r < Kd && n0 + c < N) ? ld_as_f32(w + (size_t)(k0 + r) * N + n0 + c) : 0.f;
        }
        for (int i = threadIdx.x; i < 32 * 64; i += 256) {
            const int r = i >> 6, c = i & 63;
            const int b = b0 + r;
            float g = 0.f;
            if (b < B && n0 + c < N) {
                g = ld_as_f32(dy + (size_t)b * N + n0 + c);
                if (mask) g *= mask[(size_t)b * N + n0 + c];
            }
            gs[r][c] = g;
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < 64; c += 4) {
            const float w0 = ws[kl][c], w1 = ws[kl][c + 1], w2 = ws[kl][c + 2], w3 = ws[kl][c + 3];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 g = *reinterpret_cast<const float4*>(&gs[bg + i][c]);
                acc[i] = fmaf(g.x, w0, acc[i]); acc[i] = fmaf(g.y, w1, acc[i]);
                acc[i] = fmaf(g.z, w2, acc[i]); acc[i] = fmaf(g.w, w3, acc[i]);
            }
        }
    }
    if (k0 + kl < Kd)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (b0 + bg + i < B) dx[(size_t)(b0 + bg + i) * Kd + k0 + kl] = acc[i];
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
int embedding_bwd(const int* idx, const float* dx, float* dtable, int B, int T, int D, int vocab, cudaStream_t st) {
    URIR_CUDA_OK(cudaMemsetAsync(dtable, 0, sizeof(float) * (size_t)vocab * D, st));
    const long long n = (long long)B * T * D;
    embedding_bwd_kernel<<<cdiv(n, 256), 256, 0, st>>>(idx, dx, dtable, (long long)B * T, D, vocab);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int dense_fwd(const void* x, const void* w, const float* bias, const float* mask, void* out, float* ws,
              int B, int Kd, int N, cudaStream_t st) {
    URIR_CHECK_ARG(Kd % 8 == 0 && ws != nullptr, "dense_fwd: Kd must be a multiple of 8 and ws non-null");
    URIR_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)B * N, st));
    dim3 grid(cdiv(N, 64), cdiv(Kd, DENSE_KC));
    dense_fwd_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, ws, B, Kd, N);
    URIR_LAUNCH_OK(0);
    const long long n = (long long)B * N;
    dense_finalize_kernel<<<cdiv(n, 256), 256, 0, st>>>(ws, bias, mask, (__nv_bfloat16*)out, n, N);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int dense_bwd(const void* x, const void* w, const void* dy, const float* mask, float* dw, float* db, float* dx,
              int B, int Kd, int N, cudaStream_t st) {
    if (dw) {
        dim3 grid(cdiv(N, 64), cdiv(Kd, 64));
        dense_bwd_w_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, mask, dw, B, Kd, N);
        URIR_LAUNCH_OK(0);
    }
    if (db) {
        dense_bwd_b_kernel<<<cdiv(N, 128), 128, 0, st>>>((const __nv_bfloat16*)dy, mask, db, B, N);
        URIR_LAUNCH_OK(0);
    }
    if (dx) {
        dim3 grid(cdiv(Kd, 64), cdiv(B, 32));
        dense_bwd_x_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, mask, (const __nv_bfloat16*)w, dx, B, Kd, N);
        URIR_LAUNCH_OK(0);
    }
    return URIR_OK;
}
int dropout_mask(float* mask, long long n, float rate, uint64_t seed, const int* step_dev, cudaStream_t st) {
    URIR_CHECK_ARG(rate >= 0.f && rate < 1.f, "dropout: rate must be in [0,1)");
    dropout_mask_kernel<<<cdiv(n, 256), 256, 0, st>>>(mask, n, rate, seed, step_dev);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
