// The stem convolution enc1.down (dl_models/u_net.py:269-276 applied to the 2-channel fp32
// spectrogram): Cin = 2 makes it a K_gemm = 18 (k=3) or 72 (k=6) problem -- bandwidth-bound, not
// GEMM-shaped, so it runs on CUDA cores with everything staged on chip:
//   fprop : one thread per output pixel, all 32 output channels in registers, weights as fp32 in
//           shared memory (128-bit broadcast reads), 64-byte contiguous bf16 store per thread.
//   wgrad : one warp per pixel lane (lane = output channel), the 18/72 (tap, cin) inputs are
//           warp-uniform loads, 18/72 register accumulators, block reduction, one atomic per output.
#include "urir_common.cuh"

namespace urir {

struct StemP {
    int N, H, W, C, K, R, S, stride, pt, pl, P, Q;
    int x_ld, x_coff, y_ld, y_coff;
};

constexpr int STEM_KT = 32;

__global__ void __launch_bounds__(128)
stem_fprop_kernel(StemP p, const float* __restrict__ x, const __nv_bfloat16* __restrict__ w_ck,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ y) {
    extern __shared__ __align__(16) float ws[];            // [R*S*C][32] then bias[32]
    const int J = p.R * p.S * p.C;
    const int k0 = blockIdx.y * STEM_KT;
    for (int i = threadIdx.x; i < J * STEM_KT; i += blockDim.x) {
        const int j = i / STEM_KT, kk = i % STEM_KT;
        ws[i] = bf2f(w_ck[(size_t)j * p.K + k0 + kk]);     // w_ck = [tap][c][k] = [j][k]
    }
    if (threadIdx.x < STEM_KT) ws[J * STEM_KT + threadIdx.x] = bias ? bias[k0 + threadIdx.x] : 0.f;
    __syncthreads();
    const long long M = (long long)p.N * p.P * p.Q;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int q = (int)(m % p.Q), pp = (int)((m / p.Q) % p.P), n = (int)(m / ((long long)p.P * p.Q));
    float acc[STEM_KT];
#pragma unroll
    for (int i = 0; i < STEM_KT; ++i) acc[i] = ws[J * STEM_KT + i];
    for (int r = 0; r < p.R; ++r) {
        const int ih = pp * p.stride + r - p.pt;
        if (ih < 0 || ih >= p.H) continue;
        for (int s = 0; s < p.S; ++s) {
            const int iw = q * p.stride + s - p.pl;
            if (iw < 0 || iw >= p.W) continue;
            const float* xp = x + ((size_t)(n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff;
            for (int c = 0; c < p.C; ++c) {
                const float xv = __ldg(xp + c);
                const float4* wr = reinterpret_cast<const float4*>(ws + ((r * p.S + s) * p.C + c) * STEM_KT);
#pragma unroll
                for (int i = 0; i < STEM_KT / 4; ++i) {
                    const float4 w4 = wr[i];
                    acc[4 * i] = fmaf(xv, w4.x, acc[4 * i]); acc[4 * i + 1] = fmaf(xv, w4.y, acc[4 * i + 1]);
                    acc[4 * i + 2] = fmaf(xv, w4.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(xv, w4.w, acc[4 * i + 3]);
                }
            }
        }
    }
    __nv_bfloat16* yp = y + (size_t)m * p.y_ld + p.y_coff + k0;
#pragma unroll
    for (int i = 0; i < STEM_KT / 8; ++i)
        *reinterpret_cast<uint4*>(yp + 8 * i) =
            make_uint4(pack_bf16x2(acc[8 * i], acc[8 * i + 1]), pack_bf16x2(acc[8 * i + 2], acc[8 * i + 3]),
                       pack_bf16x2(acc[8 * i + 4], acc[8 * i + 5]), pack_bf16x2(acc[8 * i + 6], acc[8 * i + 7]));
}

// dw[j = (tap, c)][k]: warp w of a block walks pixels m0 + w, m0 + w + 8, ...; lane = k.
template <int R>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(StemP p, const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                  float* __restrict__ dw, int pix_per_block) {
    constexpr int J = R * R * 2;
    __shared__ float red[J][STEM_KT + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.y * STEM_KT + lane;
    for (int i = threadIdx.x; i < J * (STEM_KT + 1); i += blockDim.x) (&red[0][0])[i] = 0.f;
    __syncthreads();
    const long long M = (long long)p.N * p.P * p.Q;
    const long long m0 = (long long)blockIdx.x * pix_per_block;
    const long long m1 = (m0 + pix_per_block < M) ? m0 + pix_per_block : M;
    float acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = 0.f;
    for (long long m = m0 + warp; m < m1; m += 8) {
        const int q = (int)(m % p.Q), pp = (int)((m / p.Q) % p.P), n = (int)(m / ((long long)p.P * p.Q));
        const float g = ld_as_f32(dy + (size_t)m * p.y_ld + p.y_coff + k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int ih = pp * p.stride + r - p.pt;
#pragma unroll
            for (int s = 0; s < R; ++s) {
                const int iw = q * p.stride + s - p.pl;
                const bool in = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
                const float2 xv = in ? __ldg(reinterpret_cast<const float2*>(
                                           x + ((size_t)(n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff))
                                     : make_float2(0.f, 0.f);
                acc[(r * R + s) * 2] = fmaf(xv.x, g, acc[(r * R + s) * 2]);
                acc[(r * R + s) * 2 + 1] = fmaf(xv.y, g, acc[(r * R + s) * 2 + 1]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) atomicAdd(&red[j][lane], acc[j]);
    __syncthreads();
    for (int i = threadIdx.x; i < J * STEM_KT; i += blockDim.x) {
        const int j = i / STEM_KT, kk = i % STEM_KT;
        atomicAdd(dw + (size_t)j * p.K + blockIdx.y * STEM_KT + kk, red[j][kk]);
    }
}

static StemP to_stem(const urir_conv_desc* d) {
    StemP p;
    p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C; p.K = d->K; p.R = d->R; p.S = d->S; p.stride = d->stride;
    p.pt = d->pad_top; p.pl = d->pad_left; p.P = d->P; p.Q = d->Q;
    p.x_ld = d->x_ld; p.x_coff = d->x_coff; p.y_ld = d->y_ld; p.y_coff = d->y_coff;
    return p;
}

bool stem_fprop_supported(const urir_conv_desc* d, const float* stats) {
    return d->x_dtype == URIR_F32 && d->y_dtype == URIR_BF16 && d->C <= 4 && d->K % STEM_KT == 0 && stats == nullptr &&
           d->act == URIR_ACT_NONE && !d->accumulate && d->y_ld % 8 == 0 && d->y_coff % 8 == 0 && d->R * d->S * d->C <= 144;
}
bool stem_wgrad_supported(const urir_conv_desc* d) {
    const int J = d->R * d->S * d->C;
    return d->x_dtype == URIR_F32 && d->y_dtype == URIR_BF16 && d->C == 2 && d->K % STEM_KT == 0 && d->R == d->S &&
           (J == 18 || J == 72) && d->x_ld % 2 == 0 && d->x_coff % 2 == 0;
}

int stem_fprop(const urir_conv_desc* d, const void* x, const void* w_ck, const float* bias, void* y, cudaStream_t st) {
    StemP p = to_stem(d);
    const long long M = (long long)p.N * p.P * p.Q;
    const int J = p.R * p.S * p.C;
    dim3 grid(cdiv(M, 128), p.K / STEM_KT);
    stem_fprop_kernel<<<grid, 128, (J * STEM_KT + STEM_KT) * sizeof(float), st>>>(
        p, (const float*)x, (const __nv_bfloat16*)w_ck, bias, (__nv_bfloat16*)y);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int stem_wgrad(const urir_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    StemP p = to_stem(d);
    const long long M = (long long)p.N * p.P * p.Q;
    const int J = p.R * p.S * p.C;
    URIR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)J * p.K, st));
    long long chunks = 148LL * 8 / (p.K / STEM_KT);
    long long pix = (M + chunks - 1) / chunks;
    if (pix < 64) pix = 64;
    dim3 grid(cdiv(M, pix), p.K / STEM_KT);
    if (J == 18) stem_wgrad_kernel<3><<<grid, 256, 0, st>>>(p, (const float*)x, (const __nv_bfloat16*)dy, dw, (int)pix);
    else stem_wgrad_kernel<6><<<grid, 256, 0, st>>>(p, (const float*)x, (const __nv_bfloat16*)dy, dw, (int)pix);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
