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

// Input-gradient of the head conv (K = 2 output channels, u_net.py:248): dx[p][c] = sum over (tap, k) of
// dy[p - tap + pad][k] * w[tap][c][k] -- 72 (tap, k) terms per output element, again far too thin for a
// GEMM. Block = 8 x 32 pixels; the fp32 dy halo tile (one plane per k, conflict-free) and the weights
// (fp32, [tap*K + k][32 c]) live in shared memory; a thread owns one pixel x 32 channels.
constexpr int HD_TH = 8, HD_TW = 32;
__global__ void __launch_bounds__(256)
head_dgrad_kernel(StemP p, const float* __restrict__ dy, const __nv_bfloat16* __restrict__ w_ck,
                  __nv_bfloat16* __restrict__ dx, int accumulate) {
    extern __shared__ __align__(16) float sm[];
    const int J = p.R * p.S * p.K;                                    // (tap, k) terms
    const int hh = HD_TH + p.R - 1, hw = HD_TW + p.S - 1;             // halo tile extent
    float* ws = sm;                                                   // [J][32]
    float* gs = sm + J * 32;                                          // [K][hh][hw]
    const int c0 = blockIdx.z * 32;
    int t = blockIdx.x;
    const int tw = t % ((p.W + HD_TW - 1) / HD_TW); t /= ((p.W + HD_TW - 1) / HD_TW);
    const int th = t % ((p.H + HD_TH - 1) / HD_TH);
    const int n = t / ((p.H + HD_TH - 1) / HD_TH);
    const int h0 = th * HD_TH, w0 = tw * HD_TW;
    for (int i = threadIdx.x; i < J * 32; i += blockDim.x) {
        const int j = i >> 5, cc = i & 31;
        const int tap = j / p.K, k = j % p.K;
        ws[i] = bf2f(w_ck[((size_t)tap * p.C + c0 + cc) * p.K + k]);
    }
    // dy rows needed: oh = h + pt - r for h in [h0, h0+TH), r in [0,R)  ->  [h0 + pt - (R-1), h0 + TH-1 + pt]
    const int oh0 = h0 + p.pt - (p.R - 1), ow0 = w0 + p.pl - (p.S - 1);
    for (int i = threadIdx.x; i < p.K * hh * hw; i += blockDim.x) {
        const int k = i / (hh * hw), rem = i % (hh * hw);
        const int oh = oh0 + rem / hw, ow = ow0 + rem % hw;
        float v = 0.f;
        if (oh >= 0 && oh < p.P && ow >= 0 && ow < p.Q)
            v = __ldg(dy + ((size_t)(n * p.P + oh) * p.Q + ow) * p.y_ld + p.y_coff + k);
        gs[i] = v;
    }
    __syncthreads();
    const int ly = threadIdx.x / HD_TW, lx = threadIdx.x % HD_TW;
    const int h = h0 + ly, w = w0 + lx;
    if (h >= p.H || w >= p.W) return;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int r = 0; r < p.R; ++r)
        for (int s = 0; s < p.S; ++s) {
            // oh = h + pt - r  ->  local row = oh - oh0 = ly + (R-1) - r ; same for columns
            const int gy = ly + (p.R - 1) - r, gx = lx + (p.S - 1) - s;
            for (int k = 0; k < p.K; ++k) {
                const float g = gs[(k * hh + gy) * hw + gx];
                const float4* wr = reinterpret_cast<const float4*>(ws + ((r * p.S + s) * p.K + k) * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w4 = wr[i];
                    acc[4 * i] = fmaf(g, w4.x, acc[4 * i]); acc[4 * i + 1] = fmaf(g, w4.y, acc[4 * i + 1]);
                    acc[4 * i + 2] = fmaf(g, w4.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(g, w4.w, acc[4 * i + 3]);
                }
            }
        }
    __nv_bfloat16* xp = dx + ((size_t)(n * p.H + h) * p.W + w) * p.x_ld + p.x_coff + c0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 o;
        if (accumulate) {
            const uint4 e = *reinterpret_cast<const uint4*>(xp + 8 * i);
            const float2 e0 = unpack_bf16x2(e.x), e1 = unpack_bf16x2(e.y), e2 = unpack_bf16x2(e.z), e3 = unpack_bf16x2(e.w);
            o = make_uint4(pack_bf16x2(acc[8 * i] + e0.x, acc[8 * i + 1] + e0.y), pack_bf16x2(acc[8 * i + 2] + e1.x, acc[8 * i + 3] + e1.y),
                           pack_bf16x2(acc[8 * i + 4] + e2.x, acc[8 * i + 5] + e2.y), pack_bf16x2(acc[8 * i + 6] + e3.x, acc[8 * i + 7] + e3.y));
        } else {
            o = make_uint4(pack_bf16x2(acc[8 * i], acc[8 * i + 1]), pack_bf16x2(acc[8 * i + 2], acc[8 * i + 3]),
                           pack_bf16x2(acc[8 * i + 4], acc[8 * i + 5]), pack_bf16x2(acc[8 * i + 6], acc[8 * i + 7]));
        }
        *reinterpret_cast<uint4*>(xp + 8 * i) = o;
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

bool head_dgrad_supported(const urir_conv_desc* d, const float* bias, const float* stats) {
    return d->y_dtype == URIR_F32 && d->x_dtype == URIR_BF16 && d->K <= 4 && d->C % 32 == 0 && d->stride == 1 &&
           d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && bias == nullptr && stats == nullptr && d->R * d->S * d->K <= 160;
}

int head_dgrad(const urir_conv_desc* d, const void* dy, const void* w_ck, void* dx, cudaStream_t st) {
    StemP p = to_stem(d);
    const int J = p.R * p.S * p.K;
    const int hh = HD_TH + p.R - 1, hw = HD_TW + p.S - 1;
    const size_t smem = (size_t)(J * 32 + p.K * hh * hw) * sizeof(float);
    dim3 grid(cdiv(p.W, HD_TW) * cdiv(p.H, HD_TH) * p.N, 1, p.C / 32);
    head_dgrad_kernel<<<grid, 256, smem, st>>>(p, (const float*)dy, (const __nv_bfloat16*)w_ck, (__nv_bfloat16*)dx,
                                              d->accumulate);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
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
    if (!d->accumulate) URIR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)J * p.K, st));
    long long chunks = (long long)sm_count() * 8 / (p.K / STEM_KT);
    long long pix = (M + chunks - 1) / chunks;
    if (pix < 64) pix = 64;
    dim3 grid(cdiv(M, pix), p.K / STEM_KT);
    if (J == 18) stem_wgrad_kernel<3><<<grid, 256, 0, st>>>(p, (const float*)x, (const __nv_bfloat16*)dy, dw, (int)pix);
    else stem_wgrad_kernel<6><<<grid, 256, 0, st>>>(p, (const float*)x, (const __nv_bfloat16*)dy, dw, (int)pix);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
