// CUDA-core direct convolutions: fprop / dgrad / wgrad for any (R,S,stride,C,K).
//
// Role in the design (DESIGN.md "kernels"):
//   * the stem (C=2, K_gemm=18) and the head (K=2) are bandwidth-bound, not GEMM-shaped, and
//     run here by design;
//   * every other shape runs on the tcgen05 implicit-GEMM kernels (conv_igemm.cu /
//     conv_wgrad_tc.cu); this file is then the on-device cross-check (URIR_IMPL_SIMT) that the
//     parity tests use to localise a tensor-core bug to a layer.
// Inputs are bf16 or fp32, weights bf16, accumulation fp32 -- the same arithmetic contract as
// the tensor-core path, so the two agree to summation-order noise.
#include "urir_common.cuh"

namespace urir {

struct ConvP {
    int N, H, W, C, K, R, S, stride, pt, pl, P, Q;
    int x_ld, x_coff, y_ld, y_coff, act, accumulate;
};

static ConvP to_p(const urir_conv_desc* d) {
    ConvP p;
    p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C; p.K = d->K; p.R = d->R; p.S = d->S;
    p.stride = d->stride; p.pt = d->pad_top; p.pl = d->pad_left; p.P = d->P; p.Q = d->Q;
    p.x_ld = d->x_ld; p.x_coff = d->x_coff; p.y_ld = d->y_ld; p.y_coff = d->y_coff;
    p.act = d->act; p.accumulate = d->accumulate;
    return p;
}

template <int NB>
__device__ __forceinline__ void load_w(const __nv_bfloat16* wp, float (&w)[NB]) {
    if constexpr (NB % 8 == 0) {
#pragma unroll
        for (int i = 0; i < NB / 8; ++i) {
            uint4 u = __ldg(reinterpret_cast<const uint4*>(wp) + i);
            float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
            w[i * 8 + 0] = a.x; w[i * 8 + 1] = a.y; w[i * 8 + 2] = b.x; w[i * 8 + 3] = b.y;
            w[i * 8 + 4] = c.x; w[i * 8 + 5] = c.y; w[i * 8 + 6] = e.x; w[i * 8 + 7] = e.y;
        }
    } else if constexpr (NB % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NB / 2; ++i) {
            float2 a = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(wp) + i));
            w[2 * i] = a.x; w[2 * i + 1] = a.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NB; ++i) w[i] = ld_as_f32(wp + i);
    }
}

// out[pixel, k0..k0+KB) ; one thread per (output pixel, KB-channel group)
template <typename TX, typename TY, int KB>
__global__ void __launch_bounds__(128)
conv_fprop_simt(ConvP p, const TX* __restrict__ x, const __nv_bfloat16* __restrict__ w_ck,
                const float* __restrict__ bias, TY* __restrict__ y, float* __restrict__ stats) {
    const long long M = (long long)p.N * p.P * p.Q;
    const long long m = (long long)blockIdx.x * 128 + threadIdx.x;
    const int k0 = blockIdx.y * KB;
    const bool live = m < M;
    float acc[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) acc[i] = 0.f;
    if (live) {
        const int q = (int)(m % p.Q), pp = (int)((m / p.Q) % p.P), n = (int)(m / ((long long)p.P * p.Q));
        for (int r = 0; r < p.R; ++r) {
            const int ih = pp * p.stride + r - p.pt;
            if (ih < 0 || ih >= p.H) continue;
            for (int s = 0; s < p.S; ++s) {
                const int iw = q * p.stride + s - p.pl;
                if (iw < 0 || iw >= p.W) continue;
                const TX* xp = x + ((size_t)(n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff;
                const __nv_bfloat16* wp = w_ck + (size_t)(r * p.S + s) * p.C * p.K + k0;
                for (int c = 0; c < p.C; ++c) {
                    const float xv = ld_as_f32(xp + c);
                    float w[KB];
                    load_w<KB>(wp + (size_t)c * p.K, w);
#pragma unroll
                    for (int i = 0; i < KB; ++i) acc[i] = fmaf(xv, w[i], acc[i]);
                }
            }
        }
    }
    TY* yp = live ? y + (size_t)m * p.y_ld + p.y_coff + k0 : nullptr;
#pragma unroll
    for (int i = 0; i < KB; ++i) {
        float v = acc[i] + (bias ? __ldg(bias + k0 + i) : 0.f);
        if (stats) {   // pre-activation, pre-rounding sums for the BatchNorm that follows
            float s1 = warp_sum(live ? v : 0.f), s2 = warp_sum(live ? v * v : 0.f);
            if ((threadIdx.x & 31) == 0) { atomicAdd(stats + k0 + i, s1); atomicAdd(stats + p.K + k0 + i, s2); }
        }
        if (live) {
            if (p.act == URIR_ACT_SIGMOID) v = 1.f / (1.f + __expf(-v));
            if (p.act == URIR_ACT_RELU) v = fmaxf(v, 0.f);
            if (p.accumulate) v += ld_as_f32(yp + i);
            st_from_f32(yp + i, v);
        }
    }
}

// dx[pixel, c0..c0+CB) ; one thread per (input pixel, CB-channel group)
template <typename TDY, typename TDX, int CB>
__global__ void __launch_bounds__(128)
conv_dgrad_simt(ConvP p, const TDY* __restrict__ dy, const __nv_bfloat16* __restrict__ w_kc,
                const float* __restrict__ bias, TDX* __restrict__ dx, float* __restrict__ stats) {
    const long long M = (long long)p.N * p.H * p.W;
    const long long m = (long long)blockIdx.x * 128 + threadIdx.x;
    const int c0 = blockIdx.y * CB;
    const bool live = m < M;
    float acc[CB];
#pragma unroll
    for (int i = 0; i < CB; ++i) acc[i] = 0.f;
    if (live) {
        const int wq = (int)(m % p.W), h = (int)((m / p.W) % p.H), n = (int)(m / ((long long)p.H * p.W));
        for (int r = 0; r < p.R; ++r) {
            const int th = h + p.pt - r;
            if (th < 0 || (th % p.stride) != 0) continue;
            const int oh = th / p.stride;
            if (oh >= p.P) continue;
            for (int s = 0; s < p.S; ++s) {
                const int tw = wq + p.pl - s;
                if (tw < 0 || (tw % p.stride) != 0) continue;
                const int ow = tw / p.stride;
                if (ow >= p.Q) continue;
                const TDY* yp = dy + ((size_t)(n * p.P + oh) * p.Q + ow) * p.y_ld + p.y_coff;
                const __nv_bfloat16* wp = w_kc + (size_t)(r * p.S + s) * p.K * p.C + c0;
                for (int k = 0; k < p.K; ++k) {
                    const float gv = ld_as_f32(yp + k);
                    float w[CB];
                    load_w<CB>(wp + (size_t)k * p.C, w);
#pragma unroll
                    for (int i = 0; i < CB; ++i) acc[i] = fmaf(gv, w[i], acc[i]);
                }
            }
        }
    }
    TDX* xp = live ? dx + (size_t)m * p.x_ld + p.x_coff + c0 : nullptr;
#pragma unroll
    for (int i = 0; i < CB; ++i) {
        float v = acc[i] + (bias ? __ldg(bias + c0 + i) : 0.f);
        if (stats) {
            float s1 = warp_sum(live ? v : 0.f), s2 = warp_sum(live ? v * v : 0.f);
            if ((threadIdx.x & 31) == 0) { atomicAdd(stats + c0 + i, s1); atomicAdd(stats + p.C + c0 + i, s2); }
        }
        if (live) {
            if (p.accumulate) v += ld_as_f32(xp + i);
            st_from_f32(xp + i, v);
        }
    }
}

// dw[tap][c][k] += sum over a chunk of output pixels. Block = CT x KT threads, one (c,k) each.
template <typename TX, typename TDY>
__global__ void __launch_bounds__(256)
conv_wgrad_simt(ConvP p, const TX* __restrict__ x, const TDY* __restrict__ dy,
                float* __restrict__ dw, int KT, int CT, int pix_per_block) {
    const int kt_count = (p.K + KT - 1) / KT, ct_count = (p.C + CT - 1) / CT;
    int b = blockIdx.x;
    const int kt = b % kt_count; b /= kt_count;
    const int ct = b % ct_count; b /= ct_count;
    const int tap = b;
    const int r = tap / p.S, s = tap % p.S;
    const int k = kt * KT + threadIdx.x % KT;
    const int c = ct * CT + threadIdx.x / KT;
    const bool live = (k < p.K) && (c < p.C) && (threadIdx.x < KT * CT);
    const long long M = (long long)p.N * p.P * p.Q;
    const long long m0 = (long long)blockIdx.y * pix_per_block;
    const long long m1 = (m0 + pix_per_block < M) ? m0 + pix_per_block : M;
    float acc = 0.f;
    if (live) {
        for (long long m = m0; m < m1; ++m) {
            const int q = (int)(m % p.Q), pp = (int)((m / p.Q) % p.P), n = (int)(m / ((long long)p.P * p.Q));
            const int ih = pp * p.stride + r - p.pt, iw = q * p.stride + s - p.pl;
            if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) continue;
            const float xv = ld_as_f32(x + ((size_t)(n * p.H + ih) * p.W + iw) * p.x_ld + p.x_coff + c);
            const float gv = ld_as_f32(dy + (size_t)m * p.y_ld + p.y_coff + k);
            acc = fmaf(xv, gv, acc);
        }
        atomicAdd(dw + ((size_t)tap * p.C + c) * p.K + k, acc);
    }
}

// ---- host dispatch -----------------------------------------------------------------------
template <typename TX, typename TY>
static int launch_fprop(const ConvP& p, const void* x, const void* w_ck, const float* bias, void* y,
                        float* stats, cudaStream_t st) {
    const long long M = (long long)p.N * p.P * p.Q;
    dim3 block(128);
#define URIR_FP(KB) do { dim3 grid(cdiv(M, 128), p.K / KB); \
    conv_fprop_simt<TX, TY, KB><<<grid, block, 0, st>>>(p, (const TX*)x, (const __nv_bfloat16*)w_ck, bias, (TY*)y, stats); } while (0)
    if (p.K % 16 == 0) URIR_FP(16);
    else if (p.K % 8 == 0) URIR_FP(8);
    else if (p.K % 2 == 0) URIR_FP(2);
    else URIR_FP(1);
#undef URIR_FP
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

template <typename TDY, typename TDX>
static int launch_dgrad(const ConvP& p, const void* dy, const void* w_kc, const float* bias, void* dx,
                        float* stats, cudaStream_t st) {
    const long long M = (long long)p.N * p.H * p.W;
    dim3 block(128);
#define URIR_DG(CB) do { dim3 grid(cdiv(M, 128), p.C / CB); \
    conv_dgrad_simt<TDY, TDX, CB><<<grid, block, 0, st>>>(p, (const TDY*)dy, (const __nv_bfloat16*)w_kc, bias, (TDX*)dx, stats); } while (0)
    if (p.C % 16 == 0) URIR_DG(16);
    else if (p.C % 8 == 0) URIR_DG(8);
    else if (p.C % 2 == 0) URIR_DG(2);
    else URIR_DG(1);
#undef URIR_DG
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

template <typename TX, typename TDY>
static int launch_wgrad(const ConvP& p, const void* x, const void* dy, float* dw, cudaStream_t st) {
    const long long M = (long long)p.N * p.P * p.Q;
    int KT = p.K < 32 ? p.K : 32;
    int CT = 256 / KT; if (CT > p.C) CT = p.C;
    const int tiles = p.R * p.S * cdiv(p.C, CT) * cdiv(p.K, KT);
    // enough pixel chunks to fill the machine a few times over, but >= 256 pixels per block
    long long want_chunks = ((long long)sm_count() * 8 + tiles - 1) / tiles;
    long long pix = (M + want_chunks - 1) / want_chunks;
    if (pix < 256) pix = 256;
    if (pix > M) pix = M;
    const int chunks = cdiv(M, pix);
    if (!p.accumulate) URIR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p.R * p.S * p.C * p.K, st));
    dim3 grid(tiles, chunks), block(256);
    conv_wgrad_simt<TX, TDY><<<grid, block, 0, st>>>(p, (const TX*)x, (const TDY*)dy, dw, KT, CT, (int)pix);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

bool stem_fprop_supported(const urir_conv_desc* d, const float* stats);
bool stem_wgrad_supported(const urir_conv_desc* d);
int stem_fprop(const urir_conv_desc* d, const void* x, const void* w_ck, const float* bias, void* y, cudaStream_t st);
int stem_wgrad(const urir_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st);

int conv_fprop_simt_dispatch(const urir_conv_desc* d, const void* x, const void* w_ck, const float* bias,
                             void* y, float* stats, cudaStream_t st) {
    ConvP p = to_p(d);
    URIR_CHECK_ARG(w_ck != nullptr, "fprop(SIMT) needs w_ck");
    if (stem_fprop_supported(d, stats)) return stem_fprop(d, x, w_ck, bias, y, st);
    if (d->x_dtype == URIR_F32 && d->y_dtype == URIR_BF16) return launch_fprop<float, __nv_bfloat16>(p, x, w_ck, bias, y, stats, st);
    if (d->x_dtype == URIR_BF16 && d->y_dtype == URIR_BF16) return launch_fprop<__nv_bfloat16, __nv_bfloat16>(p, x, w_ck, bias, y, stats, st);
    if (d->x_dtype == URIR_BF16 && d->y_dtype == URIR_F32) return launch_fprop<__nv_bfloat16, float>(p, x, w_ck, bias, y, stats, st);
    return launch_fprop<float, float>(p, x, w_ck, bias, y, stats, st);
}

bool head_dgrad_supported(const urir_conv_desc* d, const float* bias, const float* stats);
int head_dgrad(const urir_conv_desc* d, const void* dy, const void* w_ck, void* dx, cudaStream_t st);

int conv_dgrad_simt_dispatch(const urir_conv_desc* d, const void* dy, const void* w_kc, const float* bias,
                             void* dx, float* stats, cudaStream_t st, const void* w_ck) {
    ConvP p = to_p(d);
    if (w_ck && head_dgrad_supported(d, bias, stats)) return head_dgrad(d, dy, w_ck, dx, st);
    URIR_CHECK_ARG(w_kc != nullptr, "dgrad(SIMT) needs w_kc");
    if (d->y_dtype == URIR_F32 && d->x_dtype == URIR_BF16) return launch_dgrad<float, __nv_bfloat16>(p, dy, w_kc, bias, dx, stats, st);
    if (d->y_dtype == URIR_BF16 && d->x_dtype == URIR_BF16) return launch_dgrad<__nv_bfloat16, __nv_bfloat16>(p, dy, w_kc, bias, dx, stats, st);
    if (d->y_dtype == URIR_BF16 && d->x_dtype == URIR_F32) return launch_dgrad<__nv_bfloat16, float>(p, dy, w_kc, bias, dx, stats, st);
    return launch_dgrad<float, float>(p, dy, w_kc, bias, dx, stats, st);
}

int conv_wgrad_simt_dispatch(const urir_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    ConvP p = to_p(d);
    if (stem_wgrad_supported(d)) return stem_wgrad(d, x, dy, dw, st);
    if (d->x_dtype == URIR_F32 && d->y_dtype == URIR_BF16) return launch_wgrad<float, __nv_bfloat16>(p, x, dy, dw, st);
    if (d->x_dtype == URIR_BF16 && d->y_dtype == URIR_BF16) return launch_wgrad<__nv_bfloat16, __nv_bfloat16>(p, x, dy, dw, st);
    if (d->x_dtype == URIR_BF16 && d->y_dtype == URIR_F32) return launch_wgrad<__nv_bfloat16, float>(p, x, dy, dw, st);
    return launch_wgrad<float, float>(p, x, dy, dw, st);
}

// fp32 HWIO -> bf16 [tap][C][K] and [tap][K][C]
__global__ void weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_ck,
                                   __nv_bfloat16* __restrict__ w_kc, int taps, int C, int K) {
    const long long n = (long long)taps * C * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K), c = (int)((i / K) % C), t = (int)(i / ((long long)K * C));
        const __nv_bfloat16 v = f2bf(w[i]);
        if (w_ck) w_ck[i] = v;
        if (w_kc) w_kc[((size_t)t * K + k) * C + c] = v;
    }
}

// all kernels of the model in one launch: table[e] = {w fp32 ptr, w_ck ptr, w_kc ptr, taps, C, K} (int64 each).
// Work unit = one 64 (c) x 64 (k) tile of one tap: fp32 rows are read coalesced, w_ck is written coalesced,
// and the transposed w_kc goes through a padded shared-memory tile so that it is written coalesced as well
// (the earlier element-per-thread version scattered 2-byte stores C*2 bytes apart).
constexpr int WP_MAX_ENTRIES = 64;
__global__ void __launch_bounds__(256) weight_prep_batched_kernel(const long long* __restrict__ table, int n_entries) {
    __shared__ long long tile_start[WP_MAX_ENTRIES + 1];
    __shared__ float tile[64][65];
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int e = 0; e < n_entries; ++e) {
            tile_start[e] = acc;
            const long long* t = table + 6 * e;
            acc += t[3] * ((t[4] + 63) / 64) * ((t[5] + 63) / 64);
        }
        tile_start[n_entries] = acc;
    }
    __syncthreads();
    const long long total = tile_start[n_entries];
    int e = 0;
    for (long long tl = blockIdx.x; tl < total; tl += gridDim.x) {
        while (tl >= tile_start[e + 1]) ++e;
        const long long* t = table + 6 * e;
        const float* w = reinterpret_cast<const float*>(t[0]);
        __nv_bfloat16* w_ck = reinterpret_cast<__nv_bfloat16*>(t[1]);
        __nv_bfloat16* w_kc = reinterpret_cast<__nv_bfloat16*>(t[2]);
        const int C = (int)t[4], K = (int)t[5];
        const int tiles_k = (K + 63) / 64, tiles_c = (C + 63) / 64;
        long long r = tl - tile_start[e];
        const int tk = (int)(r % tiles_k); r /= tiles_k;
        const int tc = (int)(r % tiles_c); const long long tap = r / tiles_c;
        const int c0 = tc * 64, k0 = tk * 64;
        const size_t base = (size_t)tap * C * K;
        const int col = threadIdx.x & 63, row0 = threadIdx.x >> 6;
        __syncthreads();                                   // previous tile's transposed reads are done
        const bool vec = (K % 4 == 0) && (C % 2 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0) &&
                         (!w_ck || (reinterpret_cast<uintptr_t>(w_ck) & 7) == 0) &&
                         (!w_kc || (reinterpret_cast<uintptr_t>(w_kc) & 3) == 0);
        if (vec) {
            // 16-byte loads (four per thread, all in flight), 8-byte w_ck stores, 4-byte transposed w_kc stores
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = threadIdx.x + 256 * j, c = c0 + (idx >> 4), k = k0 + (idx & 15) * 4;
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < C && k < K) v[j] = *reinterpret_cast<const float4*>(w + base + (size_t)c * K + k);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = threadIdx.x + 256 * j, rr = idx >> 4, cc = (idx & 15) * 4;
                const int c = c0 + rr, k = k0 + cc;
                if (w_ck && c < C && k < K) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(v[j].x, v[j].y), hi = __floats2bfloat162_rn(v[j].z, v[j].w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(w_ck + base + (size_t)c * K + k) = pk;
                }
                tile[rr][cc] = v[j].x; tile[rr][cc + 1] = v[j].y; tile[rr][cc + 2] = v[j].z; tile[rr][cc + 3] = v[j].w;
            }
            __syncthreads();
            if (w_kc) {
                const int c2 = (threadIdx.x & 31) * 2, r0 = threadIdx.x >> 5;
#pragma unroll
                for (int rr = r0; rr < 64; rr += 8) {
                    const int k = k0 + rr, c = c0 + c2;
                    if (k < K && c < C)
                        *reinterpret_cast<__nv_bfloat162*>(w_kc + base + (size_t)k * C + c) =
                            __floats2bfloat162_rn(tile[c2][rr], tile[c2 + 1][rr]);
                }
            }
            continue;
        }
#pragma unroll 4
        for (int rr = row0; rr < 64; rr += 4) {
            const int c = c0 + rr, k = k0 + col;
            float v = 0.f;
            if (c < C && k < K) {
                v = w[base + (size_t)c * K + k];
                if (w_ck) w_ck[base + (size_t)c * K + k] = f2bf(v);
            }
            tile[rr][col] = v;
        }
        __syncthreads();
        if (w_kc) {
#pragma unroll 4
            for (int rr = row0; rr < 64; rr += 4) {
                const int k = k0 + rr, c = c0 + col;
                if (k < K && c < C) w_kc[base + (size_t)k * C + c] = f2bf(tile[col][rr]);
            }
        }
    }
}

int weight_prep_batched(const long long* table_dev, int n_entries, cudaStream_t st) {
    URIR_CHECK_ARG(n_entries <= WP_MAX_ENTRIES, "weight_prep_batched: at most %d entries", WP_MAX_ENTRIES);
    weight_prep_batched_kernel<<<sm_count() * 8, 256, 0, st>>>(table_dev, n_entries);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

// Inference-mode BatchNorm folded into the preceding convolution, all layers in one launch:
// table[e] = {w fp32, scale_shift fp32 [2K], bias fp32 [K], out w_kc bf16 [tap][K][C], out bias fp32 [K], taps, C, K}.
__global__ void __launch_bounds__(256) weight_fold_bn_batched_kernel(const long long* __restrict__ table, int n_entries) {
    __shared__ long long tile_start[WP_MAX_ENTRIES + 1];
    __shared__ float tile[64][65];
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int e = 0; e < n_entries; ++e) {
            tile_start[e] = acc;
            const long long* t = table + 8 * e;
            acc += t[5] * ((t[6] + 63) / 64) * ((t[7] + 63) / 64);
        }
        tile_start[n_entries] = acc;
    }
    __syncthreads();
    const long long total = tile_start[n_entries];
    int e = 0;
    for (long long tl = blockIdx.x; tl < total; tl += gridDim.x) {
        while (tl >= tile_start[e + 1]) ++e;
        const long long* t = table + 8 * e;
        const float* w = reinterpret_cast<const float*>(t[0]);
        const float* ss = reinterpret_cast<const float*>(t[1]);
        const float* bias = reinterpret_cast<const float*>(t[2]);
        __nv_bfloat16* w_kc = reinterpret_cast<__nv_bfloat16*>(t[3]);
        float* bias_out = reinterpret_cast<float*>(t[4]);
        const int C = (int)t[6], K = (int)t[7];
        const int tiles_k = (K + 63) / 64, tiles_c = (C + 63) / 64;
        long long r = tl - tile_start[e];
        const int tk = (int)(r % tiles_k); r /= tiles_k;
        const int tc = (int)(r % tiles_c); const long long tap = r / tiles_c;
        const int c0 = tc * 64, k0 = tk * 64;
        const size_t base = (size_t)tap * C * K;
        const int col = threadIdx.x & 63, row0 = threadIdx.x >> 6;
        if (tap == 0 && tc == 0 && threadIdx.x < 64 && k0 + threadIdx.x < K) {        // the layer's folded bias, once
            const int k = k0 + threadIdx.x;
            bias_out[k] = fmaf(bias ? bias[k] : 0.f, ss[k], ss[K + k]);
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = row0; rr < 64; rr += 4) {
            const int c = c0 + rr, k = k0 + col;
            tile[rr][col] = (c < C && k < K) ? w[base + (size_t)c * K + k] * ss[k] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = row0; rr < 64; rr += 4) {
            const int k = k0 + rr, c = c0 + col;
            if (k < K && c < C) w_kc[base + (size_t)k * C + c] = f2bf(tile[col][rr]);
        }
    }
}
int weight_fold_bn_batched(const long long* table_dev, int n_entries, cudaStream_t st) {
    URIR_CHECK_ARG(n_entries <= WP_MAX_ENTRIES, "weight_fold_bn_batched: at most %d entries", WP_MAX_ENTRIES);
    weight_fold_bn_batched_kernel<<<sm_count() * 8, 256, 0, st>>>(table_dev, n_entries);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int weight_prep(const float* w, void* w_ck, void* w_kc, int taps, int C, int K, cudaStream_t st) {
    const long long n = (long long)taps * C * K;
    int blocks = cdiv(n, 256); if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    weight_prep_kernel<<<blocks, 256, 0, st>>>(w, (__nv_bfloat16*)w_ck, (__nv_bfloat16*)w_kc, taps, C, K);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
