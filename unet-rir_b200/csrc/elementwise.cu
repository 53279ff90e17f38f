// Bandwidth-bound kernels of the train step: BatchNorm(+ReLU) forward / backward, the fused
// amp/phase loss (forward scalars + dL/dy in one pass), flat Adam/SGD, and small helpers.
// All of them are HBM-bound (DESIGN.md roofline table): 128-bit accesses, one pass per tensor,
// warp-shuffle + shared-memory reductions, a handful of atomics per block.
#include "urir_common.cuh"

namespace urir {

// =========================================================================================
// BatchNormalization (Keras: eps 1e-3, momentum .99, biased batch variance) -- u_net.py:367-369
// =========================================================================================
__global__ void bn_finalize_kernel(const float* __restrict__ stats, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                   float momentum, float eps, int unbiased, float* __restrict__ scale_shift,
                                   float* __restrict__ mean_rstd, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, var;
    if (stats) {
        const double m = (double)stats[c] / count;
        double v = (double)stats[C + c] / count - m * m;
        if (v < 0) v = 0;
        mean = (float)m; var = (float)v;
        if (moving_mean) {
            const double mv = unbiased ? v * (count / (count > 1 ? count - 1 : 1)) : v;
            moving_mean[c] = moving_mean[c] * momentum + mean * (1.f - momentum);
            moving_var[c] = moving_var[c] * momentum + (float)mv * (1.f - momentum);
        }
    } else {
        mean = moving_mean[c]; var = moving_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    scale_shift[c] = g * rstd;
    scale_shift[C + c] = b - mean * g * rstd;
    if (mean_rstd) { mean_rstd[c] = mean; mean_rstd[C + c] = rstd; }
}

// y = relu(x*scale+shift), 8 channels (one 128-bit access) per thread iteration
__global__ void __launch_bounds__(256)
bn_relu_fwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff,
                   const float* __restrict__ scale_shift, __nv_bfloat16* __restrict__ y, int y_ld,
                   int y_coff, long long npix, int C, int relu) {
    extern __shared__ float ss[];   // [2C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) ss[i] = scale_shift[i];
    __syncthreads();
    const int G = C >> 3;
    const long long total = npix * G;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // 4 independent 128-bit loads in flight per thread (memory-level parallelism, guideline 7)
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
        uint4 u[4]; long long pixs[4]; int cs[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long i = i0 + q * stride;
            pixs[q] = i / G; cs[q] = (int)(i - pixs[q] * G) << 3;
            if (i < total) u[q] = ld_nc_v4(x + pixs[q] * x_ld + x_coff + cs[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (i0 + q * stride >= total) break;
            const int c = cs[q];
            const uint32_t in[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
            uint32_t out[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 v = unpack_bf16x2(in[j]);
                v.x = fmaf(v.x, ss[c + 2 * j], ss[C + c + 2 * j]);
                v.y = fmaf(v.y, ss[c + 2 * j + 1], ss[C + c + 2 * j + 1]);
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
                out[j] = pack_bf16x2(v.x, v.y);
            }
            *reinterpret_cast<uint4*>(y + pixs[q] * y_ld + y_coff + c) = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }
}

// sums[c] = sum g, sums[C+c] = sum g*xhat, g = dy * (x*scale+shift > 0)
// block = 256 threads; thread owns one 8-channel group and strides over pixels.
__global__ void __launch_bounds__(256)
bn_relu_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld, int dy_coff,
                          const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff,
                          const float* __restrict__ scale_shift, const float* __restrict__ mean_rstd,
                          float* __restrict__ sums, long long npix, int C, unsigned int* gate) {
    extern __shared__ float sm[];              // [4C] consts, then [256][17] reduction scratch
    float* cs = sm;
    float* red = sm + 4 * C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) { cs[i] = scale_shift[i]; cs[2 * C + i] = mean_rstd[i]; }
    __syncthreads();
    const int G = C >> 3;                      // channel groups
    const int lanes = 256 / G;                 // pixel lanes per block (G <= 256)
    const int g = threadIdx.x % G, lane = threadIdx.x / G;
    const int c = g << 3;
    float a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
    if (lane < lanes) {
        const long long pstride = (long long)gridDim.x * lanes;
        for (long long pix0 = (long long)blockIdx.x * lanes + lane; pix0 < npix; pix0 += 2 * pstride) {
          uint4 uds[2], uxs[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
              const long long pix = pix0 + q * pstride;
              if (pix < npix) { uds[q] = ld_nc_v4(dy + pix * dy_ld + dy_coff + c); uxs[q] = ld_nc_v4(x + pix * x_ld + x_coff + c); }
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (pix0 + q * pstride >= npix) break;
            const uint4 ud = uds[q], ux = uxs[q];
            const uint32_t d4[4] = {ud.x, ud.y, ud.z, ud.w}, x4[4] = {ux.x, ux.y, ux.z, ux.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 dv = unpack_bf16x2(d4[j]), xv = unpack_bf16x2(x4[j]);
                const int c0 = c + 2 * j, c1 = c0 + 1;
                const float g0 = fmaf(xv.x, cs[c0], cs[C + c0]) > 0.f ? dv.x : 0.f;
                const float g1 = fmaf(xv.y, cs[c1], cs[C + c1]) > 0.f ? dv.y : 0.f;
                const float h0 = (xv.x - cs[2 * C + c0]) * cs[3 * C + c0];
                const float h1 = (xv.y - cs[2 * C + c1]) * cs[3 * C + c1];
                a1[2 * j] += g0; a1[2 * j + 1] += g1;
                a2[2 * j] = fmaf(g0, h0, a2[2 * j]); a2[2 * j + 1] = fmaf(g1, h1, a2[2 * j + 1]);
            }
          }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x * 17 + j] = a1[j]; red[threadIdx.x * 17 + 8 + j] = a2[j]; }
    __syncthreads();
    // thread t < 2C : sums entry t ; reduce over pixel lanes
    gate_enter(gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
    for (int t = threadIdx.x; t < 2 * C; t += blockDim.x) {
        const int which = t / C, ch = t % C;
        const int gg = ch >> 3, j = ch & 7;
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[(l * G + gg) * 17 + which * 8 + j];
        atomicAdd(sums + t, s);
    }
    gate_leave(gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
}

__global__ void __launch_bounds__(256)
bn_relu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld, int dy_coff,
                         const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff,
                         const float* __restrict__ scale_shift, const float* __restrict__ mean_rstd,
                         const float* __restrict__ gamma, const float* __restrict__ sums,
                         __nv_bfloat16* __restrict__ dx, int dx_ld, int dx_coff,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                         long long npix, int C, unsigned int* gate) {
    extern __shared__ float sm[];   // scale, shift, mean, rstd, a=gamma*rstd, mg, mgx : 7C ; then [256][9] scratch
    const float inv_n = 1.f / (float)npix;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        sm[i] = scale_shift[i]; sm[C + i] = scale_shift[C + i];
        sm[2 * C + i] = mean_rstd[i]; sm[3 * C + i] = mean_rstd[C + i];
        sm[4 * C + i] = (gamma ? gamma[i] : 1.f) * mean_rstd[C + i];
        sm[5 * C + i] = sums[i] * inv_n; sm[6 * C + i] = sums[C + i] * inv_n;
        if (blockIdx.x == 0) { if (dgamma) dgamma[i] = sums[C + i]; if (dbeta) dbeta[i] = sums[i]; }
    }
    __syncthreads();
    const int G = C >> 3;
    const long long total = npix * G;
    float bsum[8];                  // sum of dx over this thread's (fixed, since G | 256) channel group
#pragma unroll
    for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
      uint4 uds[2], uxs[2]; long long pixs[2]; int cs[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
          const long long i = i0 + q * stride;
          pixs[q] = i / G; cs[q] = (int)(i - pixs[q] * G) << 3;
          if (i < total) { uds[q] = ld_nc_v4(dy + pixs[q] * dy_ld + dy_coff + cs[q]); uxs[q] = ld_nc_v4(x + pixs[q] * x_ld + x_coff + cs[q]); }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (i0 + q * stride >= total) break;
        const long long pix = pixs[q];
        const int c = cs[q];
        const uint4 ud = uds[q], ux = uxs[q];
        const uint32_t d4[4] = {ud.x, ud.y, ud.z, ud.w}, x4[4] = {ux.x, ux.y, ux.z, ux.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 dv = unpack_bf16x2(d4[j]), xv = unpack_bf16x2(x4[j]);
            float r[2];
            const float dvv[2] = {dv.x, dv.y}, xvv[2] = {xv.x, xv.y};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cc = c + 2 * j + e;
                const float g = fmaf(xvv[e], sm[cc], sm[C + cc]) > 0.f ? dvv[e] : 0.f;
                const float h = (xvv[e] - sm[2 * C + cc]) * sm[3 * C + cc];
                r[e] = sm[4 * C + cc] * (g - sm[5 * C + cc] - h * sm[6 * C + cc]);
            }
            o[j] = pack_bf16x2(r[0], r[1]);
            bsum[2 * j] += r[0]; bsum[2 * j + 1] += r[1];
        }
        *reinterpret_cast<uint4*>(dx + pix * dx_ld + dx_coff + c) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    if (dbias) {                    // gradient of the preceding conv's bias = sum of dx (analytically ~0)
        float* red = sm + 7 * C;
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x * 9 + j] = bsum[j];
        __syncthreads();
        gate_enter(gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
        for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
            const int gg = ch >> 3, j = ch & 7;
            float s = 0.f;
            for (int l = 0; l < 256 / G; ++l) s += red[(l * G + gg) * 9 + j];
            atomicAdd(dbias + ch, s);
        }
        gate_leave(gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
    }
}


// -----------------------------------------------------------------------------------------
// Fast paths for power-of-two channel counts (every layer of the canonical net). A thread keeps ONE 8-channel
// group for its whole life (256 % (C/8) == 0), so the per-channel coefficients sit in registers instead of
// being re-read from shared memory for every element, the pixel index advances by a constant (no 64-bit
// divisions in the loop) and UNROLL independent 128-bit loads per tensor are in flight per thread.
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ void ld8(const float* __restrict__ p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

template <int UNROLL>
__global__ void __launch_bounds__(256)
bn_relu_fwd_pow2_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff,
                        const float* __restrict__ scale_shift, __nv_bfloat16* __restrict__ y, int y_ld,
                        int y_coff, long long npix, int C, int g_shift, int relu) {
    pdl_sync();
    const int G = C >> 3, lanes = 256 >> g_shift;
    const int c = (threadIdx.x & (G - 1)) << 3, lane = threadIdx.x >> g_shift;
    float sc[8], sh[8];
    ld8(scale_shift + c, sc); ld8(scale_shift + C + c, sh);
    const __nv_bfloat16* xp = x + x_coff + c;
    __nv_bfloat16* yp = y + y_coff + c;
    const long long pstride = (long long)gridDim.x * lanes;
    for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < npix; p0 += UNROLL * pstride) {
        uint4 u[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
            if (p0 + q * pstride < npix) u[q] = ld_nc_v4(xp + (p0 + q * pstride) * x_ld);
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            if (p0 + q * pstride >= npix) break;
            float v[8];
            unpack8(u[q], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[j] = fmaf(v[j], sc[j], sh[j]); if (relu) v[j] = fmaxf(v[j], 0.f); }
            *reinterpret_cast<uint4*>(yp + (p0 + q * pstride) * y_ld) = pack8(v);
        }
    }
}

template <int UNROLL>
__global__ void __launch_bounds__(256)
bn_relu_bwd_reduce_pow2_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld, int dy_coff,
                               const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff,
                               const float* __restrict__ scale_shift, const float* __restrict__ mean_rstd,
                               float* __restrict__ sums, long long npix, int C, int g_shift, unsigned int* gate) {
    pdl_sync();
    extern __shared__ float red[];             // [256][17]
    const int G = C >> 3, lanes = 256 >> g_shift;
    const int c = (threadIdx.x & (G - 1)) << 3, lane = threadIdx.x >> g_shift;
    float sc[8], sh[8], mu[8], rs[8], a1[8], a2[8];
    ld8(scale_shift + c, sc); ld8(scale_shift + C + c, sh); ld8(mean_rstd + c, mu); ld8(mean_rstd + C + c, rs);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
    const __nv_bfloat16* dp = dy + dy_coff + c;
    const __nv_bfloat16* xp = x + x_coff + c;
    const long long pstride = (long long)gridDim.x * lanes;
    for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < npix; p0 += UNROLL * pstride) {
        uint4 ud[UNROLL], ux[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
            if (p0 + q * pstride < npix) { ud[q] = ld_nc_v4(dp + (p0 + q * pstride) * dy_ld); ux[q] = ld_nc_v4(xp + (p0 + q * pstride) * x_ld); }
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            if (p0 + q * pstride >= npix) break;
            float dv[8], xv[8];
            unpack8(ud[q], dv); unpack8(ux[q], xv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float g = fmaf(xv[j], sc[j], sh[j]) > 0.f ? dv[j] : 0.f;
                a1[j] += g;
                a2[j] = fmaf(g, (xv[j] - mu[j]) * rs[j], a2[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x * 17 + j] = a1[j]; red[threadIdx.x * 17 + 8 + j] = a2[j]; }
    __syncthreads();
    gate_enter(gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
    for (int t = threadIdx.x; t < 2 * C; t += blockDim.x) {
        const int which = t / C, ch = t % C;
        const int gg = ch >> 3, j = ch & 7;
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[(l * G + gg) * 17 + which * 8 + j];
        atomicAdd(sums + t, s);
    }
    gate_leave(gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
}

// dx = a*g + k1*x + k0 with a = gamma*rstd, k1 = -a*rstd*mean(g*xhat), k0 = -k1*mean - a*mean(g)
template <int UNROLL>
__global__ void __launch_bounds__(256)
bn_relu_bwd_apply_pow2_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld, int dy_coff,
                              const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff,
                              const float* __restrict__ scale_shift, const float* __restrict__ mean_rstd,
                              const float* __restrict__ gamma, const float* __restrict__ sums,
                              __nv_bfloat16* __restrict__ dx, int dx_ld, int dx_coff,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                              long long npix, int C, int g_shift, unsigned int* gate) {
    pdl_sync();
    extern __shared__ float red[];             // [256][9]
    const int G = C >> 3, lanes = 256 >> g_shift;
    const int c = (threadIdx.x & (G - 1)) << 3, lane = threadIdx.x >> g_shift;
    const float inv_n = 1.f / (float)npix;
    float sc[8], sh[8], a[8], k1[8], k0[8], bsum[8];
    {
        float mu[8], rs[8], sg[8], sgx[8], gm[8];
        ld8(scale_shift + c, sc); ld8(scale_shift + C + c, sh); ld8(mean_rstd + c, mu); ld8(mean_rstd + C + c, rs);
        ld8(sums + c, sg); ld8(sums + C + c, sgx);
        if (gamma) ld8(gamma + c, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = (gamma ? gm[j] : 1.f) * rs[j];
            k1[j] = -a[j] * rs[j] * (sgx[j] * inv_n);
            k0[j] = -k1[j] * mu[j] - a[j] * (sg[j] * inv_n);
            bsum[j] = 0.f;
        }
        if (blockIdx.x == 0 && lane == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { if (dgamma) dgamma[c + j] = sgx[j]; if (dbeta) dbeta[c + j] = sg[j]; }
        }
    }
    const __nv_bfloat16* dp = dy + dy_coff + c;
    const __nv_bfloat16* xp = x + x_coff + c;
    __nv_bfloat16* op = dx + dx_coff + c;
    const long long pstride = (long long)gridDim.x * lanes;
    for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < npix; p0 += UNROLL * pstride) {
        uint4 ud[UNROLL], ux[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
            if (p0 + q * pstride < npix) { ud[q] = ld_nc_v4(dp + (p0 + q * pstride) * dy_ld); ux[q] = ld_nc_v4(xp + (p0 + q * pstride) * x_ld); }
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            if (p0 + q * pstride >= npix) break;
            float dv[8], xv[8], r[8];
            unpack8(ud[q], dv); unpack8(ux[q], xv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float g = fmaf(xv[j], sc[j], sh[j]) > 0.f ? dv[j] : 0.f;
                r[j] = fmaf(a[j], g, fmaf(k1[j], xv[j], k0[j]));
                bsum[j] += r[j];
            }
            *reinterpret_cast<uint4*>(op + (p0 + q * pstride) * dx_ld) = pack8(r);
        }
    }
    if (dbias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x * 9 + j] = bsum[j];
        __syncthreads();
        gate_enter(gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
        for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
            const int gg = ch >> 3, j = ch & 7;
            float s = 0.f;
            for (int l = 0; l < lanes; ++l) s += red[(l * G + gg) * 9 + j];
            atomicAdd(dbias + ch, s);
        }
        gate_leave(gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
    }
}


// training-mode BatchNorm + ReLU in one launch: every thread derives scale / shift of its 8 channels from the
// conv epilogue's [sum | sumsq] (same arithmetic as bn_finalize_kernel), block 0 also writes scale_shift /
// mean_rstd for the backward pass and updates the moving statistics.
template <int UNROLL>
__global__ void __launch_bounds__(256)
bn_relu_fwd_train_pow2_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff,
                              const float* __restrict__ stats, double count, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ moving_mean,
                              float* __restrict__ moving_var, float momentum, float eps, int unbiased,
                              float* __restrict__ scale_shift, float* __restrict__ mean_rstd,
                              __nv_bfloat16* __restrict__ y, int y_ld, int y_coff, long long npix, int C, int g_shift) {
    pdl_sync();
    const int G = C >> 3, lanes = 256 >> g_shift;
    const int c = (threadIdx.x & (G - 1)) << 3, lane = threadIdx.x >> g_shift;
    float sc[8], sh[8];
    {
        float s1[8], s2[8], gm[8], bt[8];
        ld8(stats + c, s1); ld8(stats + C + c, s2);
        if (gamma) ld8(gamma + c, gm);
        if (beta) ld8(beta + c, bt);
        const bool writer = blockIdx.x == 0 && lane == 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double m = (double)s1[j] / count;
            double v = (double)s2[j] / count - m * m;
            if (v < 0) v = 0;
            const float mean = (float)m, var = (float)v;
            const float rstd = rsqrtf(var + eps);
            const float g = gamma ? gm[j] : 1.f, b = beta ? bt[j] : 0.f;
            sc[j] = g * rstd;
            sh[j] = b - mean * g * rstd;
            if (writer) {
                if (moving_mean) {
                    const double mv = unbiased ? v * (count / (count > 1 ? count - 1 : 1)) : v;
                    moving_mean[c + j] = moving_mean[c + j] * momentum + mean * (1.f - momentum);
                    moving_var[c + j] = moving_var[c + j] * momentum + (float)mv * (1.f - momentum);
                }
                scale_shift[c + j] = sc[j]; scale_shift[C + c + j] = sh[j];
                if (mean_rstd) { mean_rstd[c + j] = mean; mean_rstd[C + c + j] = rstd; }
            }
        }
    }
    const __nv_bfloat16* xp = x + x_coff + c;
    __nv_bfloat16* yp = y + y_coff + c;
    const long long pstride = (long long)gridDim.x * lanes;
    for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < npix; p0 += UNROLL * pstride) {
        uint4 u[UNROLL];
#pragma unroll
        for (int q = 0; q < UNROLL; ++q)
            if (p0 + q * pstride < npix) u[q] = ld_nc_v4(xp + (p0 + q * pstride) * x_ld);
#pragma unroll
        for (int q = 0; q < UNROLL; ++q) {
            if (p0 + q * pstride >= npix) break;
            float v[8];
            unpack8(u[q], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
            *reinterpret_cast<uint4*>(yp + (p0 + q * pstride) * y_ld) = pack8(v);
        }
    }
}

// C/8 a power of two <= 256 -> log2(C/8), else -1
static int pow2_shift(int C) {
    if (C % 8) return -1;
    const int G = C / 8;
    if (G < 1 || G > 256 || (G & (G - 1))) return -1;
    int s = 0; while ((1 << s) < G) ++s;
    return s;
}
// enough blocks to fill the machine, few enough that the per-block tail (atomics) stays small
static int pow2_grid(long long npix, int lanes, int unroll, int max_per_sm) {
    long long b = (npix + (long long)lanes * unroll - 1) / ((long long)lanes * unroll);
    const long long cap = (long long)sm_count() * max_per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

static int grid_for(long long work_items, int threads) {
    long long b = (work_items + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

static bool vec8_ok(int C, int ld, int coff) { return C % 8 == 0 && ld % 8 == 0 && coff % 8 == 0; }

int bn_finalize(const float* stats, double count, const float* gamma, const float* beta, float* mm, float* mv,
                float momentum, float eps, int unbiased, float* scale_shift, float* mean_rstd, int C, cudaStream_t st) {
    URIR_CHECK_ARG(C > 0 && scale_shift, "bn_finalize: bad args");
    URIR_CHECK_ARG(stats || (mm && mv), "bn_finalize: inference mode needs moving statistics");
    bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(stats, count, gamma, beta, mm, mv, momentum, eps, unbiased,
                                                     scale_shift, mean_rstd, C);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int bn_relu_fwd(const void* x, int x_ld, int x_coff, const float* ss, void* y, int y_ld, int y_coff,
                long long npix, int C, int relu, cudaStream_t st) {
    URIR_CHECK_ARG(vec8_ok(C, x_ld, x_coff) && vec8_ok(C, y_ld, y_coff), "bn_relu_fwd: C/ld/coff must be multiples of 8");
    URIR_CHECK_ARG(C <= 2048, "bn_relu_fwd: C too large");
    if (const int gs = pow2_shift(C); gs >= 0) {
        const int lanes = 256 >> gs;
        URIR_CUDA_OK(launch_pdl(bn_relu_fwd_pow2_kernel<4>, dim3(pow2_grid(npix, lanes, 4, 4)), dim3(256), 0, st,
            (const __nv_bfloat16*)x, x_ld, x_coff, ss, (__nv_bfloat16*)y, y_ld, y_coff, npix, C, gs, relu));
        URIR_LAUNCH_OK(0);
        return URIR_OK;
    }
    bn_relu_fwd_kernel<<<grid_for(npix * (C / 8), 256), 256, 2 * C * sizeof(float), st>>>(
        (const __nv_bfloat16*)x, x_ld, x_coff, ss, (__nv_bfloat16*)y, y_ld, y_coff, npix, C, relu);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int bn_relu_fwd_train(const void* x, int x_ld, int x_coff, const float* stats, double count, const float* gamma,
                      const float* beta, float* mm, float* mv, float momentum, float eps, int unbiased,
                      float* scale_shift, float* mean_rstd, void* y, int y_ld, int y_coff, long long npix, int C,
                      cudaStream_t st) {
    URIR_CHECK_ARG(x && y && stats && scale_shift && npix > 0 && C > 0, "bn_relu_fwd_train: bad args");
    URIR_CHECK_ARG(vec8_ok(C, x_ld, x_coff) && vec8_ok(C, y_ld, y_coff), "bn_relu_fwd_train: C/ld/coff must be multiples of 8");
    const int gs = pow2_shift(C);
    if (gs < 0) {       // general channel counts: the two-launch path
        int rc = bn_finalize(stats, count, gamma, beta, mm, mv, momentum, eps, unbiased, scale_shift, mean_rstd, C, st);
        if (rc) return rc;
        return bn_relu_fwd(x, x_ld, x_coff, scale_shift, y, y_ld, y_coff, npix, C, 1, st);
    }
    const int lanes = 256 >> gs;
    URIR_CUDA_OK(launch_pdl(bn_relu_fwd_train_pow2_kernel<4>, dim3(pow2_grid(npix, lanes, 4, 4)), dim3(256), 0, st,
        (const __nv_bfloat16*)x, x_ld, x_coff, stats, count, gamma, beta, mm, mv, momentum, eps, unbiased, scale_shift,
        mean_rstd, (__nv_bfloat16*)y, y_ld, y_coff, npix, C, gs));
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int bn_relu_bwd_reduce(const void* dy, int dy_ld, int dy_coff, const void* x, int x_ld, int x_coff, const float* ss,
                       const float* mr, float* sums, long long npix, int C, int prezeroed, cudaStream_t st) {
    URIR_CHECK_ARG(vec8_ok(C, x_ld, x_coff) && vec8_ok(C, dy_ld, dy_coff), "bn_bwd_reduce: C/ld/coff must be multiples of 8");
    URIR_CHECK_ARG(C <= 2048 && (C / 8) <= 256, "bn_bwd_reduce: C too large");
    if (!prezeroed) URIR_CUDA_OK(cudaMemsetAsync(sums, 0, 2 * C * sizeof(float), st));
    if (const int gs = pow2_shift(C); gs >= 0) {
        const int ln = 256 >> gs;
        URIR_CUDA_OK(launch_pdl(bn_relu_bwd_reduce_pow2_kernel<4>, dim3(pow2_grid(npix, ln, 8, 2)), dim3(256), 256 * 17 * sizeof(float), st,
            (const __nv_bfloat16*)dy, dy_ld, dy_coff, (const __nv_bfloat16*)x, x_ld, x_coff, ss, mr, sums, npix, C, gs, next_gate()));
        URIR_LAUNCH_OK(0);
        return URIR_OK;
    }
    const int lanes = 256 / (C / 8);
    long long blocks = (npix + lanes - 1) / lanes;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    const size_t smem = (4 * C + 256 * 17) * sizeof(float);
    bn_relu_bwd_reduce_kernel<<<(int)blocks, 256, smem, st>>>((const __nv_bfloat16*)dy, dy_ld, dy_coff,
                                                              (const __nv_bfloat16*)x, x_ld, x_coff, ss, mr, sums, npix, C, next_gate());
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int bn_relu_bwd_apply(const void* dy, int dy_ld, int dy_coff, const void* x, int x_ld, int x_coff, const float* ss,
                      const float* mr, const float* gamma, const float* sums, void* dx, int dx_ld, int dx_coff,
                      float* dgamma, float* dbeta, float* dbias, long long npix, int C, int prezeroed, cudaStream_t st) {
    URIR_CHECK_ARG(vec8_ok(C, x_ld, x_coff) && vec8_ok(C, dy_ld, dy_coff) && vec8_ok(C, dx_ld, dx_coff),
                   "bn_bwd_apply: C/ld/coff must be multiples of 8");
    URIR_CHECK_ARG(C <= 1024, "bn_bwd_apply: C too large");
    URIR_CHECK_ARG(!dbias || (256 % (C / 8) == 0), "bn_bwd_apply: dbias needs C/8 to divide 256");
    if (dbias && !prezeroed) URIR_CUDA_OK(cudaMemsetAsync(dbias, 0, C * sizeof(float), st));
    if (const int gs = pow2_shift(C); gs >= 0) {
        const int ln = 256 >> gs;
        URIR_CUDA_OK(launch_pdl(bn_relu_bwd_apply_pow2_kernel<4>, dim3(pow2_grid(npix, ln, 8, 2)), dim3(256), 256 * 9 * sizeof(float), st,
            (const __nv_bfloat16*)dy, dy_ld, dy_coff, (const __nv_bfloat16*)x, x_ld, x_coff, ss, mr, gamma, sums,
            (__nv_bfloat16*)dx, dx_ld, dx_coff, dgamma, dbeta, dbias, npix, C, gs, dbias ? next_gate() : nullptr));
        URIR_LAUNCH_OK(0);
        return URIR_OK;
    }
    bn_relu_bwd_apply_kernel<<<grid_for(npix * (C / 8), 256), 256, (7 * C + 256 * 9) * sizeof(float), st>>>(
        (const __nv_bfloat16*)dy, dy_ld, dy_coff, (const __nv_bfloat16*)x, x_ld, x_coff, ss, mr, gamma, sums,
        (__nv_bfloat16*)dx, dx_ld, dx_coff, dgamma, dbeta, dbias, npix, C, dbias ? next_gate() : nullptr);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

// =========================================================================================
// per-channel sums (bias gradients)
// =========================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
channel_sum_kernel(const T* __restrict__ x, long long npix, int C, int ld, int coff, float* __restrict__ out, unsigned int* gate) {
    // thread owns channel (threadIdx.x % C) when C <= 256, strides over pixels
    const int lanes = 256 / C > 0 ? 256 / C : 1;
    __shared__ float red[256];
    gate_enter(gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);     // (deterministic mode serialises the whole CTA body)
    for (int cb = 0; cb < C; cb += 256) {
        const int c = cb + threadIdx.x % (C < 256 ? C : 256);
        const int lane = threadIdx.x / (C < 256 ? C : 256);
        float s = 0.f;
        if (c < C && lane < lanes)
            for (long long pix = (long long)blockIdx.x * lanes + lane; pix < npix; pix += (long long)gridDim.x * lanes)
                s += ld_as_f32(x + pix * ld + coff + c);
        red[threadIdx.x] = s;
        __syncthreads();
        if (lane == 0 && c < C) {
            const int stride = C < 256 ? C : 256;
            for (int l = 1; l < lanes; ++l) s += red[l * stride + (threadIdx.x % stride)];
            atomicAdd(out + c, s);
        }
        __syncthreads();
    }
    gate_leave(gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
}

int channel_sum(const void* x, int dtype, long long npix, int C, int ld, int coff, float* out, cudaStream_t st) {
    URIR_CHECK_ARG(C > 0 && npix > 0, "channel_sum: bad args");
    URIR_CUDA_OK(cudaMemsetAsync(out, 0, C * sizeof(float), st));
    const int lanes = 256 / C > 0 ? 256 / C : 1;
    long long blocks = (npix + lanes - 1) / lanes;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    unsigned int* gate = next_gate();
    if (dtype == URIR_BF16) channel_sum_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, npix, C, ld, coff, out, gate);
    else channel_sum_kernel<float><<<(int)blocks, 256, 0, st>>>((const float*)x, npix, C, ld, coff, out, gate);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

// =========================================================================================
// amp/phase loss: forward scalars and dL/dy_pred (optionally through the sigmoid) in one pass
// amp_phase_trainer.py:143-168 ; main_training.py:184-190, 203-235
// =========================================================================================
__global__ void __launch_bounds__(256)
ampphase_loss_kernel(const float4* __restrict__ yt, const float4* __restrict__ yp, long long npair,
                     long long npix, float w_amp, float w_ph, int sigmoid_bwd, float* __restrict__ losses,
                     float4* __restrict__ grad, __nv_bfloat16* __restrict__ grad16, int ld16, unsigned int* gate) {
    const float TWO_PI = 6.283185307179586f;
    float sse = 0.f, pc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npair;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 t = __ldg(yt + i), p = __ldg(yp + i);     // (amp0, ph0, amp1, ph1)
        const float da0 = p.x - t.x, da1 = p.z - t.z;
        const float d0 = TWO_PI * (t.y - p.y), d1 = TWO_PI * (t.w - p.w);
        float s0, c0, s1, c1;
        sincosf(d0, &s0, &c0); sincosf(d1, &s1, &c1);
        sse += da0 * da0 + da1 * da1;
        pc += (1.f - c0) + (1.f - c1);
        if (grad) {
            float4 g;
            g.x = 2.f * w_amp * da0; g.z = 2.f * w_amp * da1;
            g.y = -TWO_PI * w_ph * s0; g.w = -TWO_PI * w_ph * s1;
            if (sigmoid_bwd) { g.x *= p.x * (1.f - p.x); g.y *= p.y * (1.f - p.y); g.z *= p.z * (1.f - p.z); g.w *= p.w * (1.f - p.w); }
            grad[i] = g;
            if (grad16) {       // bf16 copy, `ld16` elements per pixel (only the first two are written)
                *reinterpret_cast<uint32_t*>(grad16 + (2 * i) * ld16) = pack_bf16x2(g.x, g.y);
                *reinterpret_cast<uint32_t*>(grad16 + (2 * i + 1) * ld16) = pack_bf16x2(g.z, g.w);
            }
        }
    }
    __shared__ float r1[8], r2[8];
    sse = warp_sum(sse); pc = warp_sum(pc);
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = sse; r2[threadIdx.x >> 5] = pc; }
    __syncthreads();
    gate_enter(gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += r1[i]; b += r2[i]; }
        const float inv = 1.f / (float)npix;
        atomicAdd(losses + 0, w_amp * a + w_ph * b);
        atomicAdd(losses + 1, b * inv);
        atomicAdd(losses + 2, a * inv);
    }
    gate_leave(gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
}

// Mean-squared error over BOTH channels, the loss of the generic trainer (trainer.py:146-156: `loss =
// amplitude_loss(y_true, y_pred)`), with the amp / phase metrics it reports beside it.
// losses = [w * SSE(all), mean(1 - cos) of the phase channel, mean sq err of the amp channel, mean sq err of all];
// grad = 2 w (p - t) (times p (1 - p) with sigmoid_bwd).
__global__ void __launch_bounds__(256)
mse2_loss_kernel(const float4* __restrict__ yt, const float4* __restrict__ yp, long long npair, long long npix, float w,
                 int sigmoid_bwd, float* __restrict__ losses, float4* __restrict__ grad, unsigned int* gate) {
    const float TWO_PI = 6.283185307179586f;
    float sa = 0.f, sp = 0.f, pc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += (long long)gridDim.x * blockDim.x) {
        const float4 t = __ldg(yt + i), p = __ldg(yp + i);
        const float d0 = p.x - t.x, d1 = p.y - t.y, d2 = p.z - t.z, d3 = p.w - t.w;
        sa += d0 * d0 + d2 * d2; sp += d1 * d1 + d3 * d3;
        pc += (1.f - cosf(TWO_PI * d1)) + (1.f - cosf(TWO_PI * d3));
        if (grad) {
            float4 g = make_float4(2.f * w * d0, 2.f * w * d1, 2.f * w * d2, 2.f * w * d3);
            if (sigmoid_bwd) { g.x *= p.x * (1.f - p.x); g.y *= p.y * (1.f - p.y); g.z *= p.z * (1.f - p.z); g.w *= p.w * (1.f - p.w); }
            grad[i] = g;
        }
    }
    __shared__ float r1[8], r2[8], r3[8];
    sa = warp_sum(sa); sp = warp_sum(sp); pc = warp_sum(pc);
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = sa; r2[threadIdx.x >> 5] = sp; r3[threadIdx.x >> 5] = pc; }
    __syncthreads();
    gate_enter(gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int i = 0; i < 8; ++i) { a += r1[i]; b += r2[i]; c += r3[i]; }
        const float inv = 1.f / (float)npix;
        atomicAdd(losses + 0, w * (a + b));
        atomicAdd(losses + 1, c * inv);
        atomicAdd(losses + 2, a * inv);
        atomicAdd(losses + 3, (a + b) * inv * 0.5f);
    }
    gate_leave(gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
}
int mse2_loss(const float* yt, const float* yp, long long npix, float w, int sigmoid_bwd, float* losses, float* grad, cudaStream_t st) {
    URIR_CHECK_ARG(npix > 0 && npix % 2 == 0, "mse2_loss: npix must be even");
    URIR_CUDA_OK(cudaMemsetAsync(losses, 0, 4 * sizeof(float), st));
    const long long npair = npix / 2;
    mse2_loss_kernel<<<grid_for(npair, 256), 256, 0, st>>>((const float4*)yt, (const float4*)yp, npair, npix, w, sigmoid_bwd,
                                                           losses, (float4*)grad, next_gate());
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int ampphase_loss(const float* yt, const float* yp, long long npix, float w_amp, float w_ph, int sigmoid_bwd,
                  float* losses, float* grad, void* grad16, int ld16, cudaStream_t st) {
    URIR_CHECK_ARG(npix > 0 && npix % 2 == 0, "ampphase_loss: npix must be even");
    URIR_CHECK_ARG(!grad16 || (grad && ld16 >= 2 && ld16 % 2 == 0), "ampphase_loss: grad16 needs grad and an even pitch >= 2");
    URIR_CUDA_OK(cudaMemsetAsync(losses, 0, 4 * sizeof(float), st));
    const long long npair = npix / 2;
    ampphase_loss_kernel<<<grid_for(npair, 256), 256, 0, st>>>((const float4*)yt, (const float4*)yp, npair, npix,
                                                               w_amp, w_ph, sigmoid_bwd, losses, (float4*)grad,
                                                               (__nv_bfloat16*)grad16, ld16, next_gate());
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

// =========================================================================================
// optimisers (Keras conventions: eps outside the bias-corrected sqrt)
// =========================================================================================
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, const float* __restrict__ lr_dev, const int* __restrict__ step_dev, float b1, float b2,
            float eps) {
    const float t = (float)(*step_dev + 1);
    const float lr_t = *lr_dev * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
#define URIR_ADAM1(f) mm.f = b1 * mm.f + (1.f - b1) * gg.f; vv.f = b2 * vv.f + (1.f - b2) * gg.f * gg.f; \
                      pp.f -= lr_t * mm.f / (sqrtf(vv.f) + eps);
        URIR_ADAM1(x) URIR_ADAM1(y) URIR_ADAM1(z) URIR_ADAM1(w)
#undef URIR_ADAM1
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gg = g[i];
        const float mm = b1 * m[i] + (1.f - b1) * gg, vv = b2 * v[i] + (1.f - b2) * gg * gg;
        m[i] = mm; v[i] = vv;
        p[i] -= lr_t * mm / (sqrtf(vv) + eps);
    }
}

// tf.keras.optimizers.Nadam (optimizer_v2/nadam.py; amp_phase_trainer.py:30-31 picks it for names containing "nadam"):
//   u_t = b1 * (1 - 0.5 * 0.96^(0.004 t)),  m_schedule_t = prod_{i<=t} u_i  (kept in coef[0] across steps)
//   g' = g / (1 - m_schedule_t);  m = b1 m + (1-b1) g;  m' = m / (1 - m_schedule_t * u_{t+1})
//   v = b2 v + (1-b2) g^2;  v' = v / (1 - b2^t);  w -= lr * ((1 - u_t) g' + u_{t+1} m') / (sqrt(v') + eps)
// nadam_prepare (one thread) advances the schedule product on the device, so a captured graph follows it.
// coef = [m_schedule, c_g = (1-u_t)/(1-ms_t), c_m = u_{t+1}/(1-ms_t*u_{t+1}), 1/(1-b2^t)]; a fresh state has coef[0] = 1.
__global__ void nadam_prepare_kernel(float* __restrict__ coef, const int* __restrict__ step_dev, float b1, float b2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double t = (double)(*step_dev + 1);
    const double u_t = (double)b1 * (1.0 - 0.5 * pow(0.96, 0.004 * t));
    const double u_t1 = (double)b1 * (1.0 - 0.5 * pow(0.96, 0.004 * (t + 1.0)));
    const double ms = (*step_dev == 0 ? 1.0 : (double)coef[0]) * u_t;
    const double ms_next = ms * u_t1;
    coef[0] = (float)ms;
    coef[1] = (float)((1.0 - u_t) / (1.0 - ms));
    coef[2] = (float)(u_t1 / (1.0 - ms_next));
    coef[3] = (float)(1.0 / (1.0 - pow((double)b2, t)));
}
__global__ void __launch_bounds__(256)
nadam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
             const float* __restrict__ lr_dev, const float* __restrict__ coef, float b1, float b2, float eps) {
    const float lr = *lr_dev, cg = coef[1], cm = coef[2], cv = coef[3];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gg = g[i];
        const float mm = b1 * m[i] + (1.f - b1) * gg, vv = b2 * v[i] + (1.f - b2) * gg * gg;
        m[i] = mm; v[i] = vv;
        p[i] -= lr * (cg * gg + cm * mm) / (sqrtf(vv * cv) + eps);
    }
}

// tensorflow_addons LAMB (trainer.py:37-38; defaults b1 .9, b2 .999, eps 1e-6, weight decay 0): per VARIABLE trust ratio
//   u = m_hat / (sqrt(v_hat) + eps) (+ wd * w);  r = ||w|| / ||u|| (1 when either norm is 0);  w -= lr * r * u.
// table[e] = {offset, count} of variable e in the flat buffers. Pass 1 (grid.y = variable) updates m, v, stores u in
// `upd` and accumulates the two squared norms into norms[2e], norms[2e+1] (caller-zeroed); pass 2 applies.
__global__ void __launch_bounds__(256)
lamb_stage1_kernel(const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   float* __restrict__ upd, const long long* __restrict__ table, float* __restrict__ norms,
                   const int* __restrict__ step_dev, float b1, float b2, float eps, float wd) {
    const long long off = table[2 * blockIdx.y], cnt = table[2 * blockIdx.y + 1];
    const float t = (float)(*step_dev + 1);
    const float c1 = 1.f / (1.f - powf(b1, t)), c2 = 1.f / (1.f - powf(b2, t));
    float sw = 0.f, su = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) {
        const long long j = off + i;
        const float gg = g[j], w = p[j];
        const float mm = b1 * m[j] + (1.f - b1) * gg, vv = b2 * v[j] + (1.f - b2) * gg * gg;
        m[j] = mm; v[j] = vv;
        const float u = (mm * c1) / (sqrtf(vv * c2) + eps) + wd * w;
        upd[j] = u;
        sw = fmaf(w, w, sw); su = fmaf(u, u, su);
    }
    __shared__ float r1[8], r2[8];
    sw = warp_sum(sw); su = warp_sum(su);
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = sw; r2[threadIdx.x >> 5] = su; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; ++i) { a += r1[i]; b += r2[i]; }
        atomicAdd(norms + 2 * blockIdx.y, a); atomicAdd(norms + 2 * blockIdx.y + 1, b);
    }
}
__global__ void __launch_bounds__(256)
lamb_stage2_kernel(float* __restrict__ p, const float* __restrict__ upd, const long long* __restrict__ table,
                   const float* __restrict__ norms, const float* __restrict__ lr_dev) {
    const long long off = table[2 * blockIdx.y], cnt = table[2 * blockIdx.y + 1];
    const float wn = sqrtf(norms[2 * blockIdx.y]), un = sqrtf(norms[2 * blockIdx.y + 1]);
    const float ratio = (wn > 0.f && un > 0.f) ? wn / un : 1.f;
    const float s = *lr_dev * ratio;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x)
        p[off + i] -= s * upd[off + i];
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long long n, const float* __restrict__ lr_dev) {
    const float lr = *lr_dev;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] -= lr * g[i];
}

__global__ void step_inc_kernel(int* s) { if (threadIdx.x == 0 && blockIdx.x == 0) *s += 1; }

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = fmaf(a, x[i], y[i]);
}

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, long long n, float scale, float* __restrict__ out) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s = fmaf(x[i], x[i], s);
    __shared__ float r[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float a = 0.f; for (int i = 0; i < 8; ++i) a += r[i]; atomicAdd(out, a * scale); }
}

// L2 kernel regulariser of every regularised tensor in ONE launch (main_training.py:232-233):
// table[e] = {param fp32 ptr, grad fp32 ptr, n} (int64 each); out[0] += coef * sum p^2 ; grad += 2 * coef * p.
__global__ void __launch_bounds__(256)
l2_reg_batched_kernel(const long long* __restrict__ table, float coef, float* __restrict__ out, unsigned int* gate) {
    const long long* e = table + 3 * blockIdx.y;
    const float* p = reinterpret_cast<const float*>(e[0]);
    float* g = reinterpret_cast<float*>(e[1]);
    const long long n = e[2];
    const float c2 = 2.f * coef;
    float s = 0.f;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
    const long long n4 = vec ? (n >> 2) : 0;
    for (long long i = tid; i < n4; i += nth) {            // 16-byte accesses
        const float4 v = reinterpret_cast<const float4*>(p)[i];
        float4 gg = reinterpret_cast<float4*>(g)[i];
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
        gg.x = fmaf(c2, v.x, gg.x); gg.y = fmaf(c2, v.y, gg.y); gg.z = fmaf(c2, v.z, gg.z); gg.w = fmaf(c2, v.w, gg.w);
        reinterpret_cast<float4*>(g)[i] = gg;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += nth) {
        const float v = p[i];
        s = fmaf(v, v, s);
        g[i] = fmaf(c2, v, g[i]);
    }
    __shared__ float r[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = s;
    __syncthreads();
    gate_enter(gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += r[i];
        if (a != 0.f) atomicAdd(out, a * coef);
    }
    gate_leave(gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
}

__global__ void add_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 ua = a[i], ub = b[i];
        const uint32_t aa[4] = {ua.x, ua.y, ua.z, ua.w}, bb[4] = {ub.x, ub.y, ub.z, ub.w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { float2 x = unpack_bf16x2(aa[j]), y = unpack_bf16x2(bb[j]); r[j] = pack_bf16x2(x.x + y.x, x.y + y.y); }
        o[i] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = f2bf(x[i]);
}

// fp32 [npix][C] -> bf16 [npix][ld] (first C channels; the rest of each row is left untouched / pre-zeroed)
__global__ void cast_pad_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long npix, int C, int ld) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix * C; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i / C;
        y[pix * ld + (i - pix * C)] = f2bf(x[i]);
    }
}
int cast_pad_bf16(const float* x, void* y, long long npix, int C, int ld, cudaStream_t st) {
    URIR_CHECK_ARG(npix > 0 && C > 0 && ld >= C, "cast_pad_bf16: bad args");
    cast_pad_bf16_kernel<<<grid_for(npix * C, 256), 256, 0, st>>>(x, (__nv_bfloat16*)y, npix, C, ld);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int adam(float* p, const float* g, float* m, float* v, long long n, const float* lr, const int* step, float b1,
         float b2, float eps, cudaStream_t st) {
    URIR_CHECK_ARG(n > 0 && ((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0),
                   "adam: buffers must be 16-byte aligned");
    adam_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, st>>>(p, g, m, v, n, lr, step, b1, b2, eps);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int nadam(float* p, const float* g, float* m, float* v, long long n, const float* lr, const int* step, float* coef,
          float b1, float b2, float eps, cudaStream_t st) {
    nadam_prepare_kernel<<<1, 32, 0, st>>>(coef, step, b1, b2);
    URIR_LAUNCH_OK(0);
    nadam_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, g, m, v, n, lr, coef, b1, b2, eps);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int lamb(float* p, const float* g, float* m, float* v, float* upd, const long long* table, int n_vars, float* norms,
         const float* lr, const int* step, float b1, float b2, float eps, float wd, cudaStream_t st) {
    URIR_CUDA_OK(cudaMemsetAsync(norms, 0, sizeof(float) * 2 * (size_t)n_vars, st));
    lamb_stage1_kernel<<<dim3(64, n_vars), 256, 0, st>>>(p, g, m, v, upd, table, norms, step, b1, b2, eps, wd);
    URIR_LAUNCH_OK(0);
    lamb_stage2_kernel<<<dim3(64, n_vars), 256, 0, st>>>(p, upd, table, norms, lr);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int sgd(float* p, const float* g, long long n, const float* lr, cudaStream_t st) {
    sgd_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, g, n, lr);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int step_increment(int* s, cudaStream_t st) { step_inc_kernel<<<1, 32, 0, st>>>(s); URIR_LAUNCH_OK(0); return URIR_OK; }
int axpy(float* y, const float* x, float a, long long n, cudaStream_t st) {
    axpy_kernel<<<grid_for(n, 256), 256, 0, st>>>(y, x, a, n);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int sumsq(const float* x, long long n, float scale, float* out, int accumulate, cudaStream_t st) {
    if (!accumulate) URIR_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float), st));
    sumsq_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, n, scale, out);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int l2_reg_batched(const long long* table_dev, int n_entries, float coef, float* out, cudaStream_t st) {
    URIR_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float), st));
    l2_reg_batched_kernel<<<dim3(sm_count() * 2, n_entries), 256, 0, st>>>(table_dev, coef, out, next_gate());
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
// out[p, c] = a[p, c] + b[p, c] for channel slices of NHWC buffers (the residual Adds of block modes 2 / 3 and their
// gradient fan-in); out may alias a or b. 8 channels (128 bits) per thread.
__global__ void __launch_bounds__(256)
add_bf16_strided_kernel(const __nv_bfloat16* __restrict__ a, int a_ld, int a_coff, const __nv_bfloat16* __restrict__ b,
                        int b_ld, int b_coff, __nv_bfloat16* __restrict__ o, int o_ld, int o_coff, long long npix, int C) {
    const int G = C >> 3;
    const long long total = npix * G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i / G;
        const int c = (int)(i - pix * G) << 3;
        const uint4 ua = *reinterpret_cast<const uint4*>(a + pix * a_ld + a_coff + c);
        const uint4 ub = *reinterpret_cast<const uint4*>(b + pix * b_ld + b_coff + c);
        const uint32_t aa[4] = {ua.x, ua.y, ua.z, ua.w}, bb[4] = {ub.x, ub.y, ub.z, ub.w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 x = unpack_bf16x2(aa[j]), y = unpack_bf16x2(bb[j]);
            r[j] = pack_bf16x2(x.x + y.x, x.y + y.y);
        }
        *reinterpret_cast<uint4*>(o + pix * o_ld + o_coff + c) = make_uint4(r[0], r[1], r[2], r[3]);
    }
}
int add_bf16_strided(const void* a, int a_ld, int a_coff, const void* b, int b_ld, int b_coff, void* o, int o_ld, int o_coff,
                     long long npix, int C, cudaStream_t st) {
    URIR_CHECK_ARG(vec8_ok(C, a_ld, a_coff) && vec8_ok(C, b_ld, b_coff) && vec8_ok(C, o_ld, o_coff),
                   "add_bf16_strided: C/ld/coff must be multiples of 8");
    add_bf16_strided_kernel<<<grid_for(npix * (C / 8), 256), 256, 0, st>>>((const __nv_bfloat16*)a, a_ld, a_coff,
        (const __nv_bfloat16*)b, b_ld, b_coff, (__nv_bfloat16*)o, o_ld, o_coff, npix, C);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int add_bf16(const void* a, const void* b, void* o, long long n, cudaStream_t st) {
    URIR_CHECK_ARG(n % 8 == 0, "add_bf16: n must be a multiple of 8");
    add_bf16_kernel<<<grid_for(n / 8, 256), 256, 0, st>>>((const uint4*)a, (const uint4*)b, (uint4*)o, n / 8);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}
int cast_f32_to_bf16(const float* x, void* y, long long n, cudaStream_t st) {
    cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, (__nv_bfloat16*)y, n);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
