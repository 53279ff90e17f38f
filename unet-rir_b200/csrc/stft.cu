// Signal path on the GPU: replaces the per-file librosa/numpy CPU code of
//   preprocess.py:13-18 (librosa.stft, abs, angle), :26-32 (normalise), :70-105 (zero-pad),
//   :56 (mean removal)                                     -> stft_ampphase_kernel
//   postprocess.py:87-113 (un-pad, denormalise), :128-129 (polar -> complex, librosa.istft)
//                                                          -> istft_ampphase_kernel
// Both are HBM-bound by their algorithmic traffic (38.4 kB of waveform <-> 184.3 kB of padded spectrogram per
// sample), so the design is about (a) 128-bit, line-aligned global accesses in both directions and (b) keeping the
// arithmetic per frame small enough not to get in the way:
//   * the real 256-point transform of a frame is ONE 128-point complex FFT of z[m] = x[2m] + i x[2m+1] plus an
//     untangling pass (and its mirror image for the inverse) -- half the butterflies of a complex 256-point FFT;
//   * a half-warp owns a frame: 8 points per thread, 128 = 8 x 4 x 4, i.e. one radix-8 and two radix-4 passes in
//     registers with two shared-memory exchanges (16-byte accesses, rows pitched at 80 bytes so that they are
//     bank-conflict free) instead of the eight dependent shared-memory passes of a radix-2 FFT;
//   * a CTA pass covers 16 frames: the waveform slice is staged with float4 loads, every spectrogram row is written
//     as one 128-byte line (eight float4 lanes per row), the inverse reads [bin][16 frames] tiles the same way and
//     converts polar -> complex once per element while staging them;
//   * the waveform mean (Loader.load) is computed once per sample: the CTAs of a sample form a thread-block cluster,
//     each sums one slice and the partial sums are exchanged through distributed shared memory.
#include <cooperative_groups.h>
#include "urir_common.cuh"

namespace cg = cooperative_groups;

namespace urir {

constexpr int NFFT = 256;
constexpr int HOP = 64;
constexpr int FR_PER_CTA = 16;              // frames per CTA pass = half-warps per CTA
constexpr int EX_PITCH = 10;                // complex values per exchange row (8 used): 80 bytes
constexpr int EX_FRAME = 16 * EX_PITCH;     // exchange buffer of one frame: 160 complex = 1280 bytes
constexpr int ISTFT_HOPS = 13;              // 13 hops of output overlap exactly 16 frames when n_fft = 4 hop
constexpr int OUT_PITCH = 36;               // floats per staged spectrogram row (32 used)
constexpr int MAX_CLUSTER = 8;

// MUFU approximations with denormals flushed (the default-precision intrinsics wrap each one in ~10 instructions of
// denormal scaling; per spectrogram bin that was half of the conversion's instruction count)
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_ftz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sin_ftz(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_ftz(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// forward 4-point DFT in place (w4 = -i)
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
    const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y), d13 = make_float2(a1.x - a3.x, a1.y - a3.y);
    a0 = make_float2(s02.x + s13.x, s02.y + s13.y);
    a2 = make_float2(s02.x - s13.x, s02.y - s13.y);
    a1 = make_float2(d02.x + d13.y, d02.y - d13.x);
    a3 = make_float2(d02.x - d13.y, d02.y + d13.x);
}

// forward 8-point DFT, natural order in and out
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    const float R = 0.70710678118654752f;
    float2 e0 = make_float2(v[0].x + v[4].x, v[0].y + v[4].y), o0 = make_float2(v[0].x - v[4].x, v[0].y - v[4].y);
    float2 e1 = make_float2(v[1].x + v[5].x, v[1].y + v[5].y), t1 = make_float2(v[1].x - v[5].x, v[1].y - v[5].y);
    float2 e2 = make_float2(v[2].x + v[6].x, v[2].y + v[6].y), t2 = make_float2(v[2].x - v[6].x, v[2].y - v[6].y);
    float2 e3 = make_float2(v[3].x + v[7].x, v[3].y + v[7].y), t3 = make_float2(v[3].x - v[7].x, v[3].y - v[7].y);
    float2 o1 = make_float2((t1.x + t1.y) * R, (t1.y - t1.x) * R);           // * w8
    float2 o2 = make_float2(t2.y, -t2.x);                                     // * w8^2 = -i
    float2 o3 = make_float2((t3.y - t3.x) * R, -(t3.x + t3.y) * R);          // * w8^3
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    v[0] = e0; v[2] = e1; v[4] = e2; v[6] = e3;
    v[1] = o0; v[3] = o1; v[5] = o2; v[7] = o3;
}

// 128-point forward complex FFT by one half-warp (u = lane & 15), n = 16 a + u, k = c + 8 d:
//   pass A  radix-8 over a -> c, twiddle w128^(u c)
//   pass B1 radix-4 over b1 (u = 4 b1 + b0) -> d0, twiddle w16^(b0 d0)          [thread = (c >> 1, b0)]
//   pass B2 radix-4 over b0 -> d1                                                 [thread = (c >> 1, d0)]
// in : v[a] = z[16 a + u];  out: v[2 d1 + c_lo] = Z[2 c_hi + c_lo + 8 d0 + 32 d1] with c_hi = u >> 2, d0 = u & 3.
// E: this frame's exchange buffer (EX_FRAME complex); tw[k] = exp(-2 pi i k / 256), k < 256.
// Both half-warps of a warp must call it together (full-mask __syncwarp).
__device__ __forceinline__ void fft128_halfwarp(float2 (&v)[8], float2* E, const float2* tw, int u) {
    dft8(v);
#pragma unroll
    for (int c = 1; c < 8; ++c) v[c] = cmul(v[c], tw[(2 * u * c) & 255]);
    {
        float4* row = reinterpret_cast<float4*>(E + u * EX_PITCH);
#pragma unroll
        for (int j = 0; j < 4; ++j) row[j] = make_float4(v[2 * j].x, v[2 * j].y, v[2 * j + 1].x, v[2 * j + 1].y);
    }
    __syncwarp();
    const int hi = u >> 2, lo = u & 3;
    float4 r[4];
#pragma unroll
    for (int b1 = 0; b1 < 4; ++b1) r[b1] = *reinterpret_cast<const float4*>(E + (4 * b1 + lo) * EX_PITCH + 2 * hi);
    __syncwarp();
    {
        float2 p0 = make_float2(r[0].x, r[0].y), p1 = make_float2(r[1].x, r[1].y), p2 = make_float2(r[2].x, r[2].y), p3 = make_float2(r[3].x, r[3].y);
        float2 q0 = make_float2(r[0].z, r[0].w), q1 = make_float2(r[1].z, r[1].w), q2 = make_float2(r[2].z, r[2].w), q3 = make_float2(r[3].z, r[3].w);
        dft4(p0, p1, p2, p3);
        dft4(q0, q1, q2, q3);
        const float2 w1 = tw[16 * lo], w2 = tw[32 * lo], w3 = tw[48 * lo];
        p1 = cmul(p1, w1); q1 = cmul(q1, w1); p2 = cmul(p2, w2); q2 = cmul(q2, w2); p3 = cmul(p3, w3); q3 = cmul(q3, w3);
        float2* dst = E + hi * 40 + lo * 2;                    // (c_hi, d0, b0) at c_hi * 320 + d0 * 80 + b0 * 16 bytes
        *reinterpret_cast<float4*>(dst) = make_float4(p0.x, p0.y, q0.x, q0.y);
        *reinterpret_cast<float4*>(dst + 10) = make_float4(p1.x, p1.y, q1.x, q1.y);
        *reinterpret_cast<float4*>(dst + 20) = make_float4(p2.x, p2.y, q2.x, q2.y);
        *reinterpret_cast<float4*>(dst + 30) = make_float4(p3.x, p3.y, q3.x, q3.y);
    }
    __syncwarp();
#pragma unroll
    for (int b0 = 0; b0 < 4; ++b0) r[b0] = *reinterpret_cast<const float4*>(E + hi * 40 + lo * 10 + b0 * 2);
    __syncwarp();
    {
        float2 p0 = make_float2(r[0].x, r[0].y), p1 = make_float2(r[1].x, r[1].y), p2 = make_float2(r[2].x, r[2].y), p3 = make_float2(r[3].x, r[3].y);
        float2 q0 = make_float2(r[0].z, r[0].w), q1 = make_float2(r[1].z, r[1].w), q2 = make_float2(r[2].z, r[2].w), q3 = make_float2(r[3].z, r[3].w);
        dft4(p0, p1, p2, p3);
        dft4(q0, q1, q2, q3);
        v[0] = p0; v[1] = q0; v[2] = p1; v[3] = q1; v[4] = p2; v[5] = q2; v[6] = p3; v[7] = q3;
    }
}

// atan2 to ~1e-7 rad: ratio of the smaller to the larger magnitude, Cephes atanf reduction to |t| <= tan(pi/8) and
// its degree-9 odd polynomial, then the octant fix-ups. (-0 counts as +0: numpy's real FFT returns +0 imaginary
// parts at DC / Nyquist, so negative real bins there have phase +pi.)
__device__ __forceinline__ float fast_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mn * rcp_ftz(mx);
    const bool big = a > 0.41421356237f;
    const float t = big ? (a - 1.f) * rcp_ftz(a + 1.f) : a;
    const float z = t * t;
    float r = fmaf(fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f), z * t, t);
    if (big) r += 0.78539816339745f;
    if (ay > ax) r = 1.57079632679490f - r;
    if (x < 0.f) r = 3.14159265358979f - r;
    if (!(mx > 0.f)) r = 0.f;
    return y < 0.f ? -r : r;
}

__device__ __forceinline__ float window_at(int j, int n_fft, int win_length) {
    // periodic Hann(win_length) centred in the n_fft frame (librosa pad_center)
    const int lp = (n_fft - win_length) >> 1;
    const int k = j - lp;
    if (k < 0 || k >= win_length) return 0.f;
    return 0.5f - 0.5f * cospif(2.f * (float)k / (float)win_length);
}

__device__ __forceinline__ void fill_tables(float2* tw, float* win, int win_length) {
    if (threadIdx.x < NFFT) {
        float s, c;
        sincospif(-2.f * (float)threadIdx.x / (float)NFFT, &s, &c);
        tw[threadIdx.x] = make_float2(c, s);
        win[threadIdx.x] = window_at(threadIdx.x, NFFT, win_length);
    }
}

// grid = (G, B) with the G CTAs of a sample forming one cluster; block = 256. CTA r takes the 16-frame groups
// r, r + G, ... of its sample.
__global__ void __launch_bounds__(256)
stft_ampphase_kernel(const float* __restrict__ wav, urir_stft_desc d, float* __restrict__ spec, int n_groups, int vec_ok) {
    __shared__ float2 tw[NFFT];
    __shared__ float win[NFFT];
    __shared__ __align__(16) float2 ex[FR_PER_CTA * EX_FRAME];
    __shared__ __align__(16) float outs[NFFT / 2 + 1][OUT_PITCH];
    float* seg = &outs[0][0];                // the staged waveform slice (1216 floats) is dead once the frames sit in registers
    __shared__ float red[8];
    __shared__ float part;
    __shared__ float mean_s;

    cg::cluster_group cluster = cg::this_cluster();
    const int G = gridDim.x, rank = blockIdx.x, b = blockIdx.y;
    const int T = d.n_samples;
    const float* x = wav + (size_t)b * T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    fill_tables(tw, win, d.win_length);
    // mean over the whole waveform (Loader.load: signal -= mean), one slice per CTA of the cluster
    if (d.remove_mean) {
        const int lo = (int)((long long)T * rank / G), hi = (int)((long long)T * (rank + 1) / G);
        float m = 0.f;
        for (int i = lo + tid; i < hi; i += 256) m += __ldg(x + i);
        m = warp_sum(m);
        if (lane == 0) red[warp] = m;
        __syncthreads();
        if (tid == 0) { float a = 0.f; for (int i = 0; i < 8; ++i) a += red[i]; part = a; }
        cluster.sync();
        if (tid == 0) {
            float a = 0.f;
            for (int r = 0; r < G; ++r) a += *cluster.map_shared_rank(&part, r);
            mean_s = __fdividef(a, (float)T);
        }
        cluster.sync();                    // nobody leaves (or reuses `part`) while a peer may still read it
    } else {
        if (tid == 0) mean_s = 0.f;
        __syncthreads();
    }
    const float mean = mean_s;

    const int fl = tid >> 4, u = tid & 15;
    float2* E = ex + fl * EX_FRAME;
    const int nb = d.n_bins;
    for (int g = rank; g < n_groups; g += G) {
        const int f0 = g * FR_PER_CTA;
        __syncthreads();                    // the previous group's seg / outs are no longer read
        // stage the centre-padded slice: padded index f0 * hop + i <-> sample index s0 + i
        const int s0 = f0 * HOP - NFFT / 2;
        constexpr int SEG = (FR_PER_CTA - 1) * HOP + NFFT;
        if (vec_ok && s0 >= 0 && s0 + SEG <= T) {
            const float4* x4 = reinterpret_cast<const float4*>(x + s0);
            for (int i = tid; i < SEG / 4; i += 256) {
                float4 v = __ldg(x4 + i);
                v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
                reinterpret_cast<float4*>(seg)[i] = v;
            }
        } else {
            for (int i = tid; i < SEG; i += 256) {
                int si = s0 + i;
                float v = 0.f;
                if (d.pad_mode == 1) {                 // reflect (librosa < 0.10)
                    if (si < 0) si = -si;
                    if (si >= T) si = 2 * (T - 1) - si;
                    if (si >= 0 && si < T) v = __ldg(x + si) - mean;
                } else if (si >= 0 && si < T) {
                    v = __ldg(x + si) - mean;
                }
                seg[i] = v;
            }
        }
        __syncthreads();

        const bool live = f0 + (fl & ~1) < d.n_frames;   // warp-uniform: otherwise both frames of the warp are zero padding
        // windowed frame fl as 128 complex points, 8 per thread
        float2 v[8];
        if (live) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int n = 16 * a + u;
            const float2 s = *reinterpret_cast<const float2*>(seg + fl * HOP + 2 * n);
            const float2 w = *reinterpret_cast<const float2*>(win + 2 * n);
            v[a] = make_float2(s.x * w.x, s.y * w.y);
        }
        fft128_halfwarp(v, E, tw, u);
        }
        __syncthreads();                    // every warp has read its frames out of `seg`, which `outs` aliases
        if (live) {
        {   // natural order into the frame's buffer
            const int hi = u >> 2, lo = u & 3;
#pragma unroll
            for (int d1 = 0; d1 < 4; ++d1)
                *reinterpret_cast<float4*>(E + 2 * hi + 8 * lo + 32 * d1) = make_float4(v[2 * d1].x, v[2 * d1].y, v[2 * d1 + 1].x, v[2 * d1 + 1].y);
        }
        __syncwarp();
        // untangle: X[k] = Ev[k] + w256^k Od[k], X[128 - k] = conj(Ev[k] - w256^k Od[k]); amplitude / phase, normalised
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            const int k = u + 16 * m;
            if (k <= NFFT / 4) {
                const float2 za = E[k], zb = E[(NFFT / 2 - k) & (NFFT / 2 - 1)];
                const float2 ev = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
                const float2 od = make_float2(0.5f * (za.y + zb.y), -0.5f * (za.x - zb.x));
                const float2 t = cmul(od, tw[k]);
                float2 X[2] = {make_float2(ev.x + t.x, ev.y + t.y), make_float2(ev.x - t.x, -(ev.y - t.y))};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 1 && k == NFFT / 4) break;
                    const int bin = h == 0 ? k : NFFT / 2 - k;
                    if (k == 0) X[h].y = 0.f;                         // DC / Nyquist bins are real
                    const float p2 = X[h].x * X[h].x + X[h].y * X[h].y;
                    const float amp = p2 > 1e-30f ? p2 * rsqrt_ftz(p2) : 0.f;
                    const float ph = fast_atan2(X[h].y, X[h].x);
                    float2 o;
                    // Normalizer.normalize (preprocess.py:26-32): 20 log10(v) = 6.0206 log2(v)
                    o.x = d.normalized ? fmaf(6.02059991328f, lg2_ftz(fmaf(amp, 1.f / 128.f, 1e-5f)), 100.f) * 0.01f : amp;
                    o.y = d.normalized ? fmaf(ph, 1.f / 6.28318530717959f, 0.5f) : ph;
                    *reinterpret_cast<float2*>(&outs[bin][2 * fl]) = o;
                }
            }
        }
        }
        __syncthreads();
        // rows spec[b][bin][f0 .. f0 + 16)[2] = 128 contiguous bytes; zero in the padded region (TensorPadder)
        if (vec_ok) {
            const int q = tid & 7, f = f0 + 2 * q;
            const bool z0 = f >= d.n_frames, z1 = f + 1 >= d.n_frames;
            const size_t gstride = (size_t)32 * d.W_pad * 2;
            float* gp = spec + (((size_t)b * d.H_pad + (tid >> 3)) * d.W_pad + f) * 2;
            for (int bin = tid >> 3; bin < d.H_pad; bin += 32, gp += gstride) {
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bin < nb) val = *reinterpret_cast<const float4*>(&outs[bin][4 * q]);
                if (z0) { val.x = 0.f; val.y = 0.f; }
                if (z1) { val.z = 0.f; val.w = 0.f; }
                *reinterpret_cast<float4*>(gp) = val;
            }
        } else {
            const int nfr = (d.W_pad - f0 < FR_PER_CTA) ? d.W_pad - f0 : FR_PER_CTA;
            for (int i = tid; i < d.H_pad * FR_PER_CTA * 2; i += 256) {
                const int ch = i & 1, fr = (i >> 1) & (FR_PER_CTA - 1), bin = i >> 5;
                if (fr >= nfr) continue;
                const int f = f0 + fr;
                const float val = (bin < nb && f < d.n_frames) ? outs[bin][2 * fr + ch] : 0.f;
                spec[(((size_t)b * d.H_pad + bin) * d.W_pad + f) * 2 + ch] = val;
            }
        }
    }
}

// grid = (X, B), X <= ceil((n_frames - 1) / 13) chunks; block = 256. Chunk c is the output hops [13 c, 13 c + 13): samples
// n0 = 832 c ... ; output sample n <-> padded index n + n_fft/2; frame t covers padded [t hop, t hop + n_fft), so
// exactly the 16 frames 13 c - 1 ... 13 c + 14 contribute.
__global__ void __launch_bounds__(256)
istft_ampphase_kernel(const float* __restrict__ spec, urir_stft_desc d, float* __restrict__ wav, int n_chunks, int vec_ok) {
    __shared__ float2 tw[NFFT];
    __shared__ float win[NFFT];
    __shared__ __align__(16) float2 ex[FR_PER_CTA * EX_FRAME];           // exchange buffers, then the windowed frames
    __shared__ __align__(16) float2 Xs[FR_PER_CTA][NFFT / 2 + 2];        // complex spectrum columns

    const int b = blockIdx.y, tid = threadIdx.x;
    fill_tables(tw, win, d.win_length);
    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int n0 = chunk * ISTFT_HOPS * HOP;
    const int n_out = (d.n_samples - n0 < ISTFT_HOPS * HOP) ? d.n_samples - n0 : ISTFT_HOPS * HOP;
    const int t_lo = chunk * ISTFT_HOPS - 1;
    __syncthreads();                        // the previous chunk's frames have been consumed
    // stage [bin][16 frames] tiles: un-pad + denormalise + polar -> complex (postprocess.py:87-128)
    for (int i = tid; i < d.n_bins * FR_PER_CTA; i += 256) {
        const int k = i >> 4, f = i & 15, t = t_lo + f;
        float2 v = make_float2(0.f, 0.f);
        if (t >= 0 && t < d.n_frames) {
            const float2 ap = __ldg(reinterpret_cast<const float2*>(spec + (((size_t)b * d.H_pad + k) * d.W_pad + t) * 2));
            float amp = ap.x, ph = ap.y;
            if (d.normalized) {
                amp = (ex2_ftz((ap.x * 100.f - 100.f) * (0.05f * 3.32192809489f)) - 1e-5f) * 128.f;   // 10^x = 2^(x log2 10)
                ph = ap.y * 6.28318530717959f - 3.14159265358979f;
                // (phase + pi) % 2pi - pi, python modulo
                float w = ph + 3.14159265358979f;
                w -= 6.28318530717959f * floorf(w * (1.f / 6.28318530717959f));
                ph = w - 3.14159265358979f;
            }
            const float s = sin_ftz(ph), c = cos_ftz(ph);        // |ph| <= pi after the wrap: absolute error ~5e-7
            v = make_float2(amp * c, amp * s);
            if (k == 0 || k == NFFT / 2) v.y = 0.f;            // C2R ignores these imaginary parts
        }
        Xs[f][k] = v;
    }
    __syncthreads();

    const int fl = tid >> 4, u = tid & 15;
    float2* E = ex + fl * EX_FRAME;
    // Z[k] = Ev[k] + i Od[k], Ev = (X[k] + conj X[128-k]) / 2, Od = conj(w256^k) (X[k] - conj X[128-k]) / 2;
    // z = IDFT128(Z) = conj(DFT128(conj Z)) / 128, x[2m] = Re z[m], x[2m+1] = Im z[m]
    float2 v[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int k = 16 * a + u;
        const float2 xa = Xs[fl][k], xb = Xs[fl][NFFT / 2 - k];
        const float2 ev = make_float2(0.5f * (xa.x + xb.x), 0.5f * (xa.y - xb.y));
        const float2 df = make_float2(0.5f * (xa.x - xb.x), 0.5f * (xa.y + xb.y));
        const float2 w = tw[k];
        const float2 od = make_float2(df.x * w.x + df.y * w.y, df.y * w.x - df.x * w.y);      // df * conj(w)
        v[a] = make_float2(ev.x - od.y, -(ev.y + od.x));                                      // conj(Ev + i Od)
    }
    fft128_halfwarp(v, E, tw, u);
    {
        const int hi = u >> 2, lo = u & 3;
        float* fr = reinterpret_cast<float*>(E);                 // 256 windowed samples of this frame
#pragma unroll
        for (int d1 = 0; d1 < 4; ++d1) {
            const int j0 = 4 * hi + 16 * lo + 64 * d1;
            const float4 w = *reinterpret_cast<const float4*>(win + j0);
            const float sc = 1.f / (NFFT / 2);
            *reinterpret_cast<float4*>(fr + j0) = make_float4(v[2 * d1].x * sc * w.x, -v[2 * d1].y * sc * w.y,
                                                              v[2 * d1 + 1].x * sc * w.z, -v[2 * d1 + 1].y * sc * w.w);
        }
    }
    __syncthreads();

    // overlap-add with window-sum-square normalisation (librosa.istft), four samples per thread
    const float* frames = reinterpret_cast<const float*>(ex);
    constexpr int FR_FLOATS = EX_FRAME * 2;
    for (int i4 = tid * 4; i4 < n_out; i4 += 256 * 4) {
        const int p = n0 + NFFT / 2 + i4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), wss = make_float4(0.f, 0.f, 0.f, 0.f);
        const int tb = p / HOP;
#pragma unroll
        for (int q = 0; q < NFFT / HOP; ++q) {
            const int t = tb - q;
            if (t >= 0 && t < d.n_frames) {
                const int j = p - t * HOP;
                const float4 f = *reinterpret_cast<const float4*>(frames + (t - t_lo) * FR_FLOATS + j);
                const float4 w = *reinterpret_cast<const float4*>(win + j);
                acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
                wss.x = fmaf(w.x, w.x, wss.x); wss.y = fmaf(w.y, w.y, wss.y); wss.z = fmaf(w.z, w.z, wss.z); wss.w = fmaf(w.w, w.w, wss.w);
            }
        }
        if (wss.x > 1.17549435e-38f) acc.x *= rcp_ftz(wss.x);
        if (wss.y > 1.17549435e-38f) acc.y *= rcp_ftz(wss.y);
        if (wss.z > 1.17549435e-38f) acc.z *= rcp_ftz(wss.z);
        if (wss.w > 1.17549435e-38f) acc.w *= rcp_ftz(wss.w);
        float* o = wav + (size_t)b * d.n_samples + n0 + i4;
        if (vec_ok) *reinterpret_cast<float4*>(o) = acc;
        else { o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w; }
    }
    }
}

static int check_desc(const urir_stft_desc* d) {
    URIR_CHECK_ARG(d != nullptr, "stft: null descriptor");
    URIR_CHECK_ARG(d->n_fft == NFFT, "stft: only n_fft=256 is built (dataset.py:62)");
    URIR_CHECK_ARG(d->hop_length == HOP, "stft: hop_length must be 64 (dataset.py:64)");
    URIR_CHECK_ARG(d->win_length > 0 && d->win_length <= NFFT && d->win_length % 2 == 0, "stft: bad win_length");
    URIR_CHECK_ARG(d->n_bins == NFFT / 2 + 1, "stft: n_bins must be n_fft/2+1");
    URIR_CHECK_ARG(d->n_frames == 1 + d->n_samples / d->hop_length, "stft: n_frames must be 1 + n_samples/hop");
    URIR_CHECK_ARG(d->H_pad >= d->n_bins && d->H_pad <= 160 && d->W_pad >= d->n_frames, "stft: padded shape too small");
    return URIR_OK;
}

int stft_ampphase(const float* wav, int B, const urir_stft_desc* d, float* spec, cudaStream_t st) {
    int rc = check_desc(d); if (rc) return rc;
    URIR_CHECK_ARG(B > 0 && B <= 65535, "stft: B must be in 1..65535");
    const int n_groups = cdiv(d->W_pad, FR_PER_CTA);
    // CTAs per sample (= cluster size): minimise (waves of resident CTAs) x (frame groups one CTA works through, plus
    // about half a group of per-CTA cost: tables, mean, two cluster barriers). 10 groups at B = 256 -> 5 CTAs of 2 groups,
    // not 8 with two of them doing double work; a handful of samples -> 8 CTAs each.
    int G = 1;
    {
        const long long slots = 5LL * sm_count();
        const int gmax = n_groups < MAX_CLUSTER ? n_groups : MAX_CLUSTER;
        long long best = -1;
        for (int g = 1; g <= gmax; ++g) {
            const long long waves = ((long long)B * g + slots - 1) / slots;
            const long long cost = waves * (2 * cdiv(n_groups, g) + 1);
            if (best < 0 || cost < best) { best = cost; G = g; }
        }
    }
    // 128-bit paths need 16-byte aligned rows in both tensors
    const int vec_ok = (d->W_pad % FR_PER_CTA == 0) && (d->n_samples % 4 == 0) &&
                       ((uintptr_t)wav % 16 == 0) && ((uintptr_t)spec % 16 == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G, B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    URIR_CUDA_OK(cudaLaunchKernelEx(&cfg, stft_ampphase_kernel, wav, *d, spec, n_groups, vec_ok));
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int istft_from_ampphase(const float* spec, int B, const urir_stft_desc* d, float* wav, cudaStream_t st) {
    int rc = check_desc(d); if (rc) return rc;
    URIR_CHECK_ARG(B > 0 && B <= 65535, "istft: B must be in 1..65535");
    URIR_CHECK_ARG(d->n_samples == d->hop_length * (d->n_frames - 1), "istft: n_samples must be hop*(n_frames-1)");
    const int vec_ok = ((uintptr_t)wav % 16 == 0);
    const int n_chunks = cdiv(d->n_frames - 1, ISTFT_HOPS);
    // one CTA per chunk: fewer CTAs looping over several chunks measured slower (B = 256: 61 against 54 us) -- the phases
    // of a chunk are strictly serial, so it is the number of resident CTAs that hides their latencies
    const int X = n_chunks;
    dim3 grid(X, B);
    istft_ampphase_kernel<<<grid, 256, 0, st>>>(spec, *d, wav, n_chunks, vec_ok);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
