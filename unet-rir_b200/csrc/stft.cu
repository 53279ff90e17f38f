// Signal path on the GPU: replaces the per-file librosa/numpy CPU code of
//   preprocess.py:13-18 (librosa.stft, abs, angle), :26-32 (normalise), :70-105 (zero-pad),
//   :56 (mean removal)                                     -> stft_ampphase_kernel
//   postprocess.py:87-113 (un-pad, denormalise), :128-129 (polar -> complex, librosa.istft)
//                                                          -> istft_ampphase_kernel
// Both are HBM-bound (38.4 kB + 184 kB per sample). One CTA owns 16
// consecutive frames (STFT) or 16 hops of output (iSTFT): the waveform slice / spectrum columns
// are staged in shared memory, each warp runs radix-2 256-point FFTs in shared memory, and the
// result is written with row-contiguous 128-byte runs.
#include "urir_common.cuh"

namespace urir {

constexpr int NFFT = 256;
constexpr int LOG_NFFT = 8;
constexpr int FR_PER_CTA = 16;

__device__ __forceinline__ int bitrev8(int v) { return (int)(__brev((unsigned)v) >> 24); }

// in-place radix-2 DIT FFT of 256 complex points held in `buf` (bit-reversed order on entry);
// executed by one warp. tw[k] = exp(-2*pi*i*k/256), k < 128. inverse => conjugated twiddles.
__device__ __forceinline__ void warp_fft256(float2* buf, const float2* tw, bool inverse) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < LOG_NFFT; ++s) {
        const int half = 1 << s;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int bf = lane + 32 * j;                 // butterfly id 0..127
            const int grp = bf >> s, pos = bf & (half - 1);
            const int i0 = (grp << (s + 1)) + pos, i1 = i0 + half;
            float2 w = tw[pos << (LOG_NFFT - 1 - s)];
            if (inverse) w.y = -w.y;
            const float2 a = buf[i0], b = buf[i1];
            const float2 t = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
            buf[i0] = make_float2(a.x + t.x, a.y + t.y);
            buf[i1] = make_float2(a.x - t.x, a.y - t.y);
        }
        __syncwarp();
    }
}

__device__ __forceinline__ float window_at(int j, int n_fft, int win_length) {
    // periodic Hann(win_length) centred in the n_fft frame (librosa pad_center)
    const int lp = (n_fft - win_length) >> 1;
    const int k = j - lp;
    if (k < 0 || k >= win_length) return 0.f;
    return 0.5f - 0.5f * cospif(2.f * (float)k / (float)win_length);
}

// grid = (ceil(W_pad / 16), B); block = 256
__global__ void __launch_bounds__(256)
stft_ampphase_kernel(const float* __restrict__ wav, urir_stft_desc d, float* __restrict__ spec) {
    __shared__ float2 tw[NFFT / 2];
    __shared__ float2 work[8][NFFT];
    __shared__ float seg[(FR_PER_CTA - 1) * 64 + NFFT];       // hop <= 64 assumed by host check
    __shared__ float outs[FR_PER_CTA][2][160];                 // [frame][amp|phase][bin] (bins <= 160)
    __shared__ float red[8];
    __shared__ float mean_s;

    const int b = blockIdx.y, f0 = blockIdx.x * FR_PER_CTA;
    const int T = d.n_samples, hop = d.hop_length;
    const float* x = wav + (size_t)b * T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x < NFFT / 2) {
        float s, c;
        sincospif(-2.f * (float)threadIdx.x / (float)NFFT, &s, &c);
        tw[threadIdx.x] = make_float2(c, s);
    }
    // mean over the whole waveform (Loader.load: signal -= mean)
    float m = 0.f;
    if (d.remove_mean) {
        for (int i = threadIdx.x; i < T; i += 256) m += __ldg(x + i);
        m = warp_sum(m);
        if (lane == 0) red[warp] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        if (d.remove_mean) { for (int i = 0; i < 8; ++i) a += red[i]; a /= (float)T; }
        mean_s = a;
    }
    __syncthreads();
    const float mean = mean_s;

    // stage the centre-padded slice: padded index pi = f0*hop + i ; sample index si = pi - n_fft/2
    const int seg_len = (FR_PER_CTA - 1) * hop + NFFT;
    for (int i = threadIdx.x; i < seg_len; i += 256) {
        int si = f0 * hop + i - NFFT / 2;
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
    __syncthreads();

    const int nb = d.n_bins;
    for (int fl = warp; fl < FR_PER_CTA; fl += 8) {
        const int f = f0 + fl;
        if (f < d.n_frames) {
            float2* buf = work[warp];
            for (int j = lane; j < NFFT; j += 32)
                buf[bitrev8(j)] = make_float2(seg[fl * hop + j] * window_at(j, NFFT, d.win_length), 0.f);
            __syncwarp();
            warp_fft256(buf, tw, false);
            for (int k = lane; k < nb; k += 32) {
                const float2 v = buf[k];
                const float amp = sqrtf(v.x * v.x + v.y * v.y);
                const float ph = atan2f(v.y, v.x);
                // Normalizer.normalize (preprocess.py:26-32)
                outs[fl][0][k] = d.normalized ? (20.f * log10f(amp * (1.f / 128.f) + 1e-5f) + 100.f) * 0.01f : amp;
                outs[fl][1][k] = d.normalized ? (ph + 3.14159265358979f) * (1.f / 6.28318530717959f) : ph;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // write rows: spec[b][bin][f0..f0+16)[2], zero in the padded region (TensorPadder)
    const int nfr = (d.W_pad - f0 < FR_PER_CTA) ? d.W_pad - f0 : FR_PER_CTA;
    for (int i = threadIdx.x; i < d.H_pad * FR_PER_CTA * 2; i += 256) {
        const int ch = i & 1, fl = (i >> 1) % FR_PER_CTA, bin = i / (2 * FR_PER_CTA);
        if (fl >= nfr) continue;
        const int f = f0 + fl;
        const float v = (bin < nb && f < d.n_frames) ? outs[fl][ch][bin] : 0.f;
        spec[(((size_t)b * d.H_pad + bin) * d.W_pad + f) * 2 + ch] = v;
    }
}

// grid = (ceil(n_samples / (16*hop)), B); block = 256
// output sample n <-> padded index n + n_fft/2 ; frame t covers [t*hop, t*hop + n_fft)
constexpr int ISTFT_MAX_FR = FR_PER_CTA + 4;      // frames overlapping 16 hops when n_fft = 4*hop
__global__ void __launch_bounds__(256)
istft_ampphase_kernel(const float* __restrict__ spec, urir_stft_desc d, float* __restrict__ wav) {
    __shared__ float2 tw[NFFT / 2];
    __shared__ float2 work[8][NFFT];
    __shared__ float frames[ISTFT_MAX_FR][NFFT];

    const int b = blockIdx.y, hop = d.hop_length;
    const int n0 = blockIdx.x * FR_PER_CTA * hop;                 // first output sample of the CTA
    const int n_out = (d.n_samples - n0 < FR_PER_CTA * hop) ? d.n_samples - n0 : FR_PER_CTA * hop;
    const int p0 = n0 + NFFT / 2, p1 = p0 + n_out;                // padded index range [p0, p1)
    int t_lo = (p0 - NFFT + 1 + hop - 1) / hop; if (p0 - NFFT + 1 < 0) t_lo = 0;
    int t_hi = (p1 - 1) / hop; if (t_hi > d.n_frames - 1) t_hi = d.n_frames - 1;
    const int nt = t_hi - t_lo + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x < NFFT / 2) {
        float s, c;
        sincospif(-2.f * (float)threadIdx.x / (float)NFFT, &s, &c);
        tw[threadIdx.x] = make_float2(c, s);
    }
    __syncthreads();

    for (int tl = warp; tl < nt; tl += 8) {
        const int t = t_lo + tl;
        float2* buf = work[warp];
        // un-pad + denormalise + polar->complex (postprocess.py:87-128), Hermitian-extend
        for (int k = lane; k <= NFFT / 2; k += 32) {
            float2 v = make_float2(0.f, 0.f);
            if (k < d.n_bins) {
                const float2 ap = __ldg(reinterpret_cast<const float2*>(spec + (((size_t)b * d.H_pad + k) * d.W_pad + t) * 2));
                float amp = ap.x, ph = ap.y;
                if (d.normalized) {
                    amp = (exp10f((ap.x * 100.f - 100.f) * 0.05f) - 1e-5f) * 128.f;
                    ph = ap.y * 6.28318530717959f - 3.14159265358979f;
                    // (phase + pi) % 2pi - pi, python modulo
                    float w = ph + 3.14159265358979f;
                    w -= 6.28318530717959f * floorf(w * (1.f / 6.28318530717959f));
                    ph = w - 3.14159265358979f;
                }
                float s, c;
                sincosf(ph, &s, &c);
                v = make_float2(amp * c, amp * s);
            }
            if (k == 0 || k == NFFT / 2) v.y = 0.f;            // C2R ignores these imaginary parts
            buf[bitrev8(k)] = v;
            if (k > 0 && k < NFFT / 2) buf[bitrev8(NFFT - k)] = make_float2(v.x, -v.y);
        }
        __syncwarp();
        warp_fft256(buf, tw, true);
        for (int j = lane; j < NFFT; j += 32)
            frames[tl][j] = buf[j].x * (1.f / NFFT) * window_at(j, NFFT, d.win_length);
        __syncwarp();
    }
    __syncthreads();

    for (int i = threadIdx.x; i < n_out; i += 256) {
        const int p = p0 + i;
        float acc = 0.f, wss = 0.f;
        int ta = (p - NFFT + 1 + hop - 1) / hop; if (p - NFFT + 1 < 0) ta = 0;
        int tb = p / hop; if (tb > d.n_frames - 1) tb = d.n_frames - 1;
        for (int t = ta; t <= tb; ++t) {
            const int j = p - t * hop;
            const float w = window_at(j, NFFT, d.win_length);
            acc += frames[t - t_lo][j];
            wss = fmaf(w, w, wss);
        }
        if (wss > 1.17549435e-38f) acc /= wss;
        wav[(size_t)b * d.n_samples + n0 + i] = acc;
    }
}

static int check_desc(const urir_stft_desc* d) {
    URIR_CHECK_ARG(d != nullptr, "stft: null descriptor");
    URIR_CHECK_ARG(d->n_fft == NFFT, "stft: only n_fft=256 is built (dataset.py:62)");
    URIR_CHECK_ARG(d->hop_length == 64, "stft: hop_length must be 64 (dataset.py:64)");
    URIR_CHECK_ARG(d->win_length > 0 && d->win_length <= NFFT && d->win_length % 2 == 0, "stft: bad win_length");
    URIR_CHECK_ARG(d->n_bins == NFFT / 2 + 1, "stft: n_bins must be n_fft/2+1");
    URIR_CHECK_ARG(d->n_frames == 1 + d->n_samples / d->hop_length, "stft: n_frames must be 1 + n_samples/hop");
    URIR_CHECK_ARG(d->H_pad >= d->n_bins && d->H_pad <= 160 && d->W_pad >= d->n_frames, "stft: padded shape too small");
    return URIR_OK;
}

int stft_ampphase(const float* wav, int B, const urir_stft_desc* d, float* spec, cudaStream_t st) {
    int rc = check_desc(d); if (rc) return rc;
    URIR_CHECK_ARG(B > 0, "stft: B must be positive");
    dim3 grid(cdiv(d->W_pad, FR_PER_CTA), B);
    stft_ampphase_kernel<<<grid, 256, 0, st>>>(wav, *d, spec);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

int istft_from_ampphase(const float* spec, int B, const urir_stft_desc* d, float* wav, cudaStream_t st) {
    int rc = check_desc(d); if (rc) return rc;
    URIR_CHECK_ARG(B > 0, "istft: B must be positive");
    URIR_CHECK_ARG(d->n_samples == d->hop_length * (d->n_frames - 1), "istft: n_samples must be hop*(n_frames-1)");
    dim3 grid(cdiv(d->n_samples, FR_PER_CTA * d->hop_length), B);
    istft_ampphase_kernel<<<grid, 256, 0, st>>>(spec, *d, wav);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
