// Thin-channel convolutions (sm_100a): the layers where ONE side has 2 channels --
//   * the stem  enc1.down : Conv2D(2 -> F0, k)            (dl_models/u_net.py:269-276 on the spectrogram)
//   * the head            : Conv2D(F0 -> 2, 6x6) + sigmoid (dl_models/u_net.py:247-249)
// A (tap, 2-channel) im2col row of the thin tensor is only 18 / 72 values, so instead of running taps as
// GEMM-K iterations over re-fetched activation tiles (what conv_igemm.cu does for wide layers), the CTA
// builds the im2col rows of a 128-pixel tile in shared memory from a halo patch of the fp32 thin tensor
// (8 bytes per pixel) and the whole layer becomes ONE small GEMM per tile:
//   thin_gemm  (stem fprop, head dgrad): wide[p, 32] = im2col(thin)[p, J] * Wm[J, 32]   M = pixels, K = J
//   thin_wgrad (stem wgrad, head wgrad): dW[J, 32]  += im2col(thin)[p, J]^T * wide[p, 32]   K = pixels
// Both read / write the wide bf16 tensor exactly once (HBM-bound, ~94 MB per launch at B = 64) and share the
// same shared-memory image of the im2col tile: 128 rows (pixels) of 128 B, 16-byte chunks XOR-swizzled with
// (row % 8) -- consumed K-major by thin_gemm and MN-major by thin_wgrad.
// Two implementations of each product live here. The tcgen05 kernels described above (thin_gemm_kernel, thin_wgrad_kernel) take
// any tap count up to 36; the square 3x3 / 6x6 kernels of the model go to the warp-level MMA kernels further down
// (thin_expand_mma_kernel, thin_wgrad_mma_kernel), which never stage im2col rows and are 1.5 - 2.6x faster on these
// HBM-bound launches (URIR_THIN_UMMA=1 selects the tcgen05 ones for everything).
// The head's forward (wide -> thin) lives in conv_head.cu.
#include <stdlib.h>
#include <atomic>
#include "urir_common.cuh"
#include "urir_tc.cuh"

namespace urir {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes);

constexpr int TH_BW = 16, TH_BH = 8;            // the 128-pixel tile
constexpr int TH_MAX_TAPS = 40;
constexpr int TH_ATOM = 128 * 128;              // 128 rows x 128 B
constexpr int TH_PATCH_MAX = (TH_BH + 5) * (TH_BW + 5);   // k <= 6

struct ThinParams {
    int N, H, W;
    int ntaps;
    int PH, PW;                 // halo patch extent (pixels)
    int oh0, ow0;               // patch origin relative to the tile origin
    int tiles_w, tiles_h, total_tiles;
    int KS;                     // GEMM-K steps of 16 for thin_gemm = ceil(2 * ntaps / 16)
    int nchunks;                // 16-byte chunks per im2col row that are (re)written per tile = 2 * KS
    int CW_total;               // channels of the wide side in the weight tensor
    int thin_is_x;              // 1: thin tensor is the conv input (stem); 0: it is the conv output side (head)
    const float* thin; int thin_ld, thin_coff;
    const __nv_bfloat16* wide; int wide_ld, wide_coff;   // thin_wgrad_mma: the 32-channel-group side, read with plain copies
    const __nv_bfloat16* w_ck; const float* bias;
    __nv_bfloat16* out; int out_ld, out_coff;
    float* dw;
    unsigned int* gate;         // thin_wgrad, deterministic mode: CTAs add into dw in blockIdx order (urir_common.cuh)
    unsigned long long mag_w, mag_h;   // ceil(2^40 / tiles_w), ceil(2^40 / tiles_h): exact division of tile indices < 2^20
    short toff[TH_MAX_TAPS];    // patch pixel index offset of tap t
};

__device__ __forceinline__ int thin_widx(const ThinParams& p, int tap, int ct, int cw) {
    // element (tap, ct, cw) in the [tap][C][K] weight layout
    return p.thin_is_x ? (tap * 2 + ct) * p.CW_total + cw : (tap * p.CW_total + cw) * 2 + ct;
}

// (runtime integer divisions were a quarter of this kernel's instructions: tile indices are < 2^20, so
// q = (tile * ceil(2^40 / d)) >> 40 is exact)
__device__ __forceinline__ void thin_tile_coords(const ThinParams& p, int tile, int& n, int& h0, int& w0) {
    const unsigned q1 = (unsigned)(((unsigned long long)(unsigned)tile * p.mag_w) >> 40);      // tile / tiles_w
    const int tw = tile - (int)q1 * p.tiles_w;
    const unsigned q2 = (unsigned)(((unsigned long long)q1 * p.mag_h) >> 40);                  // ... / tiles_h
    const int th = (int)q1 - (int)q2 * p.tiles_h;
    n = (int)q2;
    h0 = th * TH_BH; w0 = tw * TH_BW;
}

// this thread's (up to 3) patch entries as (row << 16 | column), -1 beyond the patch: tile independent, computed once
__device__ __forceinline__ void thin_patch_coords(const ThinParams& p, int (&pyx)[3]) {
    const int npatch = p.PH * p.PW;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int i = threadIdx.x + 128 * q;
        const int py = i / p.PW, px = i - py * p.PW;
        pyx[q] = i < npatch ? ((py << 16) | px) : -1;
    }
}

// this thread's (up to 3) entries of the halo patch of `tile`, zero outside the image (SAME padding)
__device__ __forceinline__ void thin_load_patch(const ThinParams& p, int tile, const int (&pyx)[3], float2 (&pr)[3]) {
    int n, h0, w0; thin_tile_coords(p, tile, n, h0, w0);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        pr[q] = make_float2(0.f, 0.f);
        if (pyx[q] >= 0) {
            const int gh = h0 + p.oh0 + (pyx[q] >> 16), gw = w0 + p.ow0 + (pyx[q] & 0xffff);
            if (gh >= 0 && gh < p.H && gw >= 0 && gw < p.W)
                pr[q] = __ldg(reinterpret_cast<const float2*>(p.thin + ((size_t)(n * p.H + gh) * p.W + gw) * p.thin_ld + p.thin_coff));
        }
    }
}
__device__ __forceinline__ void thin_store_patch(const ThinParams& p, float2* patch, const float2 (&pr)[3]) {
    const int npatch = p.PH * p.PW;
#pragma unroll
    for (int q = 0; q < 3; ++q) { const int i = threadIdx.x + 128 * q; if (i < npatch) patch[i] = pr[q]; }
}

// im2col row of pixel `m` (= threadIdx.x) -> 16-byte chunks c = 0 .. nchunks-1 of the swizzled tile
// KSZ = 3 / 6: square kernel of that size with compile-time tap offsets (FLIP: the mirrored taps of an input gradient),
// so every tap is one shared-memory load with an immediate offset; KSZ = 0: any geometry through p.toff / p.ntaps
template <int KSZ, bool FLIP>
__device__ __forceinline__ void thin_build_row(const ThinParams& p, const float2* patch, uint8_t* sA) {
    const int m = threadIdx.x;
    if constexpr (KSZ > 0) {
        constexpr int NT = KSZ * KSZ, PW = TH_BW + KSZ - 1, NCH = 2 * ((2 * NT + 15) / 16);
        const float2* src = patch + (m / TH_BW) * PW + (m % TH_BW);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            uint32_t wd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int tap = 4 * c + e;
                wd[e] = 0u;
                if (tap < NT) {
                    const int r = tap / KSZ, sx = tap % KSZ;
                    const float2 v = src[FLIP ? (KSZ - 1 - r) * PW + (KSZ - 1 - sx) : r * PW + sx];
                    wd[e] = pack_bf16x2(v.x, v.y);
                }
            }
            *reinterpret_cast<uint4*>(sA + (c >> 3) * TH_ATOM + m * 128 + (((c & 7) ^ (m & 7)) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
        return;
    }
    const int base = (m / TH_BW) * p.PW + (m % TH_BW);
#pragma unroll
    for (int c = 0; c < TH_MAX_TAPS / 4; ++c) {
        if (c < p.nchunks) {
            uint32_t wd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int tap = 4 * c + e;
                float2 v = make_float2(0.f, 0.f);
                if (tap < p.ntaps) v = patch[base + p.toff[tap]];
                wd[e] = pack_bf16x2(v.x, v.y);
            }
            *reinterpret_cast<uint4*>(sA + (c >> 3) * TH_ATOM + m * 128 + (((c & 7) ^ (m & 7)) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// thin_gemm: out[p, n0 .. n0+32) = im2col(thin)[p, :] * Wm (+ bias)
// 128 threads: thread = pixel = A row = TMEM lane. Persistent over tiles; 4 CTAs per SM hide each other's
// global-load / MMA / store latencies. smem: A 2 atoms (32 KB) | B 2 atoms of 32 rows (8 KB) | patch | barrier
// ---------------------------------------------------------------------------------------------
constexpr int TG_B_ATOM = 32 * 128;
constexpr int TG_SMEM = 2 * TH_ATOM + 2 * TG_B_ATOM + TH_PATCH_MAX * 8 + 64 + 1024;

template <int KSZ, bool FLIP>
__global__ void __launch_bounds__(128)
thin_gemm_kernel(const __grid_constant__ ThinParams p) {
    constexpr uint32_t IDESC = make_idesc_bf16(128, 32, 0, 0);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // one 128-byte atom row holds 32 taps: the stem (9 / 36 taps) and a 3x3 head need one atom, the 6x6 head two
    const int atoms = p.nchunks > 8 ? 2 : 1;
    uint8_t* sA = smem;
    uint8_t* sB = smem + atoms * TH_ATOM;
    float2* patch = reinterpret_cast<float2*>(sB + atoms * TG_B_ATOM);
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(patch) + TH_PATCH_MAX * 8);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * 32;

    if (threadIdx.x == 0) { mbar_init(mma_bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tmem_slot, 32);
    // weights -> swizzled K-major B tile: row n = wide channel, column j = (tap, thin channel)
    for (int idx = threadIdx.x; idx < 32 * p.nchunks; idx += 128) {
        const int c = idx >> 5, n = idx & 31;
        uint32_t wd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int tap = 4 * c + e;
            float a = 0.f, b = 0.f;
            if (tap < p.ntaps) { a = bf2f(p.w_ck[thin_widx(p, tap, 0, n0 + n)]); b = bf2f(p.w_ck[thin_widx(p, tap, 1, n0 + n)]); }
            wd[e] = pack_bf16x2(a, b);
        }
        *reinterpret_cast<uint4*>(sB + (c >> 3) * TG_B_ATOM + n * 128 + (((c & 7) ^ (n & 7)) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (SWZ_128B << 29);
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);

    float bias_v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_v[j] = p.bias ? __ldg(p.bias + n0 + j) : 0.f;

    float2 pr[3];
    int pyx[3]; thin_patch_coords(p, pyx);
    int tile = blockIdx.x;
    if (tile < p.total_tiles) thin_load_patch(p, tile, pyx, pr);
    for (int it = 0; tile < p.total_tiles; ++it, tile += gridDim.x) {
        thin_store_patch(p, patch, pr);
        __syncthreads();
        thin_build_row<KSZ, FLIP>(p, patch, sA);
        const int next = tile + gridDim.x;
        if (next < p.total_tiles) thin_load_patch(p, next, pyx, pr);      // in flight across the MMA and the epilogue
        fence_proxy_async();
        fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            fence_after_sync();
            for (int i = 0; i < p.KS; ++i) {
                const uint64_t ad = ((uint64_t)d_hi << 32) | ((a_addr + (i >> 2) * TH_ATOM + (i & 3) * 32) >> 4);
                const uint64_t bd = ((uint64_t)d_hi << 32) | ((b_addr + (i >> 2) * TG_B_ATOM + (i & 3) * 32) >> 4);
                umma_bf16_elect(tmem_base, ad, bd, IDESC, i != 0);
            }
            umma_commit_elect(mma_bar);
            __syncwarp();
        }
        mbar_wait(mma_bar, it & 1);
        fence_after_sync();
        int n, h0, w0; thin_tile_coords(p, tile, n, h0, w0);
        const int m = threadIdx.x;
        __nv_bfloat16* orow = p.out + ((size_t)(n * p.H + h0 + m / TH_BW) * p.W + w0 + m % TH_BW) * p.out_ld + p.out_coff + n0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t r0[16], r1[16];
        tmem_ld16(lane_addr, r0);
        tmem_ld16(lane_addr + 16, r1);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j] = pack_bf16x2(__uint_as_float(r0[2 * j]) + bias_v[2 * j], __uint_as_float(r0[2 * j + 1]) + bias_v[2 * j + 1]);
            o[8 + j] = pack_bf16x2(__uint_as_float(r1[2 * j]) + bias_v[16 + 2 * j], __uint_as_float(r1[2 * j + 1]) + bias_v[16 + 2 * j + 1]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(orow + 8 * q) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        fence_before_sync();
    }
    __syncthreads();
    if (warp == 0) { fence_after_sync(); tmem_dealloc(tmem_base, 32); }
    (void)lane;
}

// ---------------------------------------------------------------------------------------------
// thin_expand_mma: the same product as thin_gemm with the im2col rows never materialised.
// These two launches (stem fprop, head dgrad) move 106 MB for 1.7 / 6.8 GFLOP: they are HBM-bound by two orders of
// magnitude, and the tcgen05 version above spent its time building the im2col tile in shared memory, waiting for one tiny
// UMMA per tile and reading TMEM back (41 / 61 us against a 16 us traffic floor, tensor pipe 3-6 % busy). With a
// K index of (tap, thin channel) a warp-level m16n8k16 A fragment register IS one halo-patch pixel -- the two channels
// of pixel + tap as a bf16 pair -- so a warp takes a 16-pixel row segment, reads its A operand straight out of the
// bf16 halo patch (4 shared-memory words per K step, no staging of the 18 / 72-wide rows), keeps the whole weight
// matrix as B fragments in registers (MMA columns permuted so that a lane's accumulators are 8 consecutive channels:
// 16-byte stores, 512 contiguous bytes per warp instruction). 8 warps = the 8 x 16 tile; the halo patches of the next
// three tiles are in flight as cp.async copies into a shared-memory ring (one tile's iteration is a few hundred cycles,
// a DRAM round trip several times that: with a one-tile prefetch the kernel ran at the load latency); two CTAs per SM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int TE_NBUF = 4;             // patch ring: tiles in flight per CTA (each tile's iteration is far shorter than a DRAM round trip)

__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gsrc, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int KSZ, bool FLIP>
__global__ void __launch_bounds__(256, 2)
thin_expand_mma_kernel(const __grid_constant__ ThinParams p) {
    constexpr int NT = KSZ * KSZ, PW = TH_BW + KSZ - 1, PH = TH_BH + KSZ - 1, KS = (2 * NT + 15) / 16, NP = PH * PW;
    static_assert(NP <= 512, "two patch pixels per thread");
    __shared__ float2 patch[TE_NBUF][NP];                   // (channel 0, channel 1) of a halo-patch pixel, fp32 as in HBM
    __shared__ uint32_t patch16[2][NP];                     // the current tile's patch as bf16 pairs = MMA A-fragment registers
    pdl_sync();      // lets the successor's prologue start; the weights read next may come from the immediately preceding launch (operand refresh, BN fold)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.y * 32;

    // B fragments of the whole [2 NT x 32] weight matrix: K step ks, 8-channel group nt, halves h (taps 8 ks + t + 4 h)
    uint32_t bfrag[KS][4][2];
    int aoff[KS][2];                                        // patch offset of this thread's two taps per K step, -1 = padding
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int tap = 8 * ks + t + 4 * h;
            const int r = tap / KSZ, sx = tap - r * KSZ;
            aoff[ks][h] = tap < NT ? (FLIP ? (KSZ - 1 - r) * PW + (KSZ - 1 - sx) : r * PW + sx) : -1;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                uint32_t v = 0u;
                if (tap < NT) {
                    const int n = n0 + 8 * (g >> 1) + 2 * nt + (g & 1);   // MMA column (nt, g) = this channel: see the store
                    const uint32_t lo = __bfloat16_as_ushort(p.w_ck[thin_widx(p, tap, 0, n)]);
                    const uint32_t hi = __bfloat16_as_ushort(p.w_ck[thin_widx(p, tap, 1, n)]);
                    v = lo | (hi << 16);
                }
                bfrag[ks][nt][h] = v;
            }
        }
    float bias2[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        bias2[nt][0] = p.bias ? __ldg(p.bias + n0 + 8 * t + 2 * nt) : 0.f;
        bias2[nt][1] = p.bias ? __ldg(p.bias + n0 + 8 * t + 2 * nt + 1) : 0.f;
    }
    // this thread's two patch pixels (tile independent)
    int py[2], px[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int i = threadIdx.x + 256 * q;
        py[q] = i < NP ? i / PW : -1; px[q] = i - (i / PW) * PW;
    }
    // asynchronous copy of the halo patch of `tile` into ring slot `buf` (zero fill outside the image = SAME padding);
    // always one commit group per call so that the wait counts stay uniform
    auto fetch_patch = [&](int tile, int buf) {
        if (tile < p.total_tiles) {
            int n, h0, w0; thin_tile_coords(p, tile, n, h0, w0);
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (py[q] >= 0) {
                    const int gh = h0 + p.oh0 + py[q], gw = w0 + p.ow0 + px[q];
                    const bool in = gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
                    const float* src = p.thin + ((size_t)(n * p.H + (in ? gh : 0)) * p.W + (in ? gw : 0)) * p.thin_ld + p.thin_coff;
                    cp_async8_zfill(&patch[buf][threadIdx.x + 256 * q], src, in);
                }
        }
        cp_async_commit();
    };

    int tile = blockIdx.x;
#pragma unroll
    for (int i = 0; i < TE_NBUF - 1; ++i) fetch_patch(tile + i * (int)gridDim.x, i);
    for (int it = 0; tile < p.total_tiles; tile += gridDim.x, ++it) {
        cp_async_wait<TE_NBUF - 2>();                                // this thread's part of tile `it` has landed:
#pragma unroll
        for (int q = 0; q < 2; ++q)                                  // ... convert it once (the taps re-read every pixel up to 36 times)
            if (py[q] >= 0) {
                const float2 v = patch[it % TE_NBUF][threadIdx.x + 256 * q];
                patch16[it & 1][threadIdx.x + 256 * q] = pack_bf16x2(v.x, v.y);
            }
        __syncthreads();                                             // everybody's part is there; slot (it - 1) % NBUF is free again
        fetch_patch(tile + (TE_NBUF - 1) * (int)gridDim.x, (it + TE_NBUF - 1) % TE_NBUF);
        const uint32_t* pa = patch16[it & 1] + warp * PW + g;        // pixel (row = warp, column = g / g + 8) of the tile
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][2] = bias2[nt][0]; acc[nt][1] = acc[nt][3] = bias2[nt][1]; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
            if (aoff[ks][0] >= 0) { a0 = pa[aoff[ks][0]]; a1 = pa[aoff[ks][0] + 8]; }
            if (aoff[ks][1] >= 0) { a2 = pa[aoff[ks][1]]; a3 = pa[aoff[ks][1] + 8]; }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_16816(acc[nt], a0, a1, a2, a3, bfrag[ks][nt][0], bfrag[ks][nt][1]);
        }
        int n, h0, w0; thin_tile_coords(p, tile, n, h0, w0);
        // accumulator columns (nt, 2t / 2t+1) were given the channels 8t + 2nt / + 1: a lane holds 8 consecutive channels of
        // pixel g and of pixel g + 8 -> two 16-byte stores, a warp instruction covers 8 whole pixels (512 contiguous bytes)
        __nv_bfloat16* o = p.out + ((size_t)(n * p.H + h0 + warp) * p.W + w0 + g) * p.out_ld + p.out_coff + n0 + 8 * t;
        *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(acc[0][0], acc[0][1]), pack_bf16x2(acc[1][0], acc[1][1]),
                                                  pack_bf16x2(acc[2][0], acc[2][1]), pack_bf16x2(acc[3][0], acc[3][1]));
        *reinterpret_cast<uint4*>(o + (size_t)8 * p.out_ld) = make_uint4(pack_bf16x2(acc[0][2], acc[0][3]), pack_bf16x2(acc[1][2], acc[1][3]),
                                                                          pack_bf16x2(acc[2][2], acc[2][3]), pack_bf16x2(acc[3][2], acc[3][3]));
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// thin_wgrad: dW[(tap, ct), n0 .. n0+32) += sum over this CTA's tiles of im2col(thin)^T * wide
// A = the im2col tile read MN-major (M = J padded to 128 over 2 atoms, K = 128 pixels in 8 steps of 16);
// B = the wide bf16 tile {32 ch x 16 x 8 pixels} brought by TMA as 128 rows of 64 B (MN-major, 64B swizzle).
// Two stages of {A, B}; the accumulator stays in TMEM for the CTA's whole pixel range (split-K), then is
// added into dW with global reductions.
// ---------------------------------------------------------------------------------------------
constexpr int TW_A_STAGE = 2 * TH_ATOM, TW_B_STAGE = 128 * 64;
constexpr int TW_SMEM = 2 * TW_A_STAGE + 2 * TW_B_STAGE + TH_PATCH_MAX * 8 + 128 + 1024;

template <int KSZ, bool FLIP>
__global__ void __launch_bounds__(128)
thin_wgrad_kernel(const __grid_constant__ CUtensorMap wide_map, const __grid_constant__ ThinParams p) {
    constexpr uint32_t IDESC = make_idesc_bf16(128, 32, 1, 1);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int atoms = p.nchunks > 8 ? 2 : 1;             // A stage = `atoms` 128 x 128 B tiles (see thin_gemm)
    const int a_stage = atoms * TH_ATOM;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 2 * a_stage;
    float2* patch = reinterpret_cast<float2*>(sB + 2 * TW_B_STAGE);
    uint64_t* bar_b = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(patch) + TH_PATCH_MAX * 8);
    uint64_t* bar_mma = bar_b + 2;
    uint64_t* bar_done = bar_mma + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int n0 = blockIdx.y * 32;

    if (threadIdx.x == 0) {
        mbar_init(bar_b, 1); mbar_init(bar_b + 1, 1); mbar_init(bar_mma, 1); mbar_init(bar_mma + 1, 1); mbar_init(bar_done, 1);
        fence_barrier_init();
        prefetch_tmap(&wide_map);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 32);
    // chunks the per-tile build never writes (im2col columns >= 16*KS) feed accumulator rows that are never
    // read back; zero them once so that no NaN patterns circulate
    for (int s = 0; s < 2; ++s)
        for (int idx = threadIdx.x; idx < 128 * 8 * atoms; idx += 128) {
            const int c = idx >> 7, m = idx & 127;
            if (c >= p.nchunks)
                *reinterpret_cast<uint4*>(sA + s * a_stage + (c >> 3) * TH_ATOM + m * 128 + (((c & 7) ^ (m & 7)) << 4)) = make_uint4(0, 0, 0, 0);
        }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_hi = (1024u >> 4) | (1u << 14) | (SWZ_128B << 29);
    const uint32_t b_hi = (512u >> 4) | (1u << 14) | (SWZ_64B << 29);
    const uint32_t a_lo0 = ((uint32_t)(TH_ATOM >> 4) << 16) | (smem_u32(sA) >> 4);
    const uint32_t b_lo0 = ((uint32_t)(TW_B_STAGE >> 4) << 16) | (smem_u32(sB) >> 4);

    float2 pr[3];
    int pyx[3]; thin_patch_coords(p, pyx);
    int tile = blockIdx.x;
    int n_iters = 0;
    if (tile < p.total_tiles) thin_load_patch(p, tile, pyx, pr);
    for (int it = 0; tile < p.total_tiles; ++it, tile += gridDim.x) {
        const int s = it & 1;
        if (it >= 2) mbar_wait(bar_mma + s, ((it >> 1) - 1) & 1);       // the MMAs that read stage s are done
        if (warp == 0) {
            int n, h0, w0; thin_tile_coords(p, tile, n, h0, w0);
            mbar_expect_tx_elect(bar_b + s, TW_B_STAGE);
            tma_load_4d_elect(&wide_map, bar_b + s, sB + s * TW_B_STAGE, n0, w0, h0, n);
        }
        thin_store_patch(p, patch, pr);
        __syncthreads();
        thin_build_row<KSZ, FLIP>(p, patch, sA + s * a_stage);
        const int next = tile + gridDim.x;
        if (next < p.total_tiles) thin_load_patch(p, next, pyx, pr);
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            mbar_wait(bar_b + s, (it >> 1) & 1);
            fence_after_sync();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo0 + ((s * a_stage + i * 2048) >> 4));
                const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo0 + ((s * TW_B_STAGE + i * 1024) >> 4));
                umma_bf16_elect(tmem_base, ad, bd, IDESC, (it | i) != 0);
            }
            umma_commit_elect(bar_mma + s);
            __syncwarp();
        }
        n_iters = it + 1;
    }
    if (n_iters > 0) {
        if (warp == 0) { umma_commit_elect(bar_done); __syncwarp(); }
        mbar_wait(bar_done, 0);
        fence_after_sync();
    }
    gate_enter(p.gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
    if (n_iters > 0) {
        const int j = threadIdx.x;                       // accumulator row = (tap, thin channel)
        const int tap = j >> 1, ct = j & 1;
        const bool valid = tap < p.ntaps;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(lane_addr + c0, r);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    atomicAdd(p.dw + thin_widx(p, tap, ct, n0 + c0 + q), __uint_as_float(r[q]));
            }
        }
        fence_before_sync();
    }
    gate_leave(p.gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
    __syncthreads();
    if (warp == 0) { fence_after_sync(); tmem_dealloc(tmem_base, 32); }
}

// ---------------------------------------------------------------------------------------------
// thin_wgrad_mma: the same reduction as thin_wgrad on warp-level MMAs, without building im2col rows.
// D[channel, (tap, ct)] += wide^T[channel, pixel] * im2col[pixel, (tap, ct)], GEMM-K = the 16 pixels of one tile row per
// warp. A (wide^T) comes out of the NHWC tile with ldmatrix.trans; a B fragment register is two horizontally adjacent
// pixels of one thin channel at a tap offset, i.e. one 32-bit word of a channel-planar bf16 copy of the halo patch --
// kept in two alignments (pairs starting at even / odd columns) so that every tap offset is an aligned word. Both
// tensors arrive through 4-deep cp.async rings (wide tile 8 KB, patch <= 2.2 KB per slot). 8 warps = the 8 tile rows,
// accumulators (2 x NTL m16n8 tiles) in registers for the CTA's whole tile range, folded across the warps in shared
// memory in a fixed order and added to dW once per CTA (gated in deterministic mode like thin_wgrad).
// ---------------------------------------------------------------------------------------------
constexpr int TWM_NBUF = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_row)) : "memory");
}

template <int KSZ, bool FLIP>
__global__ void __launch_bounds__(256, 2)
thin_wgrad_mma_kernel(const __grid_constant__ ThinParams p) {
    constexpr int NT = KSZ * KSZ, PW = TH_BW + KSZ - 1, PH = TH_BH + KSZ - 1, NP = PH * PW;
    constexpr int J = 2 * NT, NTL = (J + 7) / 8, JP = NTL * 8;       // (tap, thin channel) columns in n-tiles of 8
    constexpr int PITCH = (PW + 2) / 2;                              // 32-bit words (pixel pairs) per plane row
    constexpr int PLANE = PH * PITCH;
    static_assert(NP <= 512, "two patch pixels per thread");
    static_assert(32 * JP * 4 <= TWM_NBUF * 8192, "the reduction buffer aliases the wide ring");
    __shared__ __align__(128) uint8_t wide_s[TWM_NBUF][128 * 64];    // [tile pixel][32 ch], 16-byte chunks XOR-swizzled with (pixel / 2) % 4
    __shared__ float2 patch[TWM_NBUF][NP];
    __shared__ uint32_t planes[2][4 * PLANE];                        // [buffer][(thin channel, alignment)][row][pair]
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int n0 = blockIdx.y * 32;

    // B fragment word offsets: n-tile nt, column g -> (tap, ct) = ((8 nt + g) / 2, g & 1); half h = pixels 2t + 8h, + 1
    int boff[NTL][2];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
        const int j = nt * 8 + g, tap = j >> 1, ct = j & 1;
        const int r = tap / KSZ, sx = tap - r * KSZ;
        const int dr = FLIP ? KSZ - 1 - r : r, ds = FLIP ? KSZ - 1 - sx : sx;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int x = 2 * t + 8 * h + ds, al = x & 1;
            boff[nt][h] = tap < NT ? (ct * 2 + al) * PLANE + (warp + dr) * PITCH + ((x - al) >> 1) : -1;
        }
    }
    // A fragments: ldmatrix.x4.trans row addresses (matrix = lane / 8: pixels 0-7 / 8-15 of the warp's row x channel octet)
    int aoff[2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int mat = lane >> 3, P = warp * 16 + (mat >> 1) * 8 + (lane & 7), chunk = 2 * mt + (mat & 1);
        aoff[mt] = P * 64 + ((chunk ^ ((P >> 1) & 3)) << 4);
    }
    int py[2], px[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int i = threadIdx.x + 256 * q;
        py[q] = i < NP ? i / PW : -1; px[q] = i - (i / PW) * PW;
    }
    auto fetch = [&](int tile, int buf) {
        if (tile < p.total_tiles) {
            int n, h0, w0; thin_tile_coords(p, tile, n, h0, w0);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = threadIdx.x + 256 * q, P = idx >> 2, c = idx & 3;
                const __nv_bfloat16* src = p.wide + ((size_t)(n * p.H + h0 + (P >> 4)) * p.W + w0 + (P & 15)) * p.wide_ld + p.wide_coff + n0 + c * 8;
                cp_async16(&wide_s[buf][P * 64 + ((c ^ ((P >> 1) & 3)) << 4)], src);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (py[q] >= 0) {
                    const int gh = h0 + p.oh0 + py[q], gw = w0 + p.ow0 + px[q];
                    const bool in = gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
                    const float* src = p.thin + ((size_t)(n * p.H + (in ? gh : 0)) * p.W + (in ? gw : 0)) * p.thin_ld + p.thin_coff;
                    cp_async8_zfill(&patch[buf][threadIdx.x + 256 * q], src, in);
                }
        }
        cp_async_commit();
    };

    float acc[2][NTL][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

    int tile = blockIdx.x;
#pragma unroll
    for (int i = 0; i < TWM_NBUF - 1; ++i) fetch(tile + i * (int)gridDim.x, i);
    for (int it = 0; tile < p.total_tiles; tile += gridDim.x, ++it) {
        cp_async_wait<TWM_NBUF - 2>();
        {   // this thread's patch pixels -> the four bf16 planes of buffer it & 1
            unsigned short* pl = reinterpret_cast<unsigned short*>(planes[it & 1]);
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (py[q] >= 0) {
                    const float2 v = patch[it % TWM_NBUF][threadIdx.x + 256 * q];
                    const unsigned short b0 = __bfloat16_as_ushort(f2bf(v.x)), b1 = __bfloat16_as_ushort(f2bf(v.y));
                    const int e = py[q] * 2 * PITCH + px[q];
                    pl[e] = b0; pl[4 * PLANE + e] = b1;                                       // pairs from even columns
                    if (px[q] > 0) { pl[2 * PLANE + e - 1] = b0; pl[6 * PLANE + e - 1] = b1; }   // pairs from odd columns
                }
        }
        __syncthreads();
        fetch(tile + (TWM_NBUF - 1) * (int)gridDim.x, (it + TWM_NBUF - 1) % TWM_NBUF);
        uint32_t a[2][4];
        ldmatrix_x4_trans(a[0], wide_s[it % TWM_NBUF] + aoff[0]);
        ldmatrix_x4_trans(a[1], wide_s[it % TWM_NBUF] + aoff[1]);
        const uint32_t* pw = planes[it & 1];
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) {
            uint32_t b0 = 0u, b1 = 0u;
            if (boff[nt][0] >= 0) { b0 = pw[boff[nt][0]]; b1 = pw[boff[nt][1]]; }
            mma_16816(acc[0][nt], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
            mma_16816(acc[1][nt], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // fold the 8 warps' accumulators in warp order (no floating-point atomics: the sum does not depend on timing)
    float* red = reinterpret_cast<float*>(&wide_s[0][0]);           // [32 channels][JP]
#pragma unroll 1
    for (int w = 0; w < 8; ++w) {
        if (warp == w) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float* d = red + (16 * mt + g + 8 * (e >> 1)) * JP + 8 * nt + 2 * t + (e & 1);
                        *d = w == 0 ? acc[mt][nt][e] : *d + acc[mt][nt][e];
                    }
        }
        __syncthreads();
    }
    gate_enter(p.gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
    for (int idx = threadIdx.x; idx < 32 * J; idx += 256) {
        const int j = idx >> 5, ch = idx & 31;
        atomicAdd(p.dw + thin_widx(p, j >> 1, j & 1, n0 + ch), red[ch * JP + j]);
    }
    gate_leave(p.gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool thin_geometry_ok(const urir_conv_desc* d) {
    return d->stride == 1 && d->P == d->H && d->Q == d->W && d->W % TH_BW == 0 && d->H % TH_BH == 0 &&
           d->R * d->S <= 36 && d->R <= 6 && d->S <= 6 &&
           (long long)d->N * (d->W / TH_BW) * (d->H / TH_BH) < (1 << 20);       // thin_tile_coords' exact-division range
}
// op: 0 fprop (thin = x), 1 dgrad (thin = dy), 2 wgrad (either)
bool thin_supported(const urir_conv_desc* d, int op) {
    if (!thin_geometry_ok(d)) return false;
    const bool thin_x = d->x_dtype == URIR_F32 && d->C == 2 && d->x_ld % 2 == 0 && d->x_coff % 2 == 0 &&
                        d->y_dtype == URIR_BF16 && d->K % 32 == 0 && d->y_ld % 8 == 0 && d->y_coff % 8 == 0;
    const bool thin_y = d->y_dtype == URIR_F32 && d->K == 2 && d->y_ld % 2 == 0 && d->y_coff % 2 == 0 &&
                        d->x_dtype == URIR_BF16 && d->C % 32 == 0 && d->x_ld % 8 == 0 && d->x_coff % 8 == 0;
    if (op == 0) return thin_x && d->act == URIR_ACT_NONE && !d->accumulate;
    if (op == 1) return thin_y && !d->accumulate;
    return thin_x || thin_y;
}

static void thin_fill(ThinParams& p, const urir_conv_desc* d, bool thin_is_x) {
    memset(&p, 0, sizeof(p));
    p.N = d->N; p.H = d->H; p.W = d->W;
    p.ntaps = d->R * d->S;
    p.PH = TH_BH + d->R - 1; p.PW = TH_BW + d->S - 1;
    p.tiles_w = d->W / TH_BW; p.tiles_h = d->H / TH_BH; p.total_tiles = p.tiles_w * p.tiles_h * d->N;
    p.KS = (2 * p.ntaps + 15) / 16; p.nchunks = 2 * p.KS;
    p.mag_w = ((1ull << 40) + p.tiles_w - 1) / p.tiles_w; p.mag_h = ((1ull << 40) + p.tiles_h - 1) / p.tiles_h;
    p.thin_is_x = thin_is_x ? 1 : 0;
    p.CW_total = thin_is_x ? d->K : d->C;
    if (thin_is_x) { p.oh0 = -d->pad_top; p.ow0 = -d->pad_left; }
    else { p.oh0 = d->pad_top - (d->R - 1); p.ow0 = d->pad_left - (d->S - 1); }
    for (int r = 0; r < d->R; ++r)
        for (int s = 0; s < d->S; ++s)
            p.toff[r * d->S + s] = (short)(thin_is_x ? r * p.PW + s : (d->R - 1 - r) * p.PW + (d->S - 1 - s));
}

// stem fprop (thin = x fp32) and head dgrad (thin = dy fp32)
int thin_gemm(const urir_conv_desc* d, const void* thin, const void* w_ck, const float* bias, void* wide, bool thin_is_x,
              cudaStream_t st) {
    URIR_CHECK_ARG(w_ck != nullptr, "thin conv needs w_ck");
    ThinParams p; thin_fill(p, d, thin_is_x);
    p.thin = (const float*)thin; p.w_ck = (const __nv_bfloat16*)w_ck; p.bias = bias; p.out = (__nv_bfloat16*)wide;
    if (thin_is_x) { p.thin_ld = d->x_ld; p.thin_coff = d->x_coff; p.out_ld = d->y_ld; p.out_coff = d->y_coff; }
    else { p.thin_ld = d->y_ld; p.thin_coff = d->y_coff; p.out_ld = d->x_ld; p.out_coff = d->x_coff; }
    // square 3x3 / 6x6 kernels (the stem and the head of the model) get compile-time tap offsets; anything else the generic path
    const int ksz = (d->R == d->S && (d->R == 3 || d->R == 6)) ? d->R : 0;
    { static int umma = -1; if (umma < 0) { const char* e = getenv("URIR_THIN_UMMA"); umma = (e && e[0] == '1') ? 1 : 0; }
      if (ksz && !umma) {                 // register-im2col warp-MMA kernel (see thin_expand_mma_kernel); URIR_THIN_UMMA=1: the tcgen05 one
          void (*k2)(const ThinParams) = ksz == 3 ? (thin_is_x ? thin_expand_mma_kernel<3, false> : thin_expand_mma_kernel<3, true>)
                                                  : (thin_is_x ? thin_expand_mma_kernel<6, false> : thin_expand_mma_kernel<6, true>);
          int per_sm = ksz == 3 ? 3 : 2;      // 78 / 128 registers
          { static int ov = -1; if (ov < 0) { const char* e = getenv("URIR_THIN_MMA_PER_SM"); ov = e ? atoi(e) : 0; } if (ov > 0) per_sm = ov; }
          const int gx = p.total_tiles < sm_count() * per_sm ? p.total_tiles : sm_count() * per_sm;
          URIR_CUDA_OK(launch_pdl(k2, dim3(gx, p.CW_total / 32), dim3(256), 0, st, p));
          URIR_LAUNCH_OK(0);
          return URIR_OK;
      } }
    void (*kern)(const ThinParams) = thin_gemm_kernel<0, false>;
    if (ksz == 3) kern = thin_is_x ? thin_gemm_kernel<3, false> : thin_gemm_kernel<3, true>;
    if (ksz == 6) kern = thin_is_x ? thin_gemm_kernel<6, false> : thin_gemm_kernel<6, true>;
    static std::atomic<bool> attr_set[3][2];
    if (!attr_set[ksz / 3][thin_is_x]) {
        URIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
        attr_set[ksz / 3][thin_is_x] = true;
    }
    // the per-tile phases of a CTA (patch load, im2col build, MMA, store) run back to back, so throughput comes
    // from co-resident CTAs: as many as the shared memory of the layer's atom count allows
    const int atoms = p.nchunks > 8 ? 2 : 1;
    const int smem = atoms * (TH_ATOM + TG_B_ATOM) + TH_PATCH_MAX * 8 + 64 + 1024;
    // = the CTAs that are actually resident (93 registers, <= 43 KB): a grid of 8 per SM ran as 1.6 waves of persistent
    // CTAs. Measured per call (B = 64): stem fprop 49 -> 41 us, head dgrad 75 -> 66 us against 8 / 4 per SM.
    int per_sm = 5;
    { static int ov = -1; if (ov < 0) { const char* e = getenv("URIR_THIN_GEMM_PER_SM"); ov = e ? atoi(e) : 0; } if (ov > 0) per_sm = ov; }
    int gx = p.total_tiles < sm_count() * per_sm ? p.total_tiles : sm_count() * per_sm;
    dim3 grid(gx, p.CW_total / 32);
    kern<<<grid, 128, smem, st>>>(p);
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

// stem wgrad (thin = x fp32, wide = dy) and head wgrad (thin = dy fp32, wide = x); dw fp32 [tap][C][K], overwritten
int thin_wgrad(const urir_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    const bool thin_is_x = d->x_dtype == URIR_F32;
    ThinParams p; thin_fill(p, d, thin_is_x);
    const void* wide; int wide_ld, wide_coff, wide_c;
    if (thin_is_x) { p.thin = (const float*)x; p.thin_ld = d->x_ld; p.thin_coff = d->x_coff; wide = dy; wide_ld = d->y_ld; wide_coff = d->y_coff; wide_c = d->K; }
    else { p.thin = (const float*)dy; p.thin_ld = d->y_ld; p.thin_coff = d->y_coff; wide = x; wide_ld = d->x_ld; wide_coff = d->x_coff; wide_c = d->C; }
    p.dw = dw; p.gate = next_gate();
    p.wide = (const __nv_bfloat16*)wide; p.wide_ld = wide_ld; p.wide_coff = wide_coff;
    const int ksz_mma = (d->R == d->S && (d->R == 3 || d->R == 6)) ? d->R : 0;
    { static int umma = -1; if (umma < 0) { const char* e = getenv("URIR_THIN_UMMA"); umma = (e && e[0] == '1') ? 1 : 0; }
      if (ksz_mma && !umma) {               // warp-MMA kernel (thin_wgrad_mma_kernel); URIR_THIN_UMMA=1: the tcgen05 one below
          void (*k2)(const ThinParams) = ksz_mma == 3 ? (thin_is_x ? thin_wgrad_mma_kernel<3, false> : thin_wgrad_mma_kernel<3, true>)
                                                      : (thin_is_x ? thin_wgrad_mma_kernel<6, false> : thin_wgrad_mma_kernel<6, true>);
          if (!d->accumulate) URIR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p.ntaps * d->C * d->K, st));
          int per_sm = 2;
          { static int ov = -1; if (ov < 0) { const char* e = getenv("URIR_THIN_WGRAD_PER_SM"); ov = e ? atoi(e) : 0; } if (ov > 0) per_sm = ov; }
          const int gx = p.total_tiles < sm_count() * per_sm ? p.total_tiles : sm_count() * per_sm;
          URIR_CUDA_OK(launch_pdl(k2, dim3(gx, p.CW_total / 32), dim3(256), 0, st, p));
          URIR_LAUNCH_OK(0);
          return URIR_OK;
      } }
    CUtensorMap map;
    {
        const uint64_t dims[4] = {(uint64_t)wide_c, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)wide_ld * 2, (uint64_t)d->W * wide_ld * 2, (uint64_t)d->H * d->W * wide_ld * 2};
        const uint32_t box[4] = {32, TH_BW, TH_BH, 1};
        int rc = encode_map(&map, (const char*)wide + (size_t)wide_coff * 2, 4, dims, strides, box, 64);
        if (rc) return rc;
    }
    const int ksz = (d->R == d->S && (d->R == 3 || d->R == 6)) ? d->R : 0;
    void (*kern)(const CUtensorMap, const ThinParams) = thin_wgrad_kernel<0, false>;
    if (ksz == 3) kern = thin_is_x ? thin_wgrad_kernel<3, false> : thin_wgrad_kernel<3, true>;
    if (ksz == 6) kern = thin_is_x ? thin_wgrad_kernel<6, false> : thin_wgrad_kernel<6, true>;
    static std::atomic<bool> attr_set[3][2];
    if (!attr_set[ksz / 3][thin_is_x]) {
        URIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
        attr_set[ksz / 3][thin_is_x] = true;
    }
    if (!d->accumulate) URIR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p.ntaps * d->C * d->K, st));
    const int atoms = p.nchunks > 8 ? 2 : 1;
    // with one atom the M = 128 instruction also reads the 16 KB after its A stage (the other stage / the B
    // stages: finite bf16 data inside the allocation); those accumulator rows 64..127 are never read back
    const int smem = 2 * atoms * TH_ATOM + 2 * TW_B_STAGE + TH_PATCH_MAX * 8 + 128 + 1024;
    int per_sm = 2;                                    // more CTAs only add dw reductions (measured slower)
    { static int ov = -1; if (ov < 0) { const char* e = getenv("URIR_THIN_WGRAD_PER_SM"); ov = e ? atoi(e) : 0; } if (ov > 0 && atoms == 1) per_sm = ov; }
    int gx = p.total_tiles < sm_count() * per_sm ? p.total_tiles : sm_count() * per_sm;
    dim3 grid(gx, p.CW_total / 32);
    kern<<<grid, 128, smem, st>>>(map, p);
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

}  // namespace urir
