// tcgen05 weight-gradient kernel for sm_100a:
//   dw[tap][c][k] (fp32, HWIO) = sum over output pixels  x[pixel + tap][c] * dy[pixel][k]
// (Conv2D / Conv2DTranspose kernels' gradients under tape.gradient, amp_phase_trainer.py:138).
//
// GEMM view per tap: D[c (M = 128), k (N)] += A[c, pixels] * B[pixels, k], the reduction (GEMM-K)
// runs over pixels. Both operands are the NHWC tensors exactly as they sit in HBM: a TMA box
// {64 channels, bw x bh x bn pixels} lands as `KP` rows of 128 B, which is the canonical
// MN-major (channel-contiguous) swizzled UMMA operand, so no transpose is ever materialised.
// The same tap table / parity-view trick as conv_igemm.cu provides the shifted (and, for stride 2,
// parity-strided) x tiles; TMA zero-fill supplies the SAME padding.
// A CTA owns (tap group of T taps, 128-channel c tile, BLOCK_N k tile) and a contiguous range of
// pixel tiles (split-K); T accumulators of BLOCK_N columns live in TMEM; the epilogue adds them
// into dw with vectorised global reductions.
#include <stdlib.h>
#include <atomic>
#include "urir_common.cuh"
#include "urir_tc.cuh"

namespace urir {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes);

struct WgradTap { short dh, dw; short map; short wtap; };

struct WgradParams {
    int bw, bh, bn, kp;             // pixel box, kp = bw*bh*bn (multiple of 16, <= 64)
    int tiles_w, tiles_h, tiles_n;  // pixel tiles over (Q, P, N)
    int tiles_per_cta, total_tiles;
    int C, K;                       // channels of x and dy
    int a_atom, b_atom;             // channels per TMA box / swizzle atom: 64 (128B rows) or 32 (64B rows)
    int a_atoms, b_atoms;           // atoms per tap tile of A (<= 2), atoms of the B tile (BLOCK_N / b_atom)
    int T;                          // taps per CTA (1 or 3)
    int apm, n_mma;                 // A atoms consumed per MMA (128 / a_atom); accumulators per CTA
    int stages, stage_bytes, a_region_bytes;
    int tmem_cols;
    int ntaps, n_mtiles;
    float* dw;
    int direct;                     // 1: a single CTA owns each dw element (no pixel split) -> plain stores, no memset
    long long* trace;               // debug: per-iteration clock64 stamps of CTA (0,0,0) when non-null
    unsigned int* gate;             // deterministic mode: the pixel splits add into dw in blockIdx order (urir_common.cuh)
    WgradTap taps[36];
};

struct WgradMaps { CUtensorMap a[4]; CUtensorMap b; };

constexpr int WG_KP = 64;                    // max pixels per stage
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_SMEM_BUDGET = 200 * 1024;

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Shared memory of one stage: [A atoms: T taps x a_atoms, each kp rows x (a_atom*2) bytes, contiguous]
// [padding up to n_mma*apm atoms][B atoms]. One tcgen05.mma consumes `apm` consecutive A atoms as its
// M = 128 rows (leading-byte-offset = atom size), so for C = 32 three taps share one MMA and for C = 64
// two do: the tensor pipe reads the same 4 KB of A per instruction whatever N is, so stacking taps
// along M is what keeps it busy on the thin layers.
template <int BLOCK_N>
__global__ void __launch_bounds__(192)
conv_wgrad_tc_kernel(const __grid_constant__ WgradMaps maps, const __grid_constant__ WgradParams p) {
    constexpr uint32_t IDESC = make_idesc_bf16(128, BLOCK_N, 1, 1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* empty_bar = full_bar + WG_MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + WG_MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction
    const int m_tile = blockIdx.y % p.n_mtiles, n_tile = blockIdx.y / p.n_mtiles;
    const int tap0 = blockIdx.z * p.T;
    const int tile_begin = blockIdx.x * p.tiles_per_cta;
    int tile_end = tile_begin + p.tiles_per_cta; if (tile_end > p.total_tiles) tile_end = p.total_tiles;
    const int n_iters = tile_end > tile_begin ? tile_end - tile_begin : 0;
    const int STAGES = p.stages;

    const uint32_t a_row = p.a_atom * 2, b_row = p.b_atom * 2;         // bytes per smem row
    const uint32_t a_swz = p.a_atom == 64 ? SWZ_128B : SWZ_64B, b_swz = p.b_atom == 64 ? SWZ_128B : SWZ_64B;
    const uint32_t a_atom_bytes = p.kp * a_row, b_atom_bytes = p.kp * b_row;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    if (warp == 4 && lane == 0) { prefetch_tmap(&maps.b); prefetch_tmap(&maps.a[0]); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 4) {
        // ===================== TMA producer =====================
        // Everything per-iteration is incremental (no divisions, no parameter-table reads in the loop):
        // the single issuing thread must stay well under the ~45 cycles one MMA takes.
        {
            const uint32_t bytes = (uint32_t)p.T * p.a_atoms * a_atom_bytes + (uint32_t)p.b_atoms * b_atom_bytes;
            const CUtensorMap* tmap[3]; int tdw[3], tdh[3];
#pragma unroll
            for (int ti = 0; ti < 3; ++ti) {
                const WgradTap tp = p.taps[tap0 + (ti < p.T ? ti : 0)];
                // shuffles tell the compiler these are warp-uniform, so the TMA operands stay in uniform registers
                tmap[ti] = &maps.a[__shfl_sync(0xffffffffu, (int)tp.map, 0)];
                tdw[ti] = __shfl_sync(0xffffffffu, (int)tp.dw, 0); tdh[ti] = __shfl_sync(0xffffffffu, (int)tp.dh, 0);
            }
            int t = tile_begin;
            int tw = t % p.tiles_w; t /= p.tiles_w;
            int th = t % p.tiles_h; int tn = t / p.tiles_h;
            const int a_c0 = m_tile * 128, b_c0 = n_tile * BLOCK_N;
            int stage = 0; uint32_t phase = 0;
            uint8_t* st_base = smem;
            for (int it = 0; it < n_iters; ++it) {
                const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
                const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 256 && lane == 0;
                if (tr) p.trace[it * 4 + 0] = clock64();
                mbar_wait(empty_bar + stage, phase ^ 1);
                if (tr) p.trace[it * 4 + 1] = clock64();
                mbar_expect_tx_elect(full_bar + stage, bytes);
                uint8_t* dst = st_base;
#pragma unroll
                for (int ti = 0; ti < 3; ++ti) {
                    if (ti < p.T) {
                        tma_load_4d_elect(tmap[ti], full_bar + stage, dst, a_c0, w0 + tdw[ti], h0 + tdh[ti], n0);
                        dst += a_atom_bytes;
                        if (p.a_atoms > 1) {
                            tma_load_4d_elect(tmap[ti], full_bar + stage, dst, a_c0 + p.a_atom, w0 + tdw[ti], h0 + tdh[ti], n0);
                            dst += a_atom_bytes;
                        }
                    }
                }
                dst = st_base + p.a_region_bytes;
                tma_load_4d_elect(&maps.b, full_bar + stage, dst, b_c0, w0, h0, n0);
                if (p.b_atoms > 1) tma_load_4d_elect(&maps.b, full_bar + stage, dst + b_atom_bytes, b_c0 + p.b_atom, w0, h0, n0);
                if (p.b_atoms > 2) {
                    tma_load_4d_elect(&maps.b, full_bar + stage, dst + 2 * b_atom_bytes, b_c0 + 2 * p.b_atom, w0, h0, n0);
                    tma_load_4d_elect(&maps.b, full_bar + stage, dst + 3 * b_atom_bytes, b_c0 + 3 * p.b_atom, w0, h0, n0);
                }
                if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
                st_base += p.stage_bytes;
                if (++stage == STAGES) { stage = 0; phase ^= 1; st_base = smem; }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        // Descriptors: high word (SBO | version | swizzle) and the LBO half of the low word are loop
        // invariants; per MMA only the 14-bit start-address field changes (base + precomputed offset).
        const int ksteps = p.kp / 16, n_mma = p.n_mma;
        const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t a_hi = ((8 * a_row) >> 4) | (1u << 14) | (a_swz << 29);
        const uint32_t b_hi = ((8 * b_row) >> 4) | (1u << 14) | (b_swz << 29);
        const uint32_t a_lo0 = ((a_atom_bytes >> 4) << 16) | (smem_u32(smem) >> 4);
        const uint32_t b_lo0 = ((b_atom_bytes >> 4) << 16) | ((smem_u32(smem) + p.a_region_bytes) >> 4);
        const uint32_t stage16 = p.stage_bytes >> 4;
        const uint32_t a_k16 = (16 * a_row) >> 4, b_k16 = (16 * b_row) >> 4, a_j16 = (p.apm * a_atom_bytes) >> 4;
        int stage = 0; uint32_t phase = 0;
        uint32_t soff = 0;
        for (int it = 0; it < n_iters; ++it) {
            const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && it < 256 && lane == 0;
            if (tr) p.trace[it * 4 + 2] = clock64();
            mbar_wait(full_bar + stage, phase);
            if (tr) p.trace[it * 4 + 3] = clock64();
            fence_after_sync();
            {
                const uint32_t acc = it != 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < ksteps) {
                        const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo0 + soff + k * b_k16);
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            if (j < n_mma) {
                                const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo0 + soff + k * a_k16 + j * a_j16);
                                umma_bf16_elect(tm0 + j * BLOCK_N, ad, bd, IDESC, acc | (k != 0));
                            }
                        }
                    }
                }
                umma_commit_elect(empty_bar + stage);
                if (it == n_iters - 1) umma_commit_elect(tmem_full_bar);
            }
            __syncwarp();
            soff += stage16;
            if (++stage == STAGES) { stage = 0; phase ^= 1; soff = 0; }
        }
    } else if (warp < 4) {
        // ===================== epilogue: TMEM -> red.global.add into dw =====================
        if (n_iters > 0) { mbar_wait(tmem_full_bar, 0); fence_after_sync(); }
        gate_enter(p.gate, cta_linear(), threadIdx.x == 0, 1, 128);
        if (n_iters > 0) {
            const int row = warp * 32 + lane;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int j = 0; j < p.n_mma; ++j) {
                const int g = j * p.apm + row / p.a_atom;                  // A atom this accumulator row came from
                const int ti = g / p.a_atoms;
                const int c = m_tile * 128 + (g % p.a_atoms) * p.a_atom + row % p.a_atom;
                const bool valid = ti < p.T && c < p.C;
                const int wtap = p.taps[tap0 + (valid ? ti : 0)].wtap;
                float* drow = p.dw + ((size_t)wtap * p.C + (valid ? c : 0)) * p.K + n_tile * BLOCK_N;
#pragma unroll 1
                for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(lane_addr + j * BLOCK_N + c0, r);
                    tmem_ld_wait();
                    if (valid) {
                        const int n_left = p.K - n_tile * BLOCK_N - c0;      // valid columns from c0 on
                        if (p.direct && n_left >= 16 && (p.K & 3) == 0) {
#pragma unroll
                            for (int q = 0; q < 16; q += 4)
                                *reinterpret_cast<float4*>(drow + c0 + q) = make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]),
                                                                                      __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
                        } else if (n_left >= 16 && (p.K & 3) == 0) {
#pragma unroll
                            for (int q = 0; q < 16; q += 4)
                                red_add_v4(drow + c0 + q, __uint_as_float(r[q]), __uint_as_float(r[q + 1]),
                                           __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
                        } else {
#pragma unroll
                            for (int q = 0; q < 16; ++q)
                                if (q < n_left) atomicAdd(drow + c0 + q, __uint_as_float(r[q]));
                        }
                    }
                }
            }
            fence_before_sync();
        }
        gate_leave(p.gate, cta_linear(), cta_count(), threadIdx.x == 0, 1, 128);
    }
    __syncthreads();
    if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------
static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

// pixel box with product in {16,32,48,64}: every staged row must be a real (or zero-filled) pixel
static void choose_kbox(int OW, int OH, int NB, int* bw, int* bh, int* bn) {
    double best = -1; int bbw = 1, bbh = 1, bbn = 16;
    for (int w = 1; w <= OW && w <= WG_KP; ++w)
        for (int h = 1; h <= OH && w * h <= WG_KP; ++h)
            for (int n = 1; w * h * n <= WG_KP; ++n) {
                const int kp = w * h * n;
                if (kp % 16) continue;
                const double tiles = (double)((OW + w - 1) / w) * ((OH + h - 1) / h) * ((NB + n - 1) / n);
                const double eff = ((double)OW * OH * NB) / (tiles * kp);
                const double score = eff + 1e-3 * kp / WG_KP + 1e-6 * w;   // prefer full-depth stages, wide boxes
                if (score > best) { best = score; bbw = w; bbh = h; bbn = n; }
            }
    *bw = bbw; *bh = bbh; *bn = bbn;
}

template <int BLOCK_N>
static int launch_wg(const WgradMaps& maps, const WgradParams& p, dim3 grid, int smem_bytes, cudaStream_t st) {
    static std::atomic<bool> attr_set{false};    // benign if two threads both set the attribute
    auto kern = conv_wgrad_tc_kernel<BLOCK_N>;
    if (!attr_set) {
        URIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BUDGET + 4096));
        attr_set = true;
    }
    URIR_CUDA_OK(launch_pdl(kern, grid, dim3(192), smem_bytes, st, maps, p));
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

// channel counts that are not multiples of the tile run masked tiles; TMA zero-fills the missing channels
static int pick_n(int k) { return (k % 128 == 0) ? 128 : (k % 64 == 0) ? 64 : 32; }

bool wgrad_tc_supported(const urir_conv_desc* d) {
    return d->x_dtype == URIR_BF16 && d->y_dtype == URIR_BF16 && (d->C <= 64 || d->C % 128 == 0) &&
           d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && d->y_ld % 8 == 0 && d->y_coff % 8 == 0 &&
           (d->stride == 1 || d->stride == 2) && d->R * d->S <= 36;
}

int conv_wgrad_tc(const urir_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    WgradMaps maps; WgradParams p;
    memset(&p, 0, sizeof(p));
    const int BN = pick_n(d->K);
    const int ntaps = d->R * d->S;
    const int T = (ntaps % 3 == 0) ? 3 : 1;
    choose_kbox(d->Q, d->P, d->N, &p.bw, &p.bh, &p.bn);
    p.kp = p.bw * p.bh * p.bn;
    p.tiles_w = cdiv(d->Q, p.bw); p.tiles_h = cdiv(d->P, p.bh); p.tiles_n = cdiv(d->N, p.bn);
    p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.C = d->C; p.K = d->K; p.ntaps = ntaps; p.dw = dw; p.T = T;
    {   // URIR_WGRAD_TRACE=<device pointer, hex> : debug timeline of CTA (0,0,0)
        const char* e = getenv("URIR_WGRAD_TRACE");
        p.trace = e ? (long long*)strtoull(e, nullptr, 16) : nullptr;
    }
    p.a_atom = (d->C % 64 == 0) ? 64 : 32; p.b_atom = (d->K % 64 == 0) ? 64 : 32;
    const int m_valid = d->C < 128 ? d->C : 128;
    p.a_atoms = (m_valid + p.a_atom - 1) / p.a_atom; p.b_atoms = BN / p.b_atom;
    p.n_mtiles = cdiv(d->C, 128);
    p.apm = 128 / p.a_atom;
    p.n_mma = cdiv(T * p.a_atoms, p.apm);
    const int a_atom_bytes = p.kp * p.a_atom * 2, b_atom_bytes = p.kp * p.b_atom * 2;
    p.a_region_bytes = p.n_mma * p.apm * a_atom_bytes;            // includes the unused tail atoms of the last MMA
    p.stage_bytes = p.a_region_bytes + p.b_atoms * b_atom_bytes;
    p.stage_bytes = (p.stage_bytes + 1023) / 1024 * 1024;
    p.stages = WG_SMEM_BUDGET / p.stage_bytes;
    if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
    if (p.stages < 2) return fail(URIR_ERR_UNSUP, "wgrad(tcgen05): stage of %d bytes does not fit twice", p.stage_bytes);
    { int cols = p.n_mma * BN; p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }
    const int smem_bytes = p.stages * p.stage_bytes + (2 * WG_MAX_STAGES + 2) * 8 + 16 + 1024;
    const int n_ntiles = (d->K + BN - 1) / BN, tap_groups = ntaps / T;
    const int units = p.n_mtiles * n_ntiles * tap_groups;
    // One wave: as many pixel splits as fit on the machine at once (CTAs per SM follow from the shared-memory
    // footprint). A grid slightly above a whole number of waves (300 CTAs on 296 slots) costs a full extra wave.
    const int ctas_per_sm = smem_bytes <= 113 * 1024 ? 2 : 1;
    int splits = (sm_count() * ctas_per_sm) / units;
    if (splits > p.total_tiles) splits = p.total_tiles;
    if (splits < 1) splits = 1;
    p.tiles_per_cta = cdiv(p.total_tiles, splits);
    splits = cdiv(p.total_tiles, p.tiles_per_cta);

    const int s = d->stride;
    int nt = 0;
    bool used[4] = {false, false, false, false};
    for (int r = 0; r < d->R; ++r)
        for (int q = 0; q < d->S; ++q) {
            const int th = r - d->pad_top, tw = q - d->pad_left;
            const int ph = posmod(th, s), pw = posmod(tw, s);
            WgradTap& tp = p.taps[nt++];
            tp.dh = (short)floordiv(th, s); tp.dw = (short)floordiv(tw, s);
            tp.map = (short)(ph * s + pw); tp.wtap = (short)(r * d->S + q);
            used[tp.map] = true;
        }
    const uint32_t abox[4] = {(uint32_t)p.a_atom, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    const char* xb = (const char*)x + (size_t)d->x_coff * 2;
    int first_used = -1;
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw) {
            const int mi = ph * s + pw;
            if (!used[mi]) continue;
            if (first_used < 0) first_used = mi;
            const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)((d->W - pw + s - 1) / s), (uint64_t)((d->H - ph + s - 1) / s), (uint64_t)d->N};
            const uint64_t strides[3] = {(uint64_t)s * d->x_ld * 2, (uint64_t)s * d->W * d->x_ld * 2, (uint64_t)d->H * d->W * d->x_ld * 2};
            int rc = encode_map(&maps.a[mi], xb + ((size_t)ph * d->W + pw) * d->x_ld * 2, 4, dims, strides, abox, p.a_atom * 2);
            if (rc) return rc;
        }
    for (int mi = 0; mi < 4; ++mi) if (mi >= s * s || !used[mi]) maps.a[mi] = maps.a[first_used];
    {
        const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->Q, (uint64_t)d->P, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->y_ld * 2, (uint64_t)d->Q * d->y_ld * 2, (uint64_t)d->P * d->Q * d->y_ld * 2};
        const uint32_t bbox[4] = {(uint32_t)p.b_atom, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
        int rc = encode_map(&maps.b, (const char*)dy + (size_t)d->y_coff * 2, 4, dims, strides, bbox, p.b_atom * 2);
        if (rc) return rc;
    }
    // without a pixel split every dw element has exactly one producer (all channel tiles exact): store, do not add
    p.direct = !d->accumulate && splits == 1 && d->K % BN == 0 && (d->K & 3) == 0 && (d->C % 128 == 0 || d->C == m_valid) ? 1 : 0;
    if (!p.direct && !d->accumulate) URIR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)ntaps * d->C * d->K, st));
    p.gate = p.direct ? nullptr : next_gate();
    dim3 grid(splits, p.n_mtiles * n_ntiles, tap_groups);
    if (BN == 128) return launch_wg<128>(maps, p, grid, smem_bytes, st);
    if (BN == 64) return launch_wg<64>(maps, p, grid, smem_bytes, st);
    if (BN == 32) return launch_wg<32>(maps, p, grid, smem_bytes, st);
    return fail(URIR_ERR_UNSUP, "wgrad(tcgen05): no kernel for BLOCK_N=%d T=%d", BN, T);
}

}  // namespace urir
