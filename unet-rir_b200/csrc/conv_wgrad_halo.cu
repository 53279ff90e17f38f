// Persistent halo-tile weight-gradient kernel (tcgen05, sm_100a) for the 3x3 stride-1 convolutions at the wide
// resolutions (convolutional_block_1, dl_models/u_net.py:363-371, under tape.gradient, amp_phase_trainer.py:138):
//   dw[dh][dw][c][k] = sum over pixels (h, w) of  x[h + dh - 1, w + dw - 1][c] * dy[h, w][k]
//
// conv_wgrad_tc.cu stages one shifted x tile per tap (9 x the activation bytes through L2 -> shared memory) and
// issues one N = 32/64 MMA per 3 taps, which costs 44-48 cycles whatever N is. Here ALL NINE taps come out of two
// boxes per pixel tile and (for 32-channel layers) ONE instruction per 16 pixels:
//   * substitute w' = w + dw - 1:  dw[dh][dw] = sum over (h, w') of x[h + dh - 1, w'] * dy[h, w' - dw + 1];
//     out-of-range x or dy rows are TMA zero fill, which is exactly the SAME padding.
//   * the pixel tile is 8 (h) x 16 (w'), GEMM-K = pixels, both operands MN-major straight from NHWC, H fastest in
//     shared memory: x box {Cc ch, 10 rows, 16 cols} (one halo row above/below), dy box {KN ch, 8 rows, 18 cols}
//     (one halo column left/right). A K-group of 8 pixels = 8 vertically adjacent pixels of one column.
//   * the vertical taps are stacked along GEMM-M: "atom" j of the A descriptor starts j rows lower (LBO = one row,
//     SBO = 10 rows = next column), so M = 128 = 4 (or 2) overlapping views of the same x box;
//   * the horizontal taps are stacked along GEMM-N: atom i of the B descriptor starts i columns further right
//     (LBO = SBO = 8 rows = one column), so N = 3 * KN overlapping views of the same dy box.
//   The overlap is legal because the UMMA swizzle is a function of the shared-memory address
//   (profiles/r01_umma_halo_probe.txt, tools/umma_halo_test2.cu mode 2).
// Every activation byte is staged once (152 B per pixel instead of 768) and the MMA count drops 3-6 x; the
// layers become HBM-bound. CTAs are persistent (one per SM, grid.y = blocks of Cc x-channels), accumulate their
// whole pixel range in TMEM and add it into dw with red.global.add.v4.f32 once at the end.
// Warp roles (192 threads): warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#include <stdlib.h>
#include <atomic>
#include "urir_common.cuh"
#include "urir_tc.cuh"

namespace urir {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes);

constexpr int WH_TH = 8, WH_TW = 16;         // pixel tile (h, w')
constexpr int WH_PH = WH_TH + 2;             // x box rows (vertical halo)
constexpr int WH_PW = WH_TW + 2;             // dy box columns (horizontal halo)
constexpr int WH_MAX_STAGES = 8;
constexpr int WH_SMEM_BUDGET = 200 * 1024;

struct WhaloParams {
    int tiles_w, tiles_h, total_tiles;
    int C, Cc;                  // x channels in total / per CTA (32 or 64)
    int K;                      // dy channels in total (a CTA handles KN of them: blockIdx.z)
    int n_mma;                  // MMAs per k-step: 1 (Cc = 32: 4 vertical slots) or 2 (Cc = 64: 2 slots each)
    int stages, stage_bytes, a_bytes, tx_bytes;
    int tmem_cols;
    int debug;                  // URIR_WH_DEBUG: 1 = skip the epilogue adds, 2 = do not rotate the epilogue order
    float* dw;
    unsigned int* gate;         // deterministic mode: CTAs add into dw in blockIdx order (urir_common.cuh)
};

struct WhaloMaps { CUtensorMap a; CUtensorMap b; };

__device__ __forceinline__ void wh_red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int KN>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_halo_kernel(const __grid_constant__ WhaloMaps maps, const __grid_constant__ WhaloParams p) {
    constexpr int N = 3 * KN;
    constexpr uint32_t IDESC = make_idesc_bf16(128, N, 1, 1);
    constexpr uint32_t B_ROW = KN * 2;
    constexpr uint32_t B_SWZ = KN == 64 ? SWZ_128B : SWZ_64B;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // one spare KB after the ring: the unused 4th vertical slot of the last column reads one row past its box
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes + 1024);
    uint64_t* empty_bar = full_bar + WH_MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + WH_MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int cblk = blockIdx.y, kblk = blockIdx.z;
    const int STAGES = p.stages;
    const int n_iters = blockIdx.x < p.total_tiles ? (p.total_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    if (warp == 4 && lane == 0) { prefetch_tmap(&maps.a); prefetch_tmap(&maps.b); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 4) {
        // ===================== TMA producer: two boxes per pixel tile =====================
        int stage = 0; uint32_t phase = 0;
        uint8_t* dst = smem;
        const int c0 = cblk * p.Cc;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            int t = tile;
            const int tw = t % p.tiles_w; t /= p.tiles_w;
            const int th = t % p.tiles_h; const int n = t / p.tiles_h;
            const int h0 = th * WH_TH, w0 = tw * WH_TW;
            mbar_wait(empty_bar + stage, phase ^ 1);
            mbar_expect_tx_elect(full_bar + stage, (uint32_t)p.tx_bytes);
            tma_load_4d_elect(&maps.a, full_bar + stage, dst, c0, h0 - 1, w0, n);
            tma_load_4d_elect(&maps.b, full_bar + stage, dst + p.a_bytes, kblk * KN, h0, w0 - 1, n);
            dst += p.stage_bytes;
            if (++stage == STAGES) { stage = 0; phase ^= 1; dst = smem; }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t a_row = p.Cc * 2;
        const uint32_t a_swz = p.Cc == 64 ? SWZ_128B : SWZ_64B;
        const uint32_t a_col = WH_PH * a_row, b_col = WH_TH * B_ROW;          // bytes per pixel column of the boxes
        const uint32_t a_hi = (a_col >> 4) | (1u << 14) | (a_swz << 29);       // SBO = next column of the x box
        const uint32_t b_hi = (b_col >> 4) | (1u << 14) | (B_SWZ << 29);       // SBO = next column of the dy box
        const uint32_t a_lo0 = ((a_row >> 4) << 16) | (smem_u32(smem) >> 4);   // LBO = one row down  (vertical tap)
        const uint32_t b_lo0 = ((b_col >> 4) << 16) | ((smem_u32(smem) + p.a_bytes) >> 4);   // LBO = one column right
        const uint32_t a_k16 = (2 * a_col) >> 4, b_k16 = (2 * b_col) >> 4;     // one k-step = 2 pixel columns
        const uint32_t a_j16 = (2 * a_row) >> 4;                               // second MMA of a 64-channel block: taps 2,(3)
        const uint32_t stage16 = p.stage_bytes >> 4;
        const int n_mma = p.n_mma;
        int stage = 0; uint32_t phase = 0, soff = 0;
        for (int it = 0; it < n_iters; ++it) {
            mbar_wait(full_bar + stage, phase);
            fence_after_sync();
#pragma unroll
            for (int k = 0; k < WH_TW / 2; ++k) {
                const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo0 + soff + k * b_k16);
                const uint32_t acc = (it != 0) | (k != 0);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j < n_mma) {
                        const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo0 + soff + k * a_k16 + j * a_j16);
                        umma_bf16_elect(tm0 + j * N, ad, bd, IDESC, acc);
                    }
                }
            }
            umma_commit_elect(empty_bar + stage);
            if (it == n_iters - 1) umma_commit_elect(tmem_full_bar);
            __syncwarp();
            soff += stage16;
            if (++stage == STAGES) { stage = 0; phase ^= 1; soff = 0; }
        }
    } else if (warp < 4) {
        // ===================== epilogue: TMEM -> red.global.add into dw[tap][c][k] =====================
        if (n_iters > 0) { mbar_wait(tmem_full_bar, 0); fence_after_sync(); }
        gate_enter(p.gate, cta_linear(), threadIdx.x == 0, 1, 128);
        if (n_iters > 0) {
            const int row = warp * 32 + lane;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
            const int slot = row / p.Cc, c = row % p.Cc;
            // all CTAs add into the same 9*C*K floats: start each CTA at a different (accumulator, horizontal tap)
            // so that concurrent reductions hit different addresses
            const int n_units = p.n_mma * 3;
            const int u0 = p.debug == 2 ? 0 : blockIdx.x % n_units;
#pragma unroll 1
            for (int uu = 0; uu < n_units; ++uu) {
                int u = uu + u0; if (u >= n_units) u -= n_units;
                const int j = u / 3, i = u - 3 * j;
                const int dh = p.Cc == 32 ? slot : 2 * j + slot;
                const bool valid = dh < 3 && p.debug != 1;
                {
                    const int tap = (valid ? dh : 0) * 3 + (2 - i);              // B atom i <-> horizontal tap 2 - i
                    float* drow = p.dw + ((size_t)tap * p.C + cblk * p.Cc + c) * p.K + kblk * KN;
#pragma unroll
                    for (int c0 = 0; c0 < KN; c0 += 16) {
                        uint32_t r[16];
                        tmem_ld16(lane_addr + j * N + i * KN + c0, r);
                        tmem_ld_wait();
                        if (valid) {
#pragma unroll
                            for (int q = 0; q < 16; q += 4)
                                wh_red_add_v4(drow + c0 + q, __uint_as_float(r[q]), __uint_as_float(r[q + 1]),
                                              __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
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
bool wgrad_halo_supported(const urir_conv_desc* d, bool forced) {
    if (d->stride != 1 || d->R != 3 || d->S != 3 || d->pad_top != 1 || d->pad_left != 1) return false;
    if (d->P != d->H || d->Q != d->W) return false;      // ragged tiles: TMA zero fill contributes nothing
    if (d->x_dtype != URIR_BF16 || d->y_dtype != URIR_BF16) return false;
    if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return false;
    if (!(d->C == 32 || d->C % 64 == 0) || !(d->K == 32 || d->K % 64 == 0)) return false;
    const long long tiles = (long long)d->N * cdiv(d->H, WH_TH) * cdiv(d->W, WH_TW);
    // measured at 36x40 (960 ragged tiles, K = 128 in two blocks): 41 / 68 us vs 43 / 60 us for conv_wgrad_tc -- no gain,
    // so AUTO keeps this kernel for the two wide levels only
    return forced || tiles >= sm_count() * 8;
}

template <int KN>
static int launch_wh(const WhaloMaps& maps, const WhaloParams& p, dim3 grid, int smem, cudaStream_t st) {
    static std::atomic<bool> attr_set{false};    // benign if two threads both set the attribute
    auto kern = conv_wgrad_halo_kernel<KN>;
    if (!attr_set) { URIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_SMEM_BUDGET + 8192)); attr_set = true; }
    URIR_CUDA_OK(launch_pdl(kern, grid, dim3(192), smem, st, maps, p));
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

int conv_wgrad_halo(const urir_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    WhaloMaps maps; WhaloParams p; memset(&p, 0, sizeof(p));
    const int KN = d->K == 32 ? 32 : 64;              // dy channels per CTA; grid.z covers the rest
    p.C = d->C; p.K = d->K; p.Cc = d->C == 32 ? 32 : 64;
    p.n_mma = p.Cc == 32 ? 1 : 2;
    p.tiles_w = cdiv(d->W, WH_TW); p.tiles_h = cdiv(d->H, WH_TH); p.total_tiles = p.tiles_w * p.tiles_h * d->N;
    const int a_box = WH_PH * WH_TW * p.Cc * 2, b_box = WH_TH * WH_PW * KN * 2;
    p.a_bytes = (a_box + 1023) / 1024 * 1024;
    p.stage_bytes = p.a_bytes + (b_box + 1023) / 1024 * 1024;
    p.tx_bytes = a_box + b_box;
    p.stages = WH_SMEM_BUDGET / p.stage_bytes;
    if (p.stages > WH_MAX_STAGES) p.stages = WH_MAX_STAGES;
    { const int cols = p.n_mma * 3 * KN; p.tmem_cols = cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }
    p.dw = dw; p.gate = next_gate();
    { const char* e = getenv("URIR_WH_DEBUG"); p.debug = e ? atoi(e) : 0; }
    {   // x: dims (C, H, W, N), H the fastest pixel index of the box
        const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->H, (uint64_t)d->W, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->W * d->x_ld * 2, (uint64_t)d->x_ld * 2, (uint64_t)d->H * d->W * d->x_ld * 2};
        const uint32_t box[4] = {(uint32_t)p.Cc, (uint32_t)WH_PH, (uint32_t)WH_TW, 1};
        int rc = encode_map(&maps.a, (const char*)x + (size_t)d->x_coff * 2, 4, dims, strides, box, p.Cc * 2);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->P, (uint64_t)d->Q, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->Q * d->y_ld * 2, (uint64_t)d->y_ld * 2, (uint64_t)d->P * d->Q * d->y_ld * 2};
        const uint32_t box[4] = {(uint32_t)KN, (uint32_t)WH_TH, (uint32_t)WH_PW, 1};
        int rc = encode_map(&maps.b, (const char*)dy + (size_t)d->y_coff * 2, 4, dims, strides, box, KN * 2);
        if (rc) return rc;
    }
    if (!d->accumulate) URIR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)9 * d->C * d->K, st));
    const int n_cblk = d->C / p.Cc, n_kblk = d->K / KN;
    int gx = sm_count() / (n_cblk * n_kblk); if (gx < 1) gx = 1;
    if (gx > p.total_tiles) gx = p.total_tiles;
    dim3 grid(gx, n_cblk, n_kblk);
    const int smem = p.stages * p.stage_bytes + 1024 + (2 * WH_MAX_STAGES + 2) * 8 + 16 + 1024;
    if (KN == 32) return launch_wh<32>(maps, p, grid, smem, st);
    if (KN == 64) return launch_wh<64>(maps, p, grid, smem, st);
    return fail(URIR_ERR_UNSUP, "wgrad(halo): K = %d not supported", KN);
}

// =============================================================================================
// Stride-2 3x3 weight gradient (strided Conv2D / Conv2DTranspose kernels at the wide resolutions):
//   dw[r][s][c][k] = sum over (p, q) of  x[2p + r, 2q + s][c] * dy[p, q][k]          (SAME pad 0 before, 1 after)
// Tap (r, s) reads parity plane (r & 1, s & 1) of x at offset (a, b) = (r >> 1, s >> 1) on the half-resolution
// grid. Per 8 x 16 tile of that grid a stage holds the four x planes as {32 ch, 8+1 rows, 16 cols} boxes (strided
// tensor maps) and dy as {64 ch, 8 rows, 1+16 cols} boxes; with q' = q + b the sum becomes
//   dw[r][s] = sum over (p, q') of  plane[p + a, q'] * dy[p, q' - b].
// GEMM-M stacks the four planes (atom j = plane j, LBO = plane spacing): accumulator 0 starts at row p (a = 0),
// accumulator 1 one row lower (a = 1; only planes 0, 1 = row parity 0 are meaningful). GEMM-N stacks b: atom 0 is
// dy one column to the left (b = 1), atom 1 dy itself (b = 0). Two N = 128 instructions per 16 pixels and per
// 64-channel half of dy produce all nine taps; x and dy are staged once (conv_wgrad_tc.cu: 9 x and 3 x).
// =============================================================================================
constexpr int WS_PLANE_BYTES = (WH_TH + 1) * WH_TW * 64;                                   // 9216 = 9 * 1024
constexpr int WS_DY_BYTES = ((WH_TH * (WH_TW + 1) * 128) + 1023) / 1024 * 1024;            // 17408 -> 17408 (17 * 1024)

struct WhaloS2Params {
    int tiles_w, tiles_h, total_tiles;
    int C, K;
    int nh;                     // 64-channel halves of dy (1 or 2)
    int stages, stage_bytes, tx_bytes;
    int tmem_cols;
    float* dw;
    unsigned int* gate;         // deterministic mode (urir_common.cuh)
};
struct WhaloS2Maps { CUtensorMap a[4]; CUtensorMap b; };

__global__ void __launch_bounds__(192, 1)
conv_wgrad_halo_s2_kernel(const __grid_constant__ WhaloS2Maps maps, const __grid_constant__ WhaloS2Params p) {
    constexpr uint32_t IDESC = make_idesc_bf16(128, 128, 1, 1);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes + 1024);
    uint64_t* empty_bar = full_bar + WH_MAX_STAGES;
    uint64_t* tmem_full_bar = empty_bar + WH_MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int cblk = blockIdx.y;
    const int STAGES = p.stages;
    const int n_iters = blockIdx.x < p.total_tiles ? (p.total_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    if (warp == 4 && lane == 0) { prefetch_tmap(&maps.a[0]); prefetch_tmap(&maps.b); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    if (warp == 4) {
        // ===================== TMA producer: 4 x planes + nh dy boxes per tile =====================
        int stage = 0; uint32_t phase = 0;
        uint8_t* dst = smem;
        const int c0 = cblk * 32;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            int t = tile;
            const int tw = t % p.tiles_w; t /= p.tiles_w;
            const int th = t % p.tiles_h; const int n = t / p.tiles_h;
            const int p0 = th * WH_TH, q0 = tw * WH_TW;
            mbar_wait(empty_bar + stage, phase ^ 1);
            mbar_expect_tx_elect(full_bar + stage, (uint32_t)p.tx_bytes);
#pragma unroll
            for (int pl = 0; pl < 4; ++pl)
                tma_load_4d_elect(&maps.a[pl], full_bar + stage, dst + pl * WS_PLANE_BYTES, c0, p0, q0, n);
            tma_load_4d_elect(&maps.b, full_bar + stage, dst + 4 * WS_PLANE_BYTES, 0, p0, q0 - 1, n);
            if (p.nh > 1) tma_load_4d_elect(&maps.b, full_bar + stage, dst + 4 * WS_PLANE_BYTES + WS_DY_BYTES, 64, p0, q0 - 1, n);
            dst += p.stage_bytes;
            if (++stage == STAGES) { stage = 0; phase ^= 1; dst = smem; }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
        constexpr uint32_t A_COL = (WH_TH + 1) * 64, B_COL = WH_TH * 128;         // bytes per pixel column of the boxes
        const uint32_t a_hi = (A_COL >> 4) | (1u << 14) | (SWZ_64B << 29);         // SBO = next column of a plane
        const uint32_t b_hi = (B_COL >> 4) | (1u << 14) | (SWZ_128B << 29);        // SBO = next column of dy
        const uint32_t a_lo0 = ((WS_PLANE_BYTES >> 4) << 16) | (smem_u32(smem) >> 4);                  // LBO = next plane
        const uint32_t b_lo0 = ((B_COL >> 4) << 16) | ((smem_u32(smem) + 4 * WS_PLANE_BYTES) >> 4);    // LBO = one column right
        const uint32_t stage16 = p.stage_bytes >> 4;
        const int nh = p.nh;
        int stage = 0; uint32_t phase = 0, soff = 0;
        for (int it = 0; it < n_iters; ++it) {
            mbar_wait(full_bar + stage, phase);
            fence_after_sync();
#pragma unroll
            for (int k = 0; k < WH_TW / 2; ++k) {
                const uint32_t acc = (it != 0) | (k != 0);
                const uint64_t ad0 = ((uint64_t)a_hi << 32) | (a_lo0 + soff + k * ((2 * A_COL) >> 4));
                const uint64_t ad1 = ad0 + (64 >> 4);                              // one row lower: a = 1
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h < nh) {
                        const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo0 + soff + h * (WS_DY_BYTES >> 4) + k * ((2 * B_COL) >> 4));
                        umma_bf16_elect(tm0 + (h * 2 + 0) * 128, ad0, bd, IDESC, acc);
                        umma_bf16_elect(tm0 + (h * 2 + 1) * 128, ad1, bd, IDESC, acc);
                    }
                }
            }
            umma_commit_elect(empty_bar + stage);
            if (it == n_iters - 1) umma_commit_elect(tmem_full_bar);
            __syncwarp();
            soff += stage16;
            if (++stage == STAGES) { stage = 0; phase ^= 1; soff = 0; }
        }
    } else if (warp < 4) {
        // ===================== epilogue =====================
        if (n_iters > 0) { mbar_wait(tmem_full_bar, 0); fence_after_sync(); }
        gate_enter(p.gate, cta_linear(), threadIdx.x == 0, 1, 128);
        if (n_iters > 0) {
            const int row = warp * 32 + lane;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
            const int j = row >> 5, c = row & 31, ph = j >> 1, pw = j & 1;
            const int n_units = p.nh * 4;
            const int u0 = blockIdx.x % n_units;
#pragma unroll 1
            for (int uu = 0; uu < n_units; ++uu) {
                int u = uu + u0; if (u >= n_units) u -= n_units;
                const int h = u >> 2, m = (u >> 1) & 1, i = u & 1;                 // dy half, accumulator (a), B atom
                const int r = m == 0 ? ph : 2, s = pw + 2 * (1 - i);
                const bool valid = s <= 2 && (m == 0 || j < 2);
                float* drow = p.dw + ((size_t)((valid ? r : 0) * 3 + (valid ? s : 0)) * p.C + cblk * 32 + c) * p.K + h * 64;
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    uint32_t rg[16];
                    tmem_ld16(lane_addr + (h * 2 + m) * 128 + i * 64 + c0, rg);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 16; q += 4)
                            wh_red_add_v4(drow + c0 + q, __uint_as_float(rg[q]), __uint_as_float(rg[q + 1]),
                                          __uint_as_float(rg[q + 2]), __uint_as_float(rg[q + 3]));
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

bool wgrad_halo_s2_supported(const urir_conv_desc* d, bool forced) {
    if (d->stride != 2 || d->R != 3 || d->S != 3 || d->pad_top != 0 || d->pad_left != 0) return false;
    if (d->H != 2 * d->P || d->W != 2 * d->Q) return false;
    if (d->x_dtype != URIR_BF16 || d->y_dtype != URIR_BF16) return false;
    if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return false;
    if (d->C % 32 || !(d->K == 64 || d->K == 128)) return false;
    const long long tiles = (long long)d->N * cdiv(d->P, WH_TH) * cdiv(d->Q, WH_TW);
    return forced || tiles >= sm_count() * 4;
}

int conv_wgrad_halo_s2(const urir_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    WhaloS2Maps maps; WhaloS2Params p; memset(&p, 0, sizeof(p));
    p.C = d->C; p.K = d->K; p.nh = d->K / 64;
    p.tiles_w = cdiv(d->Q, WH_TW); p.tiles_h = cdiv(d->P, WH_TH); p.total_tiles = p.tiles_w * p.tiles_h * d->N;
    p.stage_bytes = 4 * WS_PLANE_BYTES + p.nh * WS_DY_BYTES;
    p.tx_bytes = 4 * (WH_TH + 1) * WH_TW * 64 + p.nh * WH_TH * (WH_TW + 1) * 128;
    p.stages = WH_SMEM_BUDGET / p.stage_bytes;
    if (p.stages > WH_MAX_STAGES) p.stages = WH_MAX_STAGES;
    p.tmem_cols = p.nh == 1 ? 256 : 512;
    p.dw = dw; p.gate = next_gate();
    for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
            const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)d->P, (uint64_t)d->Q, (uint64_t)d->N};
            const uint64_t strides[3] = {2ull * d->W * d->x_ld * 2, 2ull * d->x_ld * 2, (uint64_t)d->H * d->W * d->x_ld * 2};
            const uint32_t box[4] = {32, (uint32_t)(WH_TH + 1), (uint32_t)WH_TW, 1};
            const char* base = (const char*)x + ((size_t)d->x_coff + ((size_t)ph * d->W + pw) * d->x_ld) * 2;
            int rc = encode_map(&maps.a[ph * 2 + pw], base, 4, dims, strides, box, 64);
            if (rc) return rc;
        }
    {
        const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->P, (uint64_t)d->Q, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->Q * d->y_ld * 2, (uint64_t)d->y_ld * 2, (uint64_t)d->P * d->Q * d->y_ld * 2};
        const uint32_t box[4] = {64, (uint32_t)WH_TH, (uint32_t)(WH_TW + 1), 1};
        int rc = encode_map(&maps.b, (const char*)dy + (size_t)d->y_coff * 2, 4, dims, strides, box, 128);
        if (rc) return rc;
    }
    if (!d->accumulate) URIR_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)9 * d->C * d->K, st));
    const int n_cblk = d->C / 32;
    int gx = sm_count() / n_cblk; if (gx < 1) gx = 1;
    if (gx > p.total_tiles) gx = p.total_tiles;
    dim3 grid(gx, n_cblk);
    const int smem = p.stages * p.stage_bytes + 1024 + (2 * WH_MAX_STAGES + 2) * 8 + 16 + 1024;
    static std::atomic<bool> attr_set{false};    // benign if two threads both set the attribute
    if (!attr_set) { URIR_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_halo_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_SMEM_BUDGET + 8192)); attr_set = true; }
    URIR_CUDA_OK(launch_pdl(conv_wgrad_halo_s2_kernel, grid, dim3(192), smem, st, maps, p));
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

}  // namespace urir
