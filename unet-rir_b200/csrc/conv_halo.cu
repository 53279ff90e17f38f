// Persistent halo-tile implicit GEMM for the stride-1 convolutions at the wide resolutions
// (convolutional_block_1, dl_models/u_net.py:363-371, and their input gradients under tape.gradient,
// amp_phase_trainer.py:138): E1b / D5a / D5b at 144x160 and E2b / D4a / D4b at 72x80.
//
// conv_igemm.cu runs every tap as a GEMM-K iteration over its own re-fetched activation tile, one tile per
// CTA. At these resolutions that is bound by TMA latency (the 4-deep ring of a short-lived CTA) and by
// L2 -> SM bandwidth (9 x the activation bytes). Here:
//   * ONE TMA box per (tile, channel chunk) brings the tile WITH its halo: {BLOCK_K ch, 8+R-1 rows, 16+S-1
//     cols}, H fastest in shared memory. Tap (dh, dw) is not another load but another UMMA descriptor: it
//     starts (dw * PH + dh) rows into the box with SBO = PH rows, which walks the 16 groups of 8 vertically
//     adjacent pixels of the 8 (H) x 16 (W) output tile. The hardware swizzle is a function of the shared
//     memory address, so any row offset works (profiles/r01_umma_halo_probe*.txt).
//   * the whole weight tensor of the layer (all taps, <= 144 KB) is loaded once per CTA and stays resident;
//   * CTAs are persistent (one per SM): the TMA ring runs ahead across tiles, two TMEM accumulator stages
//     let the epilogue of tile i overlap the MMAs of tile i+1.
// Warp roles (384 threads): warp 4 TMA producer; warps 5-6 MMA issuers alternating tiles; warps 0-3 and 8-11
// two epilogue groups (TMEM lane quarter = warp % 4), group g draining TMEM stage g = the tiles of issuer g.
// Epilogue as in conv_igemm.cu: +bias, per-channel sum / sum of squares (BatchNorm statistics or bias
// gradients), bf16, 32-byte stores into the (possibly channel-sliced) NHWC destination.
#include <stdlib.h>
#include <atomic>
#include "urir_common.cuh"
#include "urir_tc.cuh"

namespace urir {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes);
bool stats_sums_only();          // set around a urir_conv2d_dgrad_sums call (capi.cu)

constexpr int HL_TH = 8, HL_TW = 16;          // output tile: 8 rows x 16 columns = 128 GEMM rows
constexpr int HL_MAX_STAGES = 8;
constexpr int HL_MAX_FSETS = 8;            // sets of `full` barriers (see HaloParams::fsets)
constexpr int HL_SMEM_BUDGET = 220 * 1024;

struct HaloParams {
    int N, H, W;
    int tiles_h, tiles_w, total_tiles;
    int PH, PW;                 // halo box extent (rows, cols)
    int oh0, ow0;               // box origin relative to the tile origin
    int ntaps, nchunks;         // taps, GEMM-K chunks of BLOCK_K channels
    int stages, a_stage_bytes, a_box_bytes, w_bytes;
    int one_issuer;             // 1: warp 5 issues every tile (fallback when fsets would exceed HL_MAX_FSETS)
    int fsets;                  // `full` barrier sets: fill number `pass` of a stage signals set pass % fsets (see kernel)
    int nplanes, plane_bytes;   // boxes per stage (1, or the 4 parity planes of a stride-2 input) and their spacing
    long long o_sn, o_sh, o_sw; // output element strides
    long long o_off;
    const float* bias;
    float* stats;
    __nv_bfloat16* out;
    int n_total;
    int bias_mod;               // bias index = GEMM-N column % bias_mod (the 4 parity classes of an up-2 layer share it)
    int accumulate;             // out += result (fp32 add before the bf16 rounding)
    int sums_only;              // statistics: channel sums only, no sums of squares (urir_conv2d_dgrad_sums: bias gradients)
    int grp_off[16];            // output element offset of every 32-column group of GEMM-N (channel / parity placement)
    unsigned int* gate;         // deterministic mode: CTAs commit their statistics in blockIdx order (urir_common.cuh)
    int debug;                  // URIR_HALO_DEBUG: 1 no global stores, 2 no epilogue math/stores, 3 no MMAs, 4 no TMA loads (timing experiments)
    long long* trace;           // debug (URIR_HALO_TRACE): clock64 stamps of CTA 0, 8 per tile, first 128 tiles
    short tap_row[36];          // first box row of tap t = dw * PH + dh
    short wtap[36];             // weight tap index of tap t
};

struct HaloMaps { CUtensorMap a[4]; CUtensorMap b; };

// Column sums of a [32 lanes] x [16 columns] block by recursive halving: step h exchanges half of the
// values with lane ^ (16 >> h) and adds, so after all five steps lanes with (lane & 1) == 0 hold the total of
// column hl_col_of_lane(lane). The steps are linear, so the epilogue runs only the first HS steps per tile,
// accumulates the surviving 16 >> HS partial sums per block in registers across all tiles of the CTA, and
// finishes the remaining steps once at the end (no shared-memory atomics in the tile loop).
#define URIR_HALVE(V, OFF, CNT, BIT) { const bool up = lane & BIT; _Pragma("unroll") for (int j = 0; j < CNT; ++j) { \
        const float send = up ? V[j] : V[j + CNT]; const float keep = up ? V[j + CNT] : V[j]; \
        V[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF); } }
template <int HS>
__device__ __forceinline__ void hl_halve_head(float (&v)[16], int lane) {
    if (HS >= 1) URIR_HALVE(v, 16, 8, 16)
    if (HS >= 2) URIR_HALVE(v, 8, 4, 8)
    if (HS >= 3) URIR_HALVE(v, 4, 2, 4)
}
template <int HS>
__device__ __forceinline__ float hl_halve_tail(float* v, int lane) {     // v holds 16 >> HS partial sums
    if (HS < 1) URIR_HALVE(v, 16, 8, 16)
    if (HS < 2) URIR_HALVE(v, 8, 4, 8)
    if (HS < 3) URIR_HALVE(v, 4, 2, 4)
    URIR_HALVE(v, 2, 1, 2)
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// 4 x 4 transpose of 16-byte chunks across the 4 lanes of a group: in  pk[4c .. 4c+3] = chunk c of this lane's pixel,
// out pk[4i .. 4i+3] = chunk (lane & 3) of pixel i of the group.
__device__ __forceinline__ void hl_transpose4(uint32_t (&pk)[16], int lane) {
    {   // exchange with lane ^ 1 inside chunk pairs (0,1) and (2,3)
        const bool odd = lane & 1;
#pragma unroll
        for (int c = 0; c < 4; c += 2)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t send = odd ? pk[4 * c + e] : pk[4 * (c + 1) + e];
                const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 1);
                if (odd) pk[4 * c + e] = got; else pk[4 * (c + 1) + e] = got;
            }
    }
    {   // exchange with lane ^ 2 between chunk pairs (0,2) and (1,3)
        const bool hi = lane & 2;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t send = hi ? pk[4 * c + e] : pk[4 * (c + 2) + e];
                const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 2);
                if (hi) pk[4 * c + e] = got; else pk[4 * (c + 2) + e] = got;
            }
    }
}
__device__ __forceinline__ int hl_col_of_lane(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

// NG = epilogue groups (2: warps 0-3, 8-11; 4: additionally warps 12-15, 16-19 -- 640 threads, <= 96 registers)
template <int BLOCK_N, int BLOCK_K, bool ACC, int NG, bool RELU>
__global__ void __launch_bounds__(NG == 4 ? 640 : 384, 1)
conv_halo_kernel(const __grid_constant__ HaloMaps maps, const __grid_constant__ HaloParams p) {
    constexpr uint32_t SWZ = (BLOCK_K == 64) ? SWZ_128B : SWZ_64B;
    constexpr uint32_t ROW_BYTES = BLOCK_K * 2;
    constexpr uint32_t B_BYTES = BLOCK_N * ROW_BYTES;
    constexpr uint32_t TMEM_COLS = 4 * BLOCK_N;             // four accumulator stages (two per issuer / epilogue group)
    constexpr uint32_t IDESC = make_idesc_bf16(128, BLOCK_N, 0, 0);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sA = smem + p.w_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + p.stages * p.a_stage_bytes);
    uint64_t* empty_bar = full_bar + HL_MAX_FSETS * HL_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + HL_MAX_STAGES;     // [4]
    uint64_t* tempty_bar = tfull_bar + 4;                // [4]
    uint64_t* w_bar = tempty_bar + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
    float* sstats = reinterpret_cast<float*>(tmem_slot + 2);     // [2 * BLOCK_N]
    float* sbias = sstats + 2 * BLOCK_N;                         // [BLOCK_N]

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int n_tile = blockIdx.y;
    const int STAGES = p.stages;
    long long trace_c0 = 0; unsigned long long trace_g0 = 0;
    if (p.trace && threadIdx.x == 0) { trace_c0 = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(trace_g0)); }

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(empty_bar + s, 1);
            for (int j = 0; j < p.fsets; ++j) mbar_init(full_bar + j * HL_MAX_STAGES + s, 1);
        }
        for (int a = 0; a < 4; ++a) { mbar_init(tfull_bar + a, 1); mbar_init(tempty_bar + a, 4); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == 4 && lane == 0) { prefetch_tmap(&maps.a[0]); prefetch_tmap(&maps.b); }
    for (int i = threadIdx.x; i < 2 * BLOCK_N; i += blockDim.x) sstats[i] = 0.f;
    for (int i = threadIdx.x; i < BLOCK_N; i += blockDim.x)
        sbias[i] = (p.bias && n_tile * BLOCK_N + i < p.n_total) ? p.bias[(n_tile * BLOCK_N + i) % p.bias_mod] : 0.f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();          // everything above overlapped the previous kernel's tail

    if (warp == 4) {
        // ===================== TMA producer =====================
        // resident weights: one box {BLOCK_K, BLOCK_N} per (tap, chunk), all on one barrier
        mbar_expect_tx_elect(w_bar, (uint32_t)(p.ntaps * p.nchunks) * B_BYTES);
        for (int t = 0; t < p.ntaps; ++t) {
            const int wt = __shfl_sync(0xffffffffu, (int)p.wtap[t], 0);
            for (int kc = 0; kc < p.nchunks; ++kc)
                tma_load_3d_elect(&maps.b, w_bar, sW + (size_t)(t * p.nchunks + kc) * B_BYTES, kc * BLOCK_K, n_tile * BLOCK_N, wt);
        }
        int stage = 0; uint32_t phase = 0;
        int fj = 0;                          // pass % fsets -> which set of full barriers this pass signals
        uint8_t* dst = sA;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            int t = tile;
            const int tw = t % p.tiles_w; t /= p.tiles_w;
            const int th = t % p.tiles_h; const int n = t / p.tiles_h;
            const int ch = th * HL_TH + p.oh0, cw = tw * HL_TW + p.ow0;
            const int itp = (tile - blockIdx.x) / gridDim.x;
            const bool trp = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && itp < 128 && lane == 0;
            for (int kc = 0; kc < p.nchunks; ++kc) {
                (void)trp;
                mbar_wait(empty_bar + stage, phase ^ 1);
                uint64_t* fb = full_bar + fj * HL_MAX_STAGES + stage;
                mbar_expect_tx_elect(fb, p.debug == 4 ? 0u : (uint32_t)(p.nplanes * p.a_box_bytes));
                if (p.debug != 4) tma_load_4d_elect(&maps.a[0], fb, dst, kc * BLOCK_K, ch, cw, n);
                if (p.nplanes > 1 && p.debug != 4) {
#pragma unroll
                    for (int pl = 1; pl < 4; ++pl)
                        tma_load_4d_elect(&maps.a[pl], fb, dst + pl * p.plane_bytes, kc * BLOCK_K, ch, cw, n);
                }

                dst += p.a_stage_bytes;
                if (++stage == STAGES) { stage = 0; phase ^= 1; dst = sA; if (++fj == p.fsets) fj = 0; }
            }
        }
    } else if (warp == 5 || warp == 6) {
        // ===================== MMA issuers (warps 5 and 6) =====================
        // Two issuing warps alternate tiles (warp 5 + mw owns the tiles with it % 2 == mw and TMEM stage mw).
        // An mbarrier wait costs ~200 cycles even when the barrier is already complete and the tensor pipe's
        // issue queue is shallow, so a single issuer leaves the pipe idle ~400 cycles per tile (measured with
        // URIR_HALO_TRACE); with two, one warp's waits overlap the other's MMAs.
        const int mw = warp - 5;
        const int issue_end = (p.one_issuer && mw == 1) ? 0 : p.total_tiles;      // the idle second issuer
        const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t a_hi = (((uint32_t)p.PH * ROW_BYTES) >> 4) | (1u << 14) | (SWZ << 29);   // SBO = PH rows
        const uint32_t b_hi = ((8 * ROW_BYTES) >> 4) | (1u << 14) | (SWZ << 29);
        const uint32_t w_lo = smem_u32(sW) >> 4;
        const uint32_t a_lo0 = smem_u32(sA) >> 4;
        const uint32_t stage16 = (uint32_t)p.a_stage_bytes >> 4;
        mbar_wait(w_bar, 0);
        int stage = 0;
        int fj = 0; uint32_t fpar = 0;       // full-barrier set of the current pass, parity within that set
        uint32_t a_lo = a_lo0;
        int it = 0;
        // Stagger the issuers by half a tile of tensor-pipe time. Started together they fall into lock step: both
        // stream their MMAs at once (interleaved at the pipe's rate), then both sit in their ~600-cycle between-tile
        // gap (commits, barrier waits) at once and the pipe idles; URIR_HALO_TRACE showed 1140 cycles per tile
        // against 800 of MMA time. Half a period apart, one warp's gap hides behind the other's MMAs.
        if (mw == 1 && !p.one_issuer && p.debug == 7) {          // measured: the issuers fall back into lock step (epilogue-coupled); off
            const long long t0 = clock64();
            const long long delay = (long long)p.ntaps * p.nchunks * (BLOCK_K / 16) * (BLOCK_N <= 64 ? 24 : BLOCK_N <= 128 ? 32 : 64);
            while (clock64() - t0 < delay) { }
        }
        for (int tile = blockIdx.x; tile < issue_end; tile += gridDim.x, ++it) {
            // An mbarrier parity wait cannot tell "fill k+2 complete" from "fill k complete", so a waiter must observe
            // EVERY phase of a barrier it waits on. Tiles alternate between the two issuers, and unless the ring length
            // is a multiple of 2 * nchunks the issuer that consumes a given stage alternates between ring passes: with
            // one `full` barrier per stage each issuer would see only every other phase and (TMA completions being
            // unordered) could sail through a wait one pass early -- observed as a hang about once per thousand steps.
            // Hence `fsets` sets of full barriers: fill number `pass` of stage s signals set pass % fsets, with fsets
            // the smallest count for which fsets * STAGES is a multiple of 2 * nchunks. Then (stage, set) always
            // belongs to the same issuer, which sees each of its phases in order, and the empty-barrier handshake
            // keeps the producer from lapping it.
            if (!p.one_issuer && (it & 1) != mw) {           // the other issuer's tile: step over its stages
                for (int kc = 0; kc < p.nchunks; ++kc) {
                    a_lo += stage16;
                    if (++stage == STAGES) { stage = 0; a_lo = a_lo0; if (++fj == p.fsets) { fj = 0; fpar ^= 1; } }
                }
                continue;
            }
            const int acc = it & 3;                  // stages {mw, mw + 2}
            const bool trm = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && it < 128 && lane == 0;
            if (trm) p.trace[it * 8 + 0] = clock64();
            mbar_wait(tempty_bar + acc, ((it >> 2) & 1) ^ 1);
            if (trm) p.trace[it * 8 + 1] = clock64();
            const uint32_t d_tm = tm0 + acc * BLOCK_N;
            for (int kc = 0; kc < p.nchunks; ++kc) {
                mbar_wait(full_bar + fj * HL_MAX_STAGES + stage, fpar);
                if (trm && kc == p.nchunks - 1) p.trace[it * 8 + 2] = clock64();
                fence_after_sync();
                uint32_t b_lo = w_lo + ((kc * B_BYTES) >> 4);
                // (a fully unrolled 9-tap variant with register-resident offsets was measured twice -- round 1 and again
                // after the epilogue's spills were gone: per kernel within +-8 % either way, no gain at step level (3.67 vs
                // 3.65 ms). URIR_HALO_TRACE shows why: ~84 cycles per MMA on an issuer against 44 on the tensor pipe, but the
                // two issuers together already match the pipe; what stretches a tile from 913 to 1174 cycles is the epilogue
                // warps sharing the issuers' schedulers -- kept simple)
                for (int t = 0; t < (p.debug == 3 ? 0 : p.ntaps); ++t) {
                    const uint32_t a_t = a_lo + (((uint32_t)__shfl_sync(0xffffffffu, (int)p.tap_row[t], 0) * ROW_BYTES) >> 4);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k) {
                        const uint64_t ad = ((uint64_t)a_hi << 32) | (a_t + 2 * k);
                        const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + 2 * k);
                        umma_bf16_elect(d_tm, ad, bd, IDESC, (kc | t | k) != 0);
                    }
                    b_lo += (p.nchunks * B_BYTES) >> 4;
                }
                umma_commit_elect(empty_bar + stage);
                a_lo += stage16;
                if (++stage == STAGES) { stage = 0; a_lo = a_lo0; if (++fj == p.fsets) { fj = 0; fpar ^= 1; } }
            }
            umma_commit_elect(tfull_bar + acc);
            __syncwarp();
            if (trm) p.trace[it * 8 + 3] = clock64();
        }
    } else if (warp < 4 || warp >= 8) {
        // ===================== epilogue: group 0 = warps 0-3 (even tiles), group 1 = warps 8-11 (odd tiles) ====
        // one tile's epilogue is a ~900-cycle latency chain (barrier wait, TMEM loads, stores); two groups give
        // each of them two tile times to finish it
        const int eg = warp >= 8 ? ((warp - 8) >> 2) + 1 : 0;      // group eg drains the tiles with it % NG == eg
        const int quarter = warp & 3;
        // halving steps per tile (see hl_halve_head): the partial sums kept across tiles cost 2 * (BLOCK_N / 16) * (16 >> HS)
        // registers; at 64 of them the 64- and 128-column variants hit the 168-register cap and spilled 130 - 150 bytes per
        // thread inside the tile loop (ptxas -v), so they halve once more per tile (32 registers) at 8 more shuffles per block
        constexpr int HS = BLOCK_N >= 128 ? 3 : (BLOCK_N >= 64 || NG == 4) ? 2 : 0;     // (four groups: 96-register cap)
        constexpr int PART = 16 >> HS;                       // partial sums kept per 16-column block
        constexpr int NB = BLOCK_N / 16;
        float acc1[NB * PART], acc2[NB * PART];
#pragma unroll
        for (int i = 0; i < NB * PART; ++i) { acc1[i] = 0.f; acc2[i] = 0.f; }
        const int row = quarter * 32 + lane;
        const int ih = row & 7, iw = row >> 3;
        const bool want_stats = p.stats != nullptr, want_sq = want_stats && !p.sums_only;
        int it = eg;
        for (int tile = blockIdx.x + eg * gridDim.x; tile < p.total_tiles; tile += NG * gridDim.x, it += NG) {
            const int acc = it & 3;
            int t = tile;
            const int tw = t % p.tiles_w; t /= p.tiles_w;
            const int th = t % p.tiles_h; const int n = t / p.tiles_h;
            const int h = th * HL_TH + ih, w = tw * HL_TW + iw;
            const bool valid = h < p.H && w < p.W;
            // the 4 pixels whose chunks this lane stores after the transpose: rows 4 * (lane / 4) + i of the quarter
            __nv_bfloat16* gout[4]; bool gvalid[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = quarter * 32 + (lane & ~3) + i;
                const int hh = th * HL_TH + (m & 7), ww = tw * HL_TW + (m >> 3);
                gvalid[i] = hh < p.H && ww < p.W;
                gout[i] = p.out + p.o_off + (long long)n * p.o_sn + (long long)hh * p.o_sh + (long long)ww * p.o_sw;
            }
            uint32_t pk[16];
            // out += (accumulate): the previous contents are fetched in the SAME transposed layout the stores use (lane
            // group of 4 = one pixel's 64 contiguous bytes per 32-channel group), one 64-column phase ahead (before
            // the accumulator barrier / during the previous phase), and added after the transpose. The sum is formed
            // in fp32 from the bf16-rounded result and the old bf16 value (one extra rounding, well inside the bf16
            // tolerance); per-lane loads of the lane's own pixel cost half-empty sectors and measured 2x slower.
            constexpr int PH_COLS = BLOCK_N < 64 ? BLOCK_N : 64;
            uint4 oldv[ACC ? PH_COLS / 8 : 1];                     // [32-channel group of the phase][pixel i of the lane group]
            auto load_old = [&](int ph) {
#pragma unroll
                for (int gi = 0; gi < PH_COLS / 32; ++gi) {
                    const int c32 = p.grp_off[n_tile * (BLOCK_N / 32) + ph * (PH_COLS / 32) + gi] + (lane & 3) * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        oldv[gi * 4 + i] = gvalid[i] ? *reinterpret_cast<const uint4*>(gout[i] + c32) : make_uint4(0, 0, 0, 0);
                }
            };
            if (ACC) load_old(0);
            const bool tre = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && it < 128 && quarter == 0 && lane == 0;
            mbar_wait(tfull_bar + acc, (it >> 2) & 1);
            if (tre) p.trace[it * 8 + 4] = clock64();
            fence_after_sync();
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
            // TMEM loads have a long latency while the tensor pipe is streaming MMAs (the dominant epilogue stall
            // in the ncu source view), so all loads of a phase (<= 64 columns) are issued before one wait.
#pragma unroll
            for (int ph = 0; ph < (p.debug == 2 ? 0 : BLOCK_N / PH_COLS); ++ph) {
                uint32_t r[PH_COLS];
#pragma unroll
                for (int c = 0; c < PH_COLS / 32; ++c) tmem_ld32(lane_addr + ph * PH_COLS + c * 32, r + c * 32);
                tmem_ld_wait();
                if (tre && ph == 0) p.trace[it * 8 + 6] = clock64();
                if (ph == BLOCK_N / PH_COLS - 1 && p.debug != 8) {
                    // the tile's last columns are in registers: hand the accumulator stage back to the MMA issuer NOW,
                    // not after the arithmetic and the global stores below (they were ~half of the time the stage was held)
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar + acc);
                }
#pragma unroll
                for (int bb = 0; bb < PH_COLS / 16; ++bb) {
                    const int b = ph * (PH_COLS / 16) + bb;
                    float v[16];
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 bs = *reinterpret_cast<const float4*>(sbias + b * 16 + 4 * j4);
                        v[4 * j4] = __uint_as_float(r[bb * 16 + 4 * j4]) + bs.x; v[4 * j4 + 1] = __uint_as_float(r[bb * 16 + 4 * j4 + 1]) + bs.y;
                        v[4 * j4 + 2] = __uint_as_float(r[bb * 16 + 4 * j4 + 2]) + bs.z; v[4 * j4 + 3] = __uint_as_float(r[bb * 16 + 4 * j4 + 3]) + bs.w;
                    }
                    if (RELU) {              // inference: BatchNorm folded into weights / bias (urir_weight_fold_bn_batched)
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    pk[(b & 1) * 8 + 0] = pack_bf16x2(v[0], v[1]); pk[(b & 1) * 8 + 1] = pack_bf16x2(v[2], v[3]);
                    pk[(b & 1) * 8 + 2] = pack_bf16x2(v[4], v[5]); pk[(b & 1) * 8 + 3] = pack_bf16x2(v[6], v[7]);
                    pk[(b & 1) * 8 + 4] = pack_bf16x2(v[8], v[9]); pk[(b & 1) * 8 + 5] = pack_bf16x2(v[10], v[11]);
                    pk[(b & 1) * 8 + 6] = pack_bf16x2(v[12], v[13]); pk[(b & 1) * 8 + 7] = pack_bf16x2(v[14], v[15]);
                    if (b & 1) {
                        // 32 channels = four 16-byte chunks per pixel. Transpose them across the 4 lanes of a group so
                        // that store i writes chunk (lane & 3) of pixel (group, i): 64 contiguous bytes per lane group
                        // (full sectors) instead of four scattered 16-byte pieces.
                        // (storing each lane's own 64 bytes without the transpose was measured 15-20 % slower at 64 channels)
                        {
                        hl_transpose4(pk, lane);
                        if (ACC) {
                            const int gi = bb >> 1;                // 32-channel group within the phase
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t ow[4] = {oldv[gi * 4 + i].x, oldv[gi * 4 + i].y, oldv[gi * 4 + i].z, oldv[gi * 4 + i].w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 a = unpack_bf16x2(pk[4 * i + j]), o = unpack_bf16x2(ow[j]);
                                    pk[4 * i + j] = pack_bf16x2(a.x + o.x, a.y + o.y);
                                }
                            }
                            if (gi == PH_COLS / 32 - 1 && ph + 1 < BLOCK_N / PH_COLS) load_old(ph + 1);
                        }
                        const int c32 = p.grp_off[n_tile * (BLOCK_N / 32) + (b >> 1)] + (lane & 3) * 8;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (gvalid[i] && p.debug != 1) *reinterpret_cast<uint4*>(gout[i] + c32) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                        }
                    }
                    if (want_stats) {
                        float q[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) { if (!valid) v[j] = 0.f; q[j] = v[j] * v[j]; }
                        hl_halve_head<HS>(v, lane);
#pragma unroll
                        for (int j = 0; j < PART; ++j) acc1[b * PART + j] += v[j];
                        if (want_sq) {
                            hl_halve_head<HS>(q, lane);
#pragma unroll
                            for (int j = 0; j < PART; ++j) acc2[b * PART + j] += q[j];
                        }
                    }
                }
            }
            if (tre) p.trace[it * 8 + 7] = clock64();
            if (p.debug == 2 || p.debug == 8) {        // (no TMEM loads issued in debug 2; debug 8 = round 1's late release)
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + acc);
            }
            if (tre) p.trace[it * 8 + 5] = clock64();
        }
        if (want_stats) {
            float s1[NB], s2[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                s1[b] = hl_halve_tail<HS>(acc1 + b * PART, lane);
                s2[b] = hl_halve_tail<HS>(acc2 + b * PART, lane);
            }
            // the NG * 4 epilogue warps fold their column sums into shared memory one warp at a time (fixed order:
            // no floating-point atomics, the result does not depend on warp timing); once per kernel
            const int me = eg * 4 + quarter;
#pragma unroll 1
            for (int turn = 0; turn < NG * 4; ++turn) {
                if (turn == me && (lane & 1) == 0) {
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const int col = b * 16 + hl_col_of_lane(lane);
                        sstats[col] += s1[b];
                        sstats[BLOCK_N + col] += s2[b];
                    }
                }
                gate_bar(1, NG * 4 * 32);
            }
        }
    }
    __syncthreads();
    if (p.stats) {
        gate_enter(p.gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
        for (int i = threadIdx.x; i < (p.sums_only ? 1 : 2) * BLOCK_N; i += blockDim.x) {
            const int which = i / BLOCK_N, col = i % BLOCK_N;
            if (n_tile * BLOCK_N + col < p.n_total)
                atomicAdd(p.stats + which * p.n_total + n_tile * BLOCK_N + col, sstats[i]);
        }
        gate_leave(p.gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
    }
    if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, TMEM_COLS); }
    if (p.trace && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 148) {
        unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
        p.trace[1024 + 2 * blockIdx.x] = clock64() - trace_c0;
        p.trace[1024 + 2 * blockIdx.x + 1] = (long long)(g1 - trace_g0);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int halo_block_k(int kg) { return kg % 64 == 0 ? 64 : 32; }
static int halo_block_n(int n) { return n % 128 == 0 ? 128 : n % 64 == 0 ? 64 : 32; }
// largest GEMM-N tile whose resident weights (all taps, kg channels) leave room for >= 3 activation stages:
// a narrower tile re-reads the activation once per N tile but keeps the halo reuse and the persistent pipeline
static int halo_block_n_fit(int ng, int kg, int ntaps, int ph, int pw) {
    const int BK = halo_block_k(kg);
    const long long a_stage = (((long long)ph * pw * BK * 2) + 1023) / 1024 * 1024;
    for (int bn = halo_block_n(ng); bn >= 32; bn >>= 1)
        if ((long long)ntaps * kg * bn * 2 + 3 * a_stage + 2048 <= HL_SMEM_BUDGET) return bn;
    return 0;
}

// Number of `full` barrier sets that makes (stage, set) ownership static for the two issuers (see the kernel's
// issuer loop); a single issuer if that would need more than HL_MAX_FSETS sets.
static void halo_fix_stages(HaloParams& p) {
    if (p.stages > HL_MAX_STAGES) p.stages = HL_MAX_STAGES;
    const int q = 2 * p.nchunks;
    p.fsets = 1;
    while ((p.fsets * p.stages) % q != 0) ++p.fsets;
    p.one_issuer = 0;
    if (p.fsets > HL_MAX_FSETS) { p.fsets = 1; p.one_issuer = 1; }
}

// op 0: fprop (GEMM-K = C, GEMM-N = K), op 1: dgrad (GEMM-K = K, GEMM-N = C)
bool halo_supported(const urir_conv_desc* d, int op, bool forced) {
    if (d->stride != 1 || d->P != d->H || d->Q != d->W || d->x_dtype != URIR_BF16 || d->y_dtype != URIR_BF16) return false;
    if ((d->act != URIR_ACT_NONE && !(d->act == URIR_ACT_RELU && op == 0)) || d->accumulate) return false;
    if (d->R > 6 || d->S > 6 || d->R * d->S > 36) return false;
    if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return false;
    const int kg = op == 0 ? d->C : d->K, ng = op == 0 ? d->K : d->C;
    if (kg % 32 || ng % 32) return false;
    const int BN = halo_block_n_fit(ng, kg, d->R * d->S, HL_TH + d->R - 1, HL_TW + d->S - 1);
    if (BN == 0) return false;
    if (BN < 64 && BN < ng && !forced) return false;      // 44-cycle N = 32 MMAs over several N tiles do not pay
    // worth it only where the one-tile-per-CTA kernel is latency / L2 bound: many tiles per SM
    const long long tiles = (long long)d->N * cdiv(d->H, HL_TH) * cdiv(d->W, HL_TW);
    return forced || tiles >= sm_count() * 4;
}

template <int BN, int BK, bool ACC = false, int NG = 2, bool RELU = false>
static int launch_halo(const HaloMaps& maps, const HaloParams& p, int n_tiles, int smem, cudaStream_t st) {
    static std::atomic<bool> attr_set{false};    // benign if two threads both set the attribute
    auto kern = conv_halo_kernel<BN, BK, ACC, NG, RELU>;
    if (!attr_set) { URIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HL_SMEM_BUDGET + 4096)); attr_set = true; }
    int gx = sm_count() / n_tiles; if (gx < 1) gx = 1;
    if (gx > p.total_tiles) gx = p.total_tiles;
    dim3 grid(gx, n_tiles);
    URIR_CUDA_OK(launch_pdl(kern, grid, dim3(NG == 4 ? 640 : 384), smem, st, maps, p));
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

// a = activation side read through the halo box (x for fprop, dy for dgrad); w = [tap][GEMM-N][GEMM-K] bf16
int conv_halo(const urir_conv_desc* d, int op, const void* a, const void* w, const float* bias, void* out, float* stats,
              cudaStream_t st) {
    URIR_CHECK_ARG(w != nullptr, "halo conv needs the [tap][N][K] weight layout");
    const int kg = op == 0 ? d->C : d->K, ng = op == 0 ? d->K : d->C;
    const int a_ld = op == 0 ? d->x_ld : d->y_ld, a_coff = op == 0 ? d->x_coff : d->y_coff;
    const int o_ld = op == 0 ? d->y_ld : d->x_ld, o_coff = op == 0 ? d->y_coff : d->x_coff;
    const int BK = halo_block_k(kg), BN = halo_block_n_fit(ng, kg, d->R * d->S, HL_TH + d->R - 1, HL_TW + d->S - 1);
    if (BN == 0) return fail(URIR_ERR_UNSUP, "halo conv: the weights of one 32-column tile do not fit in shared memory");
    HaloMaps maps; HaloParams p; memset(&p, 0, sizeof(p));
    p.N = d->N; p.H = d->H; p.W = d->W;
    p.tiles_h = cdiv(d->H, HL_TH); p.tiles_w = cdiv(d->W, HL_TW); p.total_tiles = p.tiles_h * p.tiles_w * d->N;
    p.PH = HL_TH + d->R - 1; p.PW = HL_TW + d->S - 1;
    p.ntaps = d->R * d->S; p.nchunks = kg / BK;
    p.a_box_bytes = p.PH * p.PW * BK * 2;
    p.a_stage_bytes = (p.a_box_bytes + 1023) / 1024 * 1024;
    p.nplanes = 1; p.plane_bytes = p.a_stage_bytes;
    p.w_bytes = p.ntaps * kg * BN * 2;
    p.stages = (HL_SMEM_BUDGET - 2048 - p.w_bytes) / p.a_stage_bytes;
    halo_fix_stages(p);
    if (p.stages < 2) return fail(URIR_ERR_UNSUP, "halo conv: weights of %d bytes leave no room for the activation ring", p.w_bytes);
    p.o_sn = (long long)d->H * d->W * o_ld; p.o_sh = (long long)d->W * o_ld; p.o_sw = o_ld; p.o_off = o_coff;
    p.bias = bias; p.stats = stats; p.out = (__nv_bfloat16*)out; p.n_total = ng; p.gate = stats ? next_gate() : nullptr;
    p.bias_mod = ng; p.accumulate = 0; p.sums_only = stats_sums_only() ? 1 : 0;
    for (int g = 0; g < 16; ++g) p.grp_off[g] = 32 * g;
    { const char* e = getenv("URIR_HALO_TRACE"); p.trace = e ? (long long*)strtoull(e, nullptr, 16) : nullptr; }
    { const char* e = getenv("URIR_HALO_DEBUG"); p.debug = e ? atoi(e) : 0; }
    if (op == 0) { p.oh0 = -d->pad_top; p.ow0 = -d->pad_left; }
    else { p.oh0 = d->pad_top - (d->R - 1); p.ow0 = d->pad_left - (d->S - 1); }
    for (int r = 0; r < d->R; ++r)
        for (int s = 0; s < d->S; ++s) {
            const int t = r * d->S + s;
            const int dh = op == 0 ? r : d->R - 1 - r, dw = op == 0 ? s : d->S - 1 - s;
            p.tap_row[t] = (short)(dw * p.PH + dh);
            p.wtap[t] = (short)t;
        }
    {   // activation: dims (C, H, W, N) so that H is the fastest pixel index of the box in shared memory
        const uint64_t dims[4] = {(uint64_t)kg, (uint64_t)d->H, (uint64_t)d->W, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->W * a_ld * 2, (uint64_t)a_ld * 2, (uint64_t)d->H * d->W * a_ld * 2};
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.PH, (uint32_t)p.PW, 1};
        int rc = encode_map(&maps.a[0], (const char*)a + (size_t)a_coff * 2, 4, dims, strides, box, BK * 2);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)kg, (uint64_t)ng, (uint64_t)p.ntaps};
        const uint64_t strides[2] = {(uint64_t)kg * 2, (uint64_t)kg * ng * 2};
        const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BN, 1};
        int rc = encode_map(&maps.b, w, 3, dims, strides, box, BK * 2);
        if (rc) return rc;
    }
    const int smem = p.w_bytes + p.stages * p.a_stage_bytes + ((HL_MAX_FSETS + 1) * HL_MAX_STAGES + 9) * 8 + 16 + 3 * BN * 4 + 1024;
    const int n_tiles = ng / BN;
    if (d->act == URIR_ACT_RELU) {
#define URIR_HLR(BN_, BK_) if (BN == BN_ && BK == BK_) return launch_halo<BN_, BK_, false, 2, true>(maps, p, n_tiles, smem, st);
        URIR_HLR(32, 32) URIR_HLR(32, 64) URIR_HLR(64, 32) URIR_HLR(64, 64) URIR_HLR(128, 32) URIR_HLR(128, 64)
#undef URIR_HLR
    }
    { static int ng4 = -1; if (ng4 < 0) { const char* e = getenv("URIR_HALO_NG4"); ng4 = (e && e[0] == '1') ? 1 : 0; }
      if (ng4 && BN == 32 && BK == 32) return launch_halo<32, 32, false, 4>(maps, p, n_tiles, smem, st);
      if (ng4 && BN == 32 && BK == 64) return launch_halo<32, 64, false, 4>(maps, p, n_tiles, smem, st); }
#define URIR_HL(BN_, BK_) if (BN == BN_ && BK == BK_) return launch_halo<BN_, BK_>(maps, p, n_tiles, smem, st);
    URIR_HL(32, 32) URIR_HL(32, 64) URIR_HL(64, 32) URIR_HL(64, 64) URIR_HL(128, 32) URIR_HL(128, 64)
#undef URIR_HL
    return fail(URIR_ERR_UNSUP, "halo conv: no kernel for BLOCK_N=%d BLOCK_K=%d", BN, BK);
}

// ---------------------------------------------------------------------------------------------
// Stride-2 3x3 input gradient / Conv2DTranspose forward (dl_models/u_net.py:297-304) through the same kernel.
// With TF SAME padding on an even input (pad 0 before, 1 after) the forward conv reads x[2p + r, 2q + s], so
//   dx[2i + ph, 2j + pw, c] = sum over a, b in {0,1}, k of  dy[i - a, j - b, k] * w[2a + ph, 2b + pw, c, k]
// (terms with 2a + ph > 2 or 2b + pw > 2 do not exist). That is ONE 2x2 stride-1 input-gradient problem on the
// half-resolution grid whose GEMM-N index is (ph, pw, c): N = 4C, weights "w_up2" [a*2+b][(ph,pw,c)][K] with
// zeros for the missing taps (urir_weight_prep_up2). dy is read once through the halo box; the epilogue
// scatters the four parity classes of a pixel to (2i + ph, 2j + pw) -- 64-byte channel groups, so the
// transposed full-sector stores still apply. conv_igemm.cu instead launches one short-lived CTA per
// (128-pixel tile, parity class) with N = C (32 -> a 44-cycle MMA does a quarter of the work).
bool halo_up2_supported(const urir_conv_desc* d) {
    if (d->stride != 2 || d->R != 3 || d->S != 3 || d->pad_top != 0 || d->pad_left != 0) return false;
    if (d->H != 2 * d->P || d->W != 2 * d->Q) return false;
    if (d->x_dtype != URIR_BF16 || d->y_dtype != URIR_BF16 || d->act != URIR_ACT_NONE) return false;
    if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return false;
    if (d->C % 32 || d->K % 32 || 4 * d->C > 512) return false;
    const int BK = halo_block_k(d->K), BN = halo_block_n(4 * d->C);
    const long long w_bytes = 4LL * d->K * BN * 2;
    const long long a_stage = (((long long)(HL_TH + 1) * (HL_TW + 1) * BK * 2) + 1023) / 1024 * 1024;
    if (w_bytes + 3 * a_stage + 2048 > HL_SMEM_BUDGET) return false;
    return true;
}

int conv_halo_up2(const urir_conv_desc* d, const void* dy, const void* w_up2, const float* bias, void* dx, cudaStream_t st) {
    URIR_CHECK_ARG(w_up2 != nullptr, "up-2 halo conv needs the w_up2 weight layout");
    const int kg = d->K, ng = 4 * d->C;
    const int BK = halo_block_k(kg), BN = halo_block_n(ng);
    HaloMaps maps; HaloParams p; memset(&p, 0, sizeof(p));
    p.N = d->N; p.H = d->P; p.W = d->Q;                       // the half-resolution grid
    p.tiles_h = cdiv(p.H, HL_TH); p.tiles_w = cdiv(p.W, HL_TW); p.total_tiles = p.tiles_h * p.tiles_w * d->N;
    p.PH = HL_TH + 1; p.PW = HL_TW + 1;
    p.ntaps = 4; p.nchunks = kg / BK;
    p.a_box_bytes = p.PH * p.PW * BK * 2;
    p.a_stage_bytes = (p.a_box_bytes + 1023) / 1024 * 1024;
    p.nplanes = 1; p.plane_bytes = p.a_stage_bytes;
    p.w_bytes = p.ntaps * kg * BN * 2;
    p.stages = (HL_SMEM_BUDGET - 2048 - p.w_bytes) / p.a_stage_bytes;
    halo_fix_stages(p);
    if (p.stages < 2) return fail(URIR_ERR_UNSUP, "up-2 halo conv: weights of %d bytes leave no room for the activation ring", p.w_bytes);
    p.o_sn = (long long)d->H * d->W * d->x_ld; p.o_sh = 2LL * d->W * d->x_ld; p.o_sw = 2LL * d->x_ld; p.o_off = d->x_coff;
    p.bias = bias; p.stats = nullptr; p.out = (__nv_bfloat16*)dx; p.n_total = ng;
    p.bias_mod = d->C; p.accumulate = d->accumulate;
    for (int g = 0; g < ng / 32; ++g) {
        const int n0 = 32 * g, ph = n0 / (2 * d->C), pw = (n0 / d->C) % 2, c0 = n0 % d->C;
        p.grp_off[g] = (ph * d->W + pw) * d->x_ld + c0;
    }
    p.oh0 = -1; p.ow0 = -1;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int t = a * 2 + b;
            p.tap_row[t] = (short)((1 - b) * p.PH + (1 - a));
            p.wtap[t] = (short)t;
        }
    {
        const uint64_t dims[4] = {(uint64_t)kg, (uint64_t)d->P, (uint64_t)d->Q, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->Q * d->y_ld * 2, (uint64_t)d->y_ld * 2, (uint64_t)d->P * d->Q * d->y_ld * 2};
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.PH, (uint32_t)p.PW, 1};
        int rc = encode_map(&maps.a[0], (const char*)dy + (size_t)d->y_coff * 2, 4, dims, strides, box, BK * 2);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)kg, (uint64_t)ng, 4};
        const uint64_t strides[2] = {(uint64_t)kg * 2, (uint64_t)kg * ng * 2};
        const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BN, 1};
        int rc = encode_map(&maps.b, w_up2, 3, dims, strides, box, BK * 2);
        if (rc) return rc;
    }
    const int smem = p.w_bytes + p.stages * p.a_stage_bytes + ((HL_MAX_FSETS + 1) * HL_MAX_STAGES + 9) * 8 + 16 + 3 * BN * 4 + 1024;
    const int n_tiles = ng / BN;
    if (p.accumulate) {      // 4C is a multiple of 128: only the BLOCK_N = 128 kernels exist with the accumulate epilogue
        if (BN == 128 && BK == 32) return launch_halo<128, 32, true>(maps, p, n_tiles, smem, st);
        if (BN == 128 && BK == 64) return launch_halo<128, 64, true>(maps, p, n_tiles, smem, st);
        return fail(URIR_ERR_UNSUP, "up-2 halo conv: no accumulate kernel for BLOCK_N=%d BLOCK_K=%d", BN, BK);
    }
#define URIR_HL(BN_, BK_) if (BN == BN_ && BK == BK_) return launch_halo<BN_, BK_>(maps, p, n_tiles, smem, st);
    URIR_HL(32, 32) URIR_HL(32, 64) URIR_HL(64, 32) URIR_HL(64, 64) URIR_HL(128, 32) URIR_HL(128, 64)
#undef URIR_HL
    return fail(URIR_ERR_UNSUP, "up-2 halo conv: no kernel for BLOCK_N=%d BLOCK_K=%d", BN, BK);
}

// ---------------------------------------------------------------------------------------------
// Stride-2 3x3 forward (encoding_block's strided Conv2D, dl_models/u_net.py:269-276, and the input gradient of a
// Conv2DTranspose) through the same kernel. With SAME padding on an even input the conv reads x[2p + r, 2q + s]:
// tap (r, s) lives in parity plane (r & 1, s & 1) of x at offset (r >> 1, s >> 1). Each stage holds the four
// parity planes of the tile, every plane a (8+1) x (16+1) halo box fetched through its own strided tensor map
// (H/2 x W/2 view with doubled strides; the row / column past the end is TMA zero fill = the trailing SAME pad).
// The nine taps are descriptor offsets into those planes: x is read once per tile instead of once per tap, by a
// persistent CTA with resident weights.
bool halo_s2_fprop_supported(const urir_conv_desc* d, bool forced) {
    if (d->stride != 2 || d->R != 3 || d->S != 3 || d->pad_top != 0 || d->pad_left != 0) return false;
    if (d->H != 2 * d->P || d->W != 2 * d->Q) return false;
    if (d->x_dtype != URIR_BF16 || d->y_dtype != URIR_BF16 || d->act != URIR_ACT_NONE || d->accumulate) return false;
    if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return false;
    if (d->C % 32 || d->K % 32) return false;
    const int BK = 32;
    const long long stage = 4LL * ((((HL_TH + 1) * (HL_TW + 1) * BK * 2) + 1023) / 1024 * 1024);
    int BN = 0;
    for (int bn = halo_block_n(d->K); bn >= 32; bn >>= 1)
        if (9LL * d->C * bn * 2 + 3 * stage + 2048 <= HL_SMEM_BUDGET) { BN = bn; break; }
    if (BN == 0 || (BN < 64 && BN < d->K)) return false;
    const long long tiles = (long long)d->N * cdiv(d->P, HL_TH) * cdiv(d->Q, HL_TW);
    return forced || tiles >= sm_count() * 4;
}

int conv_halo_s2_fprop(const urir_conv_desc* d, const void* x, const void* w_kc, const float* bias, void* y, float* stats,
                       cudaStream_t st) {
    URIR_CHECK_ARG(w_kc != nullptr, "stride-2 halo conv needs the [tap][K][C] weight layout");
    const int kg = d->C, ng = d->K;
    const int BK = 32;                                   // 4 planes per stage: keep the stages small
    HaloMaps maps; HaloParams p; memset(&p, 0, sizeof(p));
    p.N = d->N; p.H = d->P; p.W = d->Q;
    p.tiles_h = cdiv(p.H, HL_TH); p.tiles_w = cdiv(p.W, HL_TW); p.total_tiles = p.tiles_h * p.tiles_w * d->N;
    p.PH = HL_TH + 1; p.PW = HL_TW + 1;
    p.ntaps = 9; p.nchunks = kg / BK;
    p.a_box_bytes = p.PH * p.PW * BK * 2;
    p.plane_bytes = (p.a_box_bytes + 1023) / 1024 * 1024;
    p.nplanes = 4; p.a_stage_bytes = 4 * p.plane_bytes;
    int BN = 0;
    for (int bn = halo_block_n(ng); bn >= 32; bn >>= 1)
        if (9LL * kg * bn * 2 + 3LL * p.a_stage_bytes + 2048 <= HL_SMEM_BUDGET) { BN = bn; break; }
    if (BN == 0) return fail(URIR_ERR_UNSUP, "stride-2 halo conv: weights do not fit in shared memory");
    p.w_bytes = p.ntaps * kg * BN * 2;
    p.stages = (HL_SMEM_BUDGET - 2048 - p.w_bytes) / p.a_stage_bytes;
    halo_fix_stages(p);
    p.o_sn = (long long)d->P * d->Q * d->y_ld; p.o_sh = (long long)d->Q * d->y_ld; p.o_sw = d->y_ld; p.o_off = d->y_coff;
    p.bias = bias; p.stats = stats; p.out = (__nv_bfloat16*)y; p.n_total = ng; p.gate = stats ? next_gate() : nullptr;
    p.bias_mod = ng; p.accumulate = 0;
    for (int g = 0; g < 16; ++g) p.grp_off[g] = 32 * g;
    p.oh0 = 0; p.ow0 = 0;
    const int plane_rows = p.plane_bytes / (BK * 2);
    for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
            const int t = r * 3 + s, plane = (r & 1) * 2 + (s & 1);
            p.tap_row[t] = (short)(plane * plane_rows + (s >> 1) * p.PH + (r >> 1));
            p.wtap[t] = (short)t;
        }
    for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
            const uint64_t dims[4] = {(uint64_t)kg, (uint64_t)d->P, (uint64_t)d->Q, (uint64_t)d->N};
            const uint64_t strides[3] = {2ull * d->W * d->x_ld * 2, 2ull * d->x_ld * 2, (uint64_t)d->H * d->W * d->x_ld * 2};
            const uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.PH, (uint32_t)p.PW, 1};
            const char* base = (const char*)x + ((size_t)d->x_coff + ((size_t)ph * d->W + pw) * d->x_ld) * 2;
            int rc = encode_map(&maps.a[ph * 2 + pw], base, 4, dims, strides, box, BK * 2);
            if (rc) return rc;
        }
    {
        const uint64_t dims[3] = {(uint64_t)kg, (uint64_t)ng, 9};
        const uint64_t strides[2] = {(uint64_t)kg * 2, (uint64_t)kg * ng * 2};
        const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BN, 1};
        int rc = encode_map(&maps.b, w_kc, 3, dims, strides, box, BK * 2);
        if (rc) return rc;
    }
    const int smem = p.w_bytes + p.stages * p.a_stage_bytes + ((HL_MAX_FSETS + 1) * HL_MAX_STAGES + 9) * 8 + 16 + 3 * BN * 4 + 1024;
    const int n_tiles = ng / BN;
    if (BN == 32) return launch_halo<32, 32>(maps, p, n_tiles, smem, st);
    if (BN == 64) return launch_halo<64, 32>(maps, p, n_tiles, smem, st);
    if (BN == 128) return launch_halo<128, 32>(maps, p, n_tiles, smem, st);
    return fail(URIR_ERR_UNSUP, "stride-2 halo conv: no kernel for BLOCK_N=%d", BN);
}

// fp32 HWIO [3][3][C][K] -> bf16 w_up2 [a*2+b][(ph,pw,c)][K], zero where 2a+ph > 2 or 2b+pw > 2
__global__ void weight_prep_up2_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int C, int K) {
    const long long n = 16LL * C * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K); long long r = i / K;
        const int c = (int)(r % C); r /= C;
        const int pw = (int)(r & 1), ph = (int)((r >> 1) & 1), b = (int)((r >> 2) & 1), a = (int)(r >> 3);
        const int rr = 2 * a + ph, ss = 2 * b + pw;
        out[i] = f2bf((rr < 3 && ss < 3) ? w[((size_t)(rr * 3 + ss) * C + c) * K + k] : 0.f);
    }
}
int weight_prep_up2(const float* w, void* out, int C, int K, cudaStream_t st) {
    const long long n = 16LL * C * K;
    int blocks = cdiv(n, 256); if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    weight_prep_up2_kernel<<<blocks, 256, 0, st>>>(w, (__nv_bfloat16*)out, C, K);
    URIR_LAUNCH_OK(0);
    return URIR_OK;
}

}  // namespace urir
