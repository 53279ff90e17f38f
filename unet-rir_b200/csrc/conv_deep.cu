// Persistent padded-sequence implicit GEMM for the DEEP stride-1 3x3 convolutions (<= 36x40, >= 128 channels:
// convolutional_block_1 at levels 3-5 and the decoder's fuse convolutions, dl_models/u_net.py:222-240, 363-371, and
// their input gradients under tape.gradient, amp_phase_trainer.py:138).
//
// conv_igemm.cu runs these layers with one 128-pixel tile per short-lived CTA and every tap as a GEMM-K step over its
// own re-fetched operands: 32 KB of A + B cross L2 -> SM per 256 cycles of MMA = 128 B/clk/SM against the ~43 B/clk/SM
// the L2 delivers with every SM pulling (ncu r1: tensor pipe 26 % busy). conv_halo.cu's cure (fetch the activation tile
// once with its halo, keep the weights resident) does not fit here: 8x16 tiles waste half an 18x20 image and the
// weights of >= 128 channels do not fit in shared memory. This kernel:
//   * flattens the batch into ONE zero-separated sequence of positions: every image row is followed by ONE zero column
//     and every image by ONE zero row (TMA out-of-bounds fill supplies them), so the 3x3 neighbourhood of position p is
//     p + dh * Wp + dw (Wp = W + 1) for EVERY p -- the zero column / row to the right / below doubles as the left / top
//     halo of the next row / image. An M tile is 128 consecutive positions (95 / 90 / 82 % real pixels at 36x40 / 18x20 /
//     9x10), its A operand per 64-channel chunk the padded rows it touches plus one row either side, brought by one
//     {64 ch, Wp, 1 row} TMA box per padded row, packed back to back in shared memory (rows start at arbitrary multiples
//     of 128 bytes: both TMA and tcgen05.mma anchor the 128-byte swizzle to absolute shared-memory address bits,
//     profiles/r02_tma_row_probe.txt, r01_umma_halo_probe.txt). The nine taps are nine descriptor start offsets.
//   * gives every CTA (one per SM, persistent) a contiguous range of M tiles of ONE 128-channel N tile and runs up to four
//     of them together (four 128-column TMEM accumulators): loop order (chunk, tap, tile), so each 16 KB weight tile
//     streamed from L2 feeds up to four M tiles. L2 -> SM traffic per 64-channel chunk: T * ~27 KB of A + 144 KB of B
//     per T * 2304 MMA cycles = 37 B/clk/SM at T = 3 (was 128).
// Warp roles (352 threads): warps 0-3 and 8-10... see below: warps 0-3 = epilogue group 0, warp 4 = A producer, warp 5 =
// weight producer, warp 6 = MMA issuer, warp 7 idle, warps 8-11 = epilogue group 1 (tiles alternate between the groups).
// Epilogue as in conv_igemm.cu: +bias, optional ReLU (inference, BN folded), per-channel sum / sum of squares (BatchNorm
// statistics or bias gradients; per-warp shared-memory slots, no floating-point atomics inside the CTA), bf16, 32-byte
// stores into the (possibly channel-sliced) NHWC destination.
#include <stdlib.h>
#include <atomic>
#include "urir_common.cuh"
#include "urir_tc.cuh"

namespace urir {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes);

constexpr int DP_BN = 128;               // GEMM-N tile (one tcgen05.mma N)
constexpr int DP_BK = 64;                // GEMM-K chunk: 64 bf16 = one 128-byte swizzled row
constexpr int DP_TMAX = 4;               // M tiles per round (4 x 128 TMEM columns)
constexpr int DP_MAX_ASLOTS = 8;
constexpr int DP_MAX_WSTAGES = 8;         // weight ring depth is a run-time parameter: a TMA box takes ~2100 cycles to land
                                          // (profiles/r01_tma_box_throughput.txt), so a 4-deep ring of 16 KB tiles delivers one
                                          // tile per ~525 cycles = less than ONE M tile per tap consumes (256 cycles of MMA)
constexpr int DP_W_BYTES = DP_BN * DP_BK * 2;           // 16 KB weight tile
constexpr int DP_SMEM_BUDGET = 224 * 1024;
constexpr int DP_THREADS = 384;

struct DeepParams {
    int N, OH, OW;              // images, rows and columns of the (output = input) grid
    int Wp, Hp;                 // padded row width / rows per image (one shared zero column / row)
    int total_tiles;            // M tiles of 128 positions over N * Hp * Wp positions
    int n_ntiles, ctas_per_nt;  // 128-channel N tiles; CTAs that share the M range of one N tile
    int nchunks, ntaps;
    int a_slots, a_stage_bytes, w_stages;
    int tmem_cols;
    long long o_sn, o_sh, o_sw, o_off;   // output element strides / offset (channel slice of an NHWC buffer)
    const float* bias;
    float* stats;
    __nv_bfloat16* out;
    int n_total;
    int relu;
    unsigned int* gate;         // deterministic mode (urir_common.cuh)
    int debug;                  // URIR_DEEP_DEBUG timing experiments: 1 no TMA loads, 2 no MMAs, 3 no epilogue math / stores
    short tap_off[9];           // dh * Wp + dw of tap t (rows of the padded sequence)
    short wtap[9];              // weight tap index of tap t
};

struct DeepMaps { CUtensorMap a; CUtensorMap b; };

// first padded row a tile needs, how many, and the row (within the loaded block) of the tile's first position
struct DeepTile { int g_lo, nrows, r0; };
__device__ __forceinline__ DeepTile deep_tile(const DeepParams& p, int tile) {
    const int p0 = p.Wp + tile * 128;                    // positions start after the leading zero row
    const int lo = p0 - p.Wp - 1, hi = p0 + 127 + p.Wp + 1;
    DeepTile t;
    t.g_lo = lo / p.Wp;                                  // lo >= -1 + ... : p0 >= Wp, so lo >= -1; -1 / Wp == 0 in C: handled below
    if (lo < 0) t.g_lo = -1;
    t.nrows = hi / p.Wp - t.g_lo + 1;
    t.r0 = p0 - t.g_lo * p.Wp;
    return t;
}

__device__ __forceinline__ float dp_colsum16(float (&v)[16], int lane) {
#define URIR_HALVE(OFF, CNT, BIT) { const bool up = lane & BIT; _Pragma("unroll") for (int j = 0; j < CNT; ++j) { \
        const float send = up ? v[j] : v[j + CNT]; const float keep = up ? v[j + CNT] : v[j]; \
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF); } }
    URIR_HALVE(16, 8, 16)
    URIR_HALVE(8, 4, 8)
    URIR_HALVE(4, 2, 4)
    URIR_HALVE(2, 1, 2)
#undef URIR_HALVE
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
__device__ __forceinline__ int dp_col_of_lane(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

__global__ void __launch_bounds__(DP_THREADS, 1)
conv_deep_kernel(const __grid_constant__ DeepMaps maps, const __grid_constant__ DeepParams p) {
    constexpr uint32_t IDESC = make_idesc_bf16(128, DP_BN, 0, 0);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;                                          // w_stages x 16 KB
    uint8_t* sA = smem + p.w_stages * DP_W_BYTES;                // a_slots x a_stage_bytes
    const int WS = p.w_stages;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(sA + (size_t)p.a_slots * p.a_stage_bytes);
    uint64_t* a_empty = a_full + DP_MAX_ASLOTS;
    uint64_t* w_full = a_empty + DP_MAX_ASLOTS;
    uint64_t* w_empty = w_full + DP_MAX_WSTAGES;
    uint64_t* t_full = w_empty + DP_MAX_WSTAGES;
    uint64_t* t_empty = t_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 1);
    float* sstats = reinterpret_cast<float*>(tmem_slot + 2);     // [8 epilogue warps][2 * DP_BN]
    float* sbias = sstats + 8 * 2 * DP_BN;                       // [DP_BN]

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    // this CTA's N tile and contiguous M-tile range
    const int nt = blockIdx.x % p.n_ntiles, ci = blockIdx.x / p.n_ntiles;
    const int tile_begin = (int)((long long)p.total_tiles * ci / p.ctas_per_nt);
    const int tile_end = (int)((long long)p.total_tiles * (ci + 1) / p.ctas_per_nt);
    const int n_my = tile_end - tile_begin;
    const int tmax = (p.a_slots - 1) < DP_TMAX ? (p.a_slots - 1) : DP_TMAX;
    const int rounds = (n_my + tmax - 1) / tmax;
    const int S = p.a_slots;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < WS; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
        mbar_init(t_full, 1); mbar_init(t_empty, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    if (warp == 4 && lane == 0) prefetch_tmap(&maps.a);
    if (warp == 5 && lane == 0) prefetch_tmap(&maps.b);
    for (int i = threadIdx.x; i < 8 * 2 * DP_BN; i += blockDim.x) sstats[i] = 0.f;
    for (int i = threadIdx.x; i < DP_BN; i += blockDim.x)
        sbias[i] = (p.bias && nt * DP_BN + i < p.n_total) ? p.bias[nt * DP_BN + i] : 0.f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    // (no integer divisions inside the pipeline loops: ring slots and phases advance incrementally, tile geometry is
    // computed once per round -- a runtime division is a ~100-cycle dependent chain, and the first version of this
    // kernel spent ~1000 cycles per tap in them on the single issuing thread: profiles/r02_deep_kernel.txt)
    if (warp == 4) {
        // ===================== A producer: one {64 ch, Wp, 1 row} box per padded row of the tile =====================
        int slot = 0; uint32_t phase = 0;
        for (int r = 0; r < rounds; ++r) {
            const int t0 = tile_begin + (int)((long long)n_my * r / rounds), t1 = tile_begin + (int)((long long)n_my * (r + 1) / rounds);
            const int T = t1 - t0;
            // this lane's row of every tile of the round: image / row coordinates, box destination, byte count
            int rn[DP_TMAX], rh[DP_TMAX], nrows[DP_TMAX];
#pragma unroll
            for (int j = 0; j < DP_TMAX; ++j) {
                rn[j] = -1; rh[j] = 0; nrows[j] = 0;
                if (j < T) {
                    const DeepTile dt = deep_tile(p, t0 + j);
                    nrows[j] = dt.nrows;
                    const int g = dt.g_lo + lane - 1;                        // padded row relative to the first image's row 0
                    if (g >= 0) { rn[j] = g / p.Hp; rh[j] = g - rn[j] * p.Hp; }   // g < 0: leading zero rows (n = -1: zero fill);
                }                                                            // h == OH: the image's trailing zero row
            }
            for (int kc = 0; kc < p.nchunks; ++kc) {
#pragma unroll
                for (int j = 0; j < DP_TMAX; ++j) {
                    if (j < T) {
                        mbar_wait(a_empty + slot, phase ^ 1);
                        if (lane == 0) mbar_expect_tx(a_full + slot, p.debug == 1 ? 0u : (uint32_t)(nrows[j] * p.Wp) * 128u);
                        __syncwarp();
                        if (lane < nrows[j] && p.debug != 1)                 // nrows <= 17 for every admitted width: one box per lane
                            tma_load_4d(&maps.a, a_full + slot, sA + (size_t)slot * p.a_stage_bytes + (size_t)lane * p.Wp * 128,
                                        kc * DP_BK, 0, rh[j], rn[j]);
                        if (++slot == S) { slot = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===================== weight producer: one 16 KB {64, 128} tile per (chunk, tap) =====================
        int ws = 0; uint32_t phase = 0;
        uint8_t* dst = sW;
        for (int r = 0; r < rounds; ++r)
            for (int kc = 0; kc < p.nchunks; ++kc)
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    mbar_wait(w_empty + ws, phase ^ 1);
                    mbar_expect_tx_elect(w_full + ws, p.debug == 1 ? 0u : (uint32_t)DP_W_BYTES);
                    if (p.debug != 1) tma_load_3d_elect(&maps.b, w_full + ws, dst, kc * DP_BK, nt * DP_BN, t);
                    dst += DP_W_BYTES;
                    if (++ws == WS) { ws = 0; phase ^= 1; dst = sW; }
                }
    } else if (warp == 6) {
        // ===================== MMA issuer =====================
        const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t d_hi = ((8 * 128) >> 4) | (1u << 14) | (SWZ_128B << 29);      // SBO = 8 rows | sm100 version | swizzle
        const uint32_t w_lo0 = smem_u32(sW) >> 4, a_lo0 = smem_u32(sA) >> 4;
        const uint32_t stage16 = (uint32_t)p.a_stage_bytes >> 4;
        int toff8[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) toff8[t] = __shfl_sync(0xffffffffu, (int)p.tap_off[t], 0) * 8;     // 128-byte rows in 16-byte units
        int ws = 0; uint32_t wphase = 0, b_lo = w_lo0;
        int aslot = 0; uint32_t aphase = 0;
        for (int r = 0; r < rounds; ++r) {
            const int t0 = tile_begin + (int)((long long)n_my * r / rounds), t1 = tile_begin + (int)((long long)n_my * (r + 1) / rounds);
            const int T = t1 - t0;
            int r0[DP_TMAX];
#pragma unroll
            for (int j = 0; j < DP_TMAX; ++j) r0[j] = j < T ? deep_tile(p, t0 + j).r0 * 8 : 0;
            mbar_wait(t_empty, (r & 1) ^ 1);                   // the epilogue has drained the previous round's accumulators
            fence_after_sync();
            for (int kc = 0; kc < p.nchunks; ++kc) {
                // ring slots of this chunk's T activation stages
                int sl[DP_TMAX]; uint32_t ph[DP_TMAX], a_base[DP_TMAX];
#pragma unroll
                for (int j = 0; j < DP_TMAX; ++j) {
                    sl[j] = aslot; ph[j] = aphase; a_base[j] = a_lo0 + (uint32_t)aslot * stage16 + (uint32_t)r0[j];
                    if (j < T) { if (++aslot == S) { aslot = 0; aphase ^= 1; } }
                }
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    mbar_wait(w_full + ws, wphase);
                    fence_after_sync();
#pragma unroll
                    for (int j = 0; j < DP_TMAX; ++j) {
                        if (j < T) {
                            if (t == 0) { mbar_wait(a_full + sl[j], ph[j]); fence_after_sync(); }
                            const uint32_t a_lo = a_base[j] + (uint32_t)toff8[t];
#pragma unroll
                            for (int k = 0; k < DP_BK / 16; ++k) {
                                const uint64_t ad = ((uint64_t)d_hi << 32) | (a_lo + 2 * k);
                                const uint64_t bd = ((uint64_t)d_hi << 32) | (b_lo + 2 * k);
                                if (p.debug != 2) umma_bf16_elect(tm0 + j * DP_BN, ad, bd, IDESC, (kc | t | k) != 0);
                            }
                            if (t == 8) umma_commit_elect(a_empty + sl[j]);
                        }
                    }
                    umma_commit_elect(w_empty + ws);
                    b_lo += DP_W_BYTES >> 4;
                    if (++ws == WS) { ws = 0; wphase ^= 1; b_lo = w_lo0; }
                }
            }
            umma_commit_elect(t_full);
            __syncwarp();
        }
    } else if (warp < 4 || (warp >= 8 && warp < 12)) {
        // ===================== epilogue: group 0 = warps 0-3, group 1 = warps 8-11; tile j of a round -> group j & 1 ====
        const int eg = warp >= 8 ? 1 : 0, quarter = warp & 3, ew = eg * 4 + quarter;
        float* my_stats = sstats + ew * 2 * DP_BN;
        const bool want_stats = p.stats != nullptr;
        const int n_left_tile = p.n_total - nt * DP_BN;
        for (int r = 0; r < rounds; ++r) {
            const int t0 = tile_begin + (int)((long long)n_my * r / rounds), t1 = tile_begin + (int)((long long)n_my * (r + 1) / rounds);
            mbar_wait(t_full, r & 1);
            fence_after_sync();
            for (int tile = t0 + eg; tile < t1; tile += 2) {
                const int j = tile - t0;
                const int row = quarter * 32 + lane;
                const int pos = tile * 128 + row;                       // position past the leading zero row
                const int g = pos / p.Wp, c = pos - g * p.Wp;
                const int n = g / p.Hp, h = g - n * p.Hp;
                const bool valid = c < p.OW && h < p.OH && n < p.N;
                __nv_bfloat16* orow = p.out + p.o_off + (long long)n * p.o_sn + (long long)h * p.o_sh + (long long)c * p.o_sw + nt * DP_BN;
                const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + j * DP_BN;
#pragma unroll 1
                for (int c0 = 0; c0 < (p.debug == 3 ? 0 : DP_BN) && c0 < n_left_tile; c0 += 32) {
                    uint32_t rr[32];
                    tmem_ld32(lane_addr + c0, rr);
                    tmem_ld_wait();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            v[q] = __uint_as_float(rr[half * 16 + q]) + sbias[c0 + half * 16 + q];
                            if (p.relu) v[q] = fmaxf(v[q], 0.f);
                            if (!valid) v[q] = 0.f;
                        }
                        if (valid) {
                            __nv_bfloat16* o = orow + c0 + half * 16;
                            *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                            *reinterpret_cast<uint4*>(o + 8) = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                        }
                        if (want_stats) {
                            float q2[16];
#pragma unroll
                            for (int q = 0; q < 16; ++q) q2[q] = v[q] * v[q];
                            const float s1 = dp_colsum16(v, lane);
                            const float s2 = dp_colsum16(q2, lane);
                            if ((lane & 1) == 0) {          // this lane owns (warp, column): plain accumulation, fixed order
                                const int col = c0 + half * 16 + dp_col_of_lane(lane);
                                my_stats[col] += s1;
                                my_stats[DP_BN + col] += s2;
                            }
                        }
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty);
        }
    }
    __syncthreads();
    if (p.stats) {
        gate_enter(p.gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
        for (int i = threadIdx.x; i < 2 * DP_BN; i += blockDim.x) {
            const int which = i / DP_BN, col = i % DP_BN;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += sstats[w * 2 * DP_BN + i];
            if (n_my > 0 && nt * DP_BN + col < p.n_total) atomicAdd(p.stats + which * p.n_total + nt * DP_BN + col, v);
        }
        gate_leave(p.gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
    }
    if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int deep_a_stage_bytes(int Wp) {
    const int rows = (127 + 2 * Wp + 2) / Wp + 2;        // padded rows a 128-position tile plus one row either side can touch
    return (rows * Wp * 128 + 1023) / 1024 * 1024;
}
static int deep_fixed_bytes() {
    return (2 * DP_MAX_ASLOTS + 2 * DP_MAX_WSTAGES + 2) * 8 + 16 + (8 * 2 * DP_BN + DP_BN) * 4 + 1024;
}
// shared-memory split: the deepest weight ring (8 .. 4 stages) that still leaves four activation slots (three M tiles per
// round + one in flight); below that, whatever fits with four weight stages
static void deep_split(int Wp, int* w_stages, int* a_slots) {
    const int stage = deep_a_stage_bytes(Wp);
    for (int ws = DP_MAX_WSTAGES; ws >= 4; --ws) {
        int s = (DP_SMEM_BUDGET - deep_fixed_bytes() - ws * DP_W_BYTES) / stage;
        if (s > DP_MAX_ASLOTS) s = DP_MAX_ASLOTS;
        if (s >= 4 || ws == 4) { *w_stages = ws; *a_slots = s; return; }
    }
}

// op 0: fprop (GEMM-K = C, GEMM-N = K), op 1: dgrad (GEMM-K = K, GEMM-N = C)
// wide_ok: also admit widths above URIR_DEEP_WMAX (forced, or no halo-tile kernel takes the layer)
bool deep_supported(const urir_conv_desc* d, int op, bool wide_ok) {
    { static int off = -1; if (off < 0) { const char* e = getenv("URIR_NO_DEEP"); off = (e && e[0] == '1') ? 1 : 0; } if (off) return false; }
    if (d->stride != 1 || d->R != 3 || d->S != 3 || d->pad_top != 1 || d->pad_left != 1 || d->P != d->H || d->Q != d->W) return false;
    if (d->x_dtype != URIR_BF16 || d->y_dtype != URIR_BF16 || d->accumulate) return false;
    if (d->act != URIR_ACT_NONE && !(d->act == URIR_ACT_RELU && op == 0)) return false;
    if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return false;
    const int kg = op == 0 ? d->C : d->K, ng = op == 0 ? d->K : d->C;
    if (kg % DP_BK || ng % DP_BN || kg < 128) return false;
    const int Wp = d->W + 1;
    { int ws, sl; deep_split(Wp, &ws, &sl);
      if (Wp > 256 || sl < 3) return false;                      // TMA box limit; at least two tiles per round + one in flight
      if ((127 + 2 * Wp + 2) / Wp + 2 > 32) return false; }      // the A producer issues one row box per lane
    // at 36x40 and above the halo-tile kernel (resident weights of one N tile) measured faster: 40 / 70 us against 54 / 95
    // for 128 -> 128 / 256 -> 128 at 36x40, B = 64 (profiles/r02_deep_kernel.txt); this kernel takes the levels below
    { static int wmax = -1; if (wmax < 0) { const char* e = getenv("URIR_DEEP_WMAX"); wmax = e ? atoi(e) : 24; }
      if (d->W > wmax && !wide_ok) return false; }
    if ((long long)d->N * (d->H + 1) * Wp + 2LL * Wp + 256 >= (1LL << 30)) return false;
    return true;
}

int conv_deep(const urir_conv_desc* d, int op, const void* a, const void* w, const float* bias, void* out, float* stats,
              cudaStream_t st) {
    URIR_CHECK_ARG(w != nullptr, "deep conv needs the [tap][N][K] weight layout");
    const int kg = op == 0 ? d->C : d->K, ng = op == 0 ? d->K : d->C;
    const int a_ld = op == 0 ? d->x_ld : d->y_ld, a_coff = op == 0 ? d->x_coff : d->y_coff;
    const int o_ld = op == 0 ? d->y_ld : d->x_ld, o_coff = op == 0 ? d->y_coff : d->x_coff;
    DeepMaps maps; DeepParams p; memset(&p, 0, sizeof(p));
    p.N = d->N; p.OH = d->H; p.OW = d->W; p.Wp = d->W + 1; p.Hp = d->H + 1;
    p.total_tiles = (int)(((long long)d->N * p.Hp * p.Wp + 127) / 128);
    p.n_ntiles = ng / DP_BN;
    const int sms = sm_count();
    p.ctas_per_nt = sms / p.n_ntiles; if (p.ctas_per_nt < 1) p.ctas_per_nt = 1;
    if (p.ctas_per_nt > p.total_tiles) p.ctas_per_nt = p.total_tiles;
    p.nchunks = kg / DP_BK; p.ntaps = 9;
    p.a_stage_bytes = deep_a_stage_bytes(p.Wp);
    deep_split(p.Wp, &p.w_stages, &p.a_slots);
    { static int ov = -1; if (ov < 0) { const char* e = getenv("URIR_DEEP_WSTAGES"); ov = e ? atoi(e) : 0; }
      if (ov >= 2 && ov <= DP_MAX_WSTAGES) {          // experiment knob: fixed ring depth, activation slots from what is left
          p.w_stages = ov;
          int sl = (DP_SMEM_BUDGET - deep_fixed_bytes() - ov * DP_W_BYTES) / p.a_stage_bytes;
          p.a_slots = sl > DP_MAX_ASLOTS ? DP_MAX_ASLOTS : sl;
          if (p.a_slots < 2) return fail(URIR_ERR_UNSUP, "deep conv: URIR_DEEP_WSTAGES=%d leaves no room for activations", ov);
      } }
    // accumulators: as many 128-column blocks as tiles run together, rounded to a power of two
    { const int per = (p.total_tiles + p.ctas_per_nt - 1) / p.ctas_per_nt;
      int t = p.a_slots - 1 < DP_TMAX ? p.a_slots - 1 : DP_TMAX; if (per < t) t = per;
      const int cols = t * DP_BN; p.tmem_cols = cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }
    p.o_sn = (long long)d->H * d->W * o_ld; p.o_sh = (long long)d->W * o_ld; p.o_sw = o_ld; p.o_off = o_coff;
    p.bias = bias; p.stats = stats; p.out = (__nv_bfloat16*)out; p.n_total = ng; p.relu = d->act == URIR_ACT_RELU;
    p.gate = stats ? next_gate() : nullptr;
    { const char* e = getenv("URIR_DEEP_DEBUG"); p.debug = e ? atoi(e) : 0; }
    for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
            const int t = r * 3 + s;
            const int dh = op == 0 ? r - 1 : 1 - r, dw = op == 0 ? s - 1 : 1 - s;
            p.tap_off[t] = (short)(dh * p.Wp + dw);
            p.wtap[t] = (short)t;
        }
    {   // activation: dims (C, W, H, N); one box = one padded row: Wp columns from column 0 (the last one out of bounds = 0)
        const uint64_t dims[4] = {(uint64_t)kg, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)a_ld * 2, (uint64_t)d->W * a_ld * 2, (uint64_t)d->H * d->W * a_ld * 2};
        const uint32_t box[4] = {(uint32_t)DP_BK, (uint32_t)p.Wp, 1, 1};
        int rc = encode_map(&maps.a, (const char*)a + (size_t)a_coff * 2, 4, dims, strides, box, 128);
        if (rc) return rc;
    }
    {
        const uint64_t dims[3] = {(uint64_t)kg, (uint64_t)ng, 9};
        const uint64_t strides[2] = {(uint64_t)kg * 2, (uint64_t)kg * ng * 2};
        const uint32_t box[3] = {(uint32_t)DP_BK, (uint32_t)DP_BN, 1};
        int rc = encode_map(&maps.b, w, 3, dims, strides, box, 128);
        if (rc) return rc;
    }
    const int smem = p.w_stages * DP_W_BYTES + p.a_slots * p.a_stage_bytes + deep_fixed_bytes();
    static std::atomic<bool> attr_set{false};
    if (!attr_set) {
        URIR_CUDA_OK(cudaFuncSetAttribute(conv_deep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_BUDGET + 2048));
        attr_set = true;
    }
    dim3 grid(p.ctas_per_nt * p.n_ntiles);
    URIR_CUDA_OK(launch_pdl(conv_deep_kernel, grid, dim3(DP_THREADS), smem, st, maps, p));
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

}  // namespace urir
