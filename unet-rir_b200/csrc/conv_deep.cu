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
bool stats_sums_only();          // set around a urir_conv2d_dgrad_sums call (capi.cu)

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
    int N, OH, OW;              // images, rows and columns of the GEMM-M grid (stride 1: output = input grid; stride 2: the half-resolution grid)
    int Wp, Hp;                 // padded row width / rows per image (one shared zero column / row)
    int total_tiles;            // M tiles of 128 positions over N * Hp * Wp positions
    int n_chtiles;              // 128-channel tiles of GEMM-N
    int nplanes;                // A-side tensors per channel chunk: 1, or the 4 parity planes of a stride-2 fprop input
    int nclasses;               // output classes: 1, or the 4 output-parity classes of a stride-2 dgrad
    int nchunks, ntaps;
    short job_cta0[17];         // job = class * n_chtiles + channel tile; its CTAs are [job_cta0[job], job_cta0[job + 1])
    short tb[4][4], te[4][4];   // [class][plane]: range of the tap list that plane contributes to that class
    long long o_cls[4];         // output element offset of class c (its parity position)
    int accumulate;             // out += result
    int sums_only;              // statistics: channel sums only (urir_conv2d_dgrad_sums)
    int a_slots, a_stage_bytes, w_stages;
    int tmax;                   // M tiles per round (<= DP_TMAX, < a_slots: the other slots run ahead)
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

struct DeepMaps { CUtensorMap a[4]; CUtensorMap b; };

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
    // this CTA's job (output class, 128-channel tile) and contiguous M-tile range within the job's CTAs
    int job = 0;
    while (job + 1 < p.nclasses * p.n_chtiles && (int)blockIdx.x >= p.job_cta0[job + 1]) ++job;
    const int cls = job / p.n_chtiles, nt = job - cls * p.n_chtiles;
    const int ci = (int)blockIdx.x - p.job_cta0[job], ctas_job = p.job_cta0[job + 1] - p.job_cta0[job];
    const int tile_begin = (int)((long long)p.total_tiles * ci / ctas_job);
    const int tile_end = (int)((long long)p.total_tiles * (ci + 1) / ctas_job);
    const int n_my = tile_end - tile_begin;
    const int tmax = p.tmax;
    const int rounds = (n_my + tmax - 1) / tmax;
    const int S = p.a_slots;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < WS; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
        mbar_init(t_full, 1); mbar_init(t_empty, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
    if (warp == 4 && lane == 0) prefetch_tmap(&maps.a[0]);
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
            for (int kc = 0; kc < p.nchunks; ++kc)
                for (int pl = 0; pl < p.nplanes; ++pl) {
#pragma unroll
                    for (int j = 0; j < DP_TMAX; ++j) {
                        if (j < T) {
                            mbar_wait(a_empty + slot, phase ^ 1);
                            if (lane == 0) mbar_expect_tx(a_full + slot, p.debug == 1 ? 0u : (uint32_t)(nrows[j] * p.Wp) * 128u);
                            __syncwarp();
                            if (lane < nrows[j] && p.debug != 1)             // nrows <= 17 for every admitted width: one box per lane
                                tma_load_4d(&maps.a[pl], a_full + slot, sA + (size_t)slot * p.a_stage_bytes + (size_t)lane * p.Wp * 128,
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
                for (int pl = 0; pl < p.nplanes; ++pl) {
                    const int tb = p.tb[cls][pl], te = p.te[cls][pl];
                    for (int t = tb; t < te; ++t) {
                        const int wt = __shfl_sync(0xffffffffu, (int)p.wtap[t], 0);
                        mbar_wait(w_empty + ws, phase ^ 1);
                        mbar_expect_tx_elect(w_full + ws, p.debug == 1 ? 0u : (uint32_t)DP_W_BYTES);
                        if (p.debug != 1) tma_load_3d_elect(&maps.b, w_full + ws, dst, kc * DP_BK, nt * DP_BN, wt);
                        dst += DP_W_BYTES;
                        if (++ws == WS) { ws = 0; phase ^= 1; dst = sW; }
                    }
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
            bool first = true;                                  // the round's first MMA overwrites the accumulators
            for (int kc = 0; kc < p.nchunks; ++kc)
                for (int pl = 0; pl < p.nplanes; ++pl) {
                    // ring slots of this (chunk, plane)'s T activation stages
                    int sl[DP_TMAX]; uint32_t ph[DP_TMAX], a_base[DP_TMAX];
#pragma unroll
                    for (int j = 0; j < DP_TMAX; ++j) {
                        sl[j] = aslot; ph[j] = aphase; a_base[j] = a_lo0 + (uint32_t)aslot * stage16 + (uint32_t)r0[j];
                        if (j < T) { if (++aslot == S) { aslot = 0; aphase ^= 1; } }
                    }
                    const int tb = p.tb[cls][pl], te = p.te[cls][pl];
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        if (t < tb || t >= te) continue;
                        mbar_wait(w_full + ws, wphase);
                        fence_after_sync();
#pragma unroll
                        for (int j = 0; j < DP_TMAX; ++j) {
                            if (j < T) {
                                if (t == tb) { mbar_wait(a_full + sl[j], ph[j]); fence_after_sync(); }
                                const uint32_t a_lo = a_base[j] + (uint32_t)toff8[t];
#pragma unroll
                                for (int k = 0; k < DP_BK / 16; ++k) {
                                    const uint64_t ad = ((uint64_t)d_hi << 32) | (a_lo + 2 * k);
                                    const uint64_t bd = ((uint64_t)d_hi << 32) | (b_lo + 2 * k);
                                    if (p.debug != 2) umma_bf16_elect(tm0 + j * DP_BN, ad, bd, IDESC, !(first && k == 0));
                                }
                                if (t == te - 1) umma_commit_elect(a_empty + sl[j]);
                            }
                        }
                        first = false;
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
                __nv_bfloat16* orow = p.out + p.o_off + p.o_cls[cls] + (long long)n * p.o_sn + (long long)h * p.o_sh + (long long)c * p.o_sw + nt * DP_BN;
                const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + j * DP_BN;
#pragma unroll 1
                for (int c0 = 0; c0 < (p.debug == 3 ? 0 : DP_BN) && c0 < n_left_tile; c0 += 32) {
                    uint32_t rr[32];
                    tmem_ld32(lane_addr + c0, rr);
                    uint4 old[4];                                       // out += : the previous 32 channels, in flight with the TMEM load
                    if (p.accumulate) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            old[q] = valid ? *reinterpret_cast<const uint4*>(orow + c0 + 8 * q) : make_uint4(0, 0, 0, 0);
                    }
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
                        if (p.accumulate) {
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const uint32_t ow[4] = {old[2 * half + q].x, old[2 * half + q].y, old[2 * half + q].z, old[2 * half + q].w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 o2 = unpack_bf16x2(ow[e]);
                                    v[8 * q + 2 * e] += o2.x; v[8 * q + 2 * e + 1] += o2.y;
                                }
                            }
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
                            const float s2 = p.sums_only ? 0.f : dp_colsum16(q2, lane);
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
        for (int i = threadIdx.x; i < (p.sums_only ? 1 : 2) * DP_BN; i += blockDim.x) {
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

static bool wide_ok_forced_s2(const urir_conv_desc* d, int op) {
    if (d->impl == URIR_IMPL_DEEP) return true;
    static int on = -1;
    if (on < 0) { const char* e = getenv("URIR_DEEP_S2"); on = e ? atoi(e) : 0; }
    return on == 1 || (on == 2 && op == 0);
}

// op 0: fprop (GEMM-K = C, GEMM-N = K), op 1: dgrad (GEMM-K = K, GEMM-N = C)
// wide_ok: also admit widths above URIR_DEEP_WMAX (forced, or no halo-tile kernel takes the layer)
// Stride 2 (3x3, TF SAME on even extents = padding 0 before / 1 after; encoding_block's strided convolution and
// decoding_block's Conv2DTranspose, dl_models/u_net.py:269-276, 296-303) runs on the half-resolution grid:
//   fprop: x[2p + r, 2q + s] lives in parity plane (r & 1, s & 1) at (p + r / 2, q + s / 2): four strided tensor maps,
//          each contributing 4 / 2 / 2 / 1 taps with forward offsets only (the shared zero column / row is the padding);
//   dgrad: dx[2p + a, 2q + b] = sum over the taps with (r & 1, s & 1) = (a, b) of dy[p - r / 2, q - s / 2] W[r, s]: four output
//          classes with 4 / 2 / 2 / 1 taps over the SAME dy tiles, written with doubled pixel strides; CTAs are divided
//          between the (class, channel tile) jobs in proportion to their taps.
bool deep_supported(const urir_conv_desc* d, int op, bool wide_ok) {
    { static int off = -1; if (off < 0) { const char* e = getenv("URIR_NO_DEEP"); off = (e && e[0] == '1') ? 1 : 0; } if (off) return false; }
    if (d->R != 3 || d->S != 3 || d->x_dtype != URIR_BF16 || d->y_dtype != URIR_BF16) return false;
    int GH, GW;                                      // GEMM-M grid
    if (d->stride == 1) {
        if (d->pad_top != 1 || d->pad_left != 1 || d->P != d->H || d->Q != d->W || d->accumulate) return false;
        GH = d->H; GW = d->W;
    } else if (d->stride == 2) {
        // Measured (B = 64, profiles/r02_deep_stride2.txt): correct, but NOT faster than conv_igemm's parity launches -- 23 / 26 us
        // against 23 / 20 us (fprop 36x40 / 18x20), 49 / 54 against 27 / 20 us (dgrad), train step 3.77 against 3.69 ms: a
        // (chunk, plane / class) step holds 1 - 4 taps, so the activation tiles are re-fetched per class and the few-tap
        // CTAs run at the L2 -> SM rate. AUTO therefore leaves the strided layers where they were; URIR_IMPL_DEEP (tests)
        // or URIR_DEEP_S2=1 (all), =2 (fprop only) selects this path.
        if (!wide_ok_forced_s2(d, op)) return false;
        if (d->pad_top != 0 || d->pad_left != 0 || (d->H & 1) || (d->W & 1) || d->P != d->H / 2 || d->Q != d->W / 2) return false;
        if (op == 0 && d->accumulate) return false;
        GH = d->P; GW = d->Q;
    } else return false;
    if (d->act != URIR_ACT_NONE && !(d->act == URIR_ACT_RELU && op == 0)) return false;
    if (d->x_ld % 8 || d->x_coff % 8 || d->y_ld % 8 || d->y_coff % 8) return false;
    const int kg = op == 0 ? d->C : d->K, ng = op == 0 ? d->K : d->C;
    if (kg % DP_BK || ng % DP_BN || kg < 128) return false;
    if (d->stride == 2 && op == 1 && ng / DP_BN > 4) return false;      // job table: 4 classes x <= 4 channel tiles
    const int Wp = GW + 1;
    { int ws, sl; deep_split(Wp, &ws, &sl);
      if (Wp > 256 || sl < 3) return false;                      // TMA box limit; at least two tiles per round + one in flight
      if ((127 + 2 * Wp + 2) / Wp + 2 > 32) return false; }      // the A producer issues one row box per lane
    // at 36x40 and above the halo-tile kernel (resident weights of one N tile) measured faster: 40 / 70 us against 54 / 95
    // for 128 -> 128 / 256 -> 128 at 36x40, B = 64 (profiles/r02_deep_kernel.txt); this kernel takes the levels below
    { static int wmax = -1; if (wmax < 0) { const char* e = getenv("URIR_DEEP_WMAX"); wmax = e ? atoi(e) : 24; }
      if (GW > wmax && !wide_ok) return false; }
    if ((long long)d->N * (GH + 1) * Wp + 2LL * Wp + 256 >= (1LL << 30)) return false;
    return true;
}

// CTAs per job: minimise the largest per-CTA work (tiles x taps) by handing CTAs one at a time to the job whose busiest
// CTA currently has the most
static void deep_assign_ctas(DeepParams& p, const int* job_taps, int njobs, int sms) {
    int cnt[16];
    for (int j = 0; j < njobs; ++j) cnt[j] = 1;
    int used = njobs;
    auto cost = [&](int j) { return (long long)((p.total_tiles + cnt[j] - 1) / cnt[j]) * job_taps[j]; };
    while (used < sms) {
        int best = -1; long long bc = -1;
        for (int j = 0; j < njobs; ++j)
            if (cnt[j] < p.total_tiles && cost(j) > bc) { bc = cost(j); best = j; }
        if (best < 0) break;
        ++cnt[best]; ++used;
    }
    p.job_cta0[0] = 0;
    for (int j = 0; j < njobs; ++j) p.job_cta0[j + 1] = (short)(p.job_cta0[j] + cnt[j]);
}

int conv_deep(const urir_conv_desc* d, int op, const void* a, const void* w, const float* bias, void* out, float* stats,
              cudaStream_t st) {
    URIR_CHECK_ARG(w != nullptr, "deep conv needs the [tap][N][K] weight layout");
    const int kg = op == 0 ? d->C : d->K, ng = op == 0 ? d->K : d->C;
    const int a_ld = op == 0 ? d->x_ld : d->y_ld, a_coff = op == 0 ? d->x_coff : d->y_coff;
    const int o_ld = op == 0 ? d->y_ld : d->x_ld, o_coff = op == 0 ? d->y_coff : d->x_coff;
    const bool s2 = d->stride == 2;
    URIR_CHECK_ARG(!(s2 && op == 1 && stats), "deep conv: no statistics epilogue on the stride-2 input gradient");
    const int GH = s2 ? d->P : d->H, GW = s2 ? d->Q : d->W;          // GEMM-M grid
    DeepMaps maps; DeepParams p; memset(&p, 0, sizeof(p));
    p.N = d->N; p.OH = GH; p.OW = GW; p.Wp = GW + 1; p.Hp = GH + 1;
    p.total_tiles = (int)(((long long)d->N * p.Hp * p.Wp + 127) / 128);
    p.n_chtiles = ng / DP_BN;
    p.nplanes = (s2 && op == 0) ? 4 : 1;
    p.nclasses = (s2 && op == 1) ? 4 : 1;
    p.nchunks = kg / DP_BK; p.ntaps = 9;
    p.accumulate = d->accumulate; p.sums_only = stats_sums_only() ? 1 : 0;
    // tap list, grouped by (class, plane)
    int job_taps[16];
    {
        int nt = 0;
        const int groups = s2 ? 4 : 1;
        for (int gidx = 0; gidx < groups; ++gidx) {
            const int gr = gidx >> 1, gs = gidx & 1;                 // (row, column) parity of the group
            const int cls = p.nclasses > 1 ? gidx : 0, pl = p.nplanes > 1 ? gidx : 0;
            p.tb[cls][pl] = (short)nt;
            for (int r = 0; r < 3; ++r)
                for (int s = 0; s < 3; ++s) {
                    if (s2 && ((r & 1) != gr || (s & 1) != gs)) continue;
                    int dh, dw;
                    if (!s2) { dh = op == 0 ? r - 1 : 1 - r; dw = op == 0 ? s - 1 : 1 - s; }
                    else if (op == 0) { dh = r >> 1; dw = s >> 1; }
                    else { dh = -(r >> 1); dw = -(s >> 1); }
                    p.tap_off[nt] = (short)(dh * p.Wp + dw);
                    p.wtap[nt] = (short)(r * 3 + s);
                    ++nt;
                }
            p.te[cls][pl] = (short)nt;
        }
        for (int c = 0; c < p.nclasses; ++c) {
            int taps = 0;
            for (int pl = 0; pl < p.nplanes; ++pl) taps += p.te[c][pl] - p.tb[c][pl];
            for (int t = 0; t < p.n_chtiles; ++t) job_taps[c * p.n_chtiles + t] = taps;
        }
    }
    const int njobs = p.nclasses * p.n_chtiles;
    if (njobs > 16) return fail(URIR_ERR_UNSUP, "deep conv: %d (class, channel tile) jobs", njobs);
    const int sms = sm_count();
    deep_assign_ctas(p, job_taps, njobs, sms > njobs ? sms : njobs);
    p.a_stage_bytes = deep_a_stage_bytes(p.Wp);
    deep_split(p.Wp, &p.w_stages, &p.a_slots);
    { static int ov = -1; if (ov < 0) { const char* e = getenv("URIR_DEEP_WSTAGES"); ov = e ? atoi(e) : 0; }
      if (ov >= 2 && ov <= DP_MAX_WSTAGES) {          // experiment knob: fixed ring depth, activation slots from what is left
          p.w_stages = ov;
          int sl = (DP_SMEM_BUDGET - deep_fixed_bytes() - ov * DP_W_BYTES) / p.a_stage_bytes;
          p.a_slots = sl > DP_MAX_ASLOTS ? DP_MAX_ASLOTS : sl;
          if (p.a_slots < 2) return fail(URIR_ERR_UNSUP, "deep conv: URIR_DEEP_WSTAGES=%d leaves no room for activations", ov);
      } }
    p.tmax = p.a_slots - 1 < DP_TMAX ? p.a_slots - 1 : DP_TMAX;
    if (s2) {
        // a (chunk, plane) / (chunk, class) step of a strided layer holds 1 - 4 taps instead of 9: the activation ring has to
        // run further ahead of the MMAs (a row-box batch takes ~2100 cycles to land), so fewer weight stages, fewer tiles per round
        static int ws2 = -1, t2 = -1;
        if (ws2 < 0) { const char* e = getenv("URIR_DEEP_S2_WSTAGES"); ws2 = e ? atoi(e) : 4; }
        if (t2 < 0) { const char* e = getenv("URIR_DEEP_S2_T"); t2 = e ? atoi(e) : 2; }
        if (ws2 >= 2 && ws2 <= DP_MAX_WSTAGES) {
            int sl = (DP_SMEM_BUDGET - deep_fixed_bytes() - ws2 * DP_W_BYTES) / p.a_stage_bytes;
            if (sl > DP_MAX_ASLOTS) sl = DP_MAX_ASLOTS;
            if (sl >= 2) { p.w_stages = ws2; p.a_slots = sl; }
        }
        p.tmax = t2 < p.a_slots ? t2 : p.a_slots - 1;
        if (p.tmax > DP_TMAX) p.tmax = DP_TMAX;
        if (p.tmax < 1) p.tmax = 1;
    }
    // accumulators: as many 128-column blocks as tiles run together, rounded to a power of two
    { int per = 1;
      for (int j = 0; j < njobs; ++j) { const int c = p.job_cta0[j + 1] - p.job_cta0[j]; const int q = (p.total_tiles + c - 1) / c; if (q > per) per = q; }
      int t = p.tmax; if (per < t) t = per;
      const int cols = t * DP_BN; p.tmem_cols = cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }
    if (s2 && op == 1) {        // class (a, b) writes dx[2p + a, 2q + b]
        p.o_sn = (long long)d->H * d->W * o_ld; p.o_sh = 2LL * d->W * o_ld; p.o_sw = 2LL * o_ld; p.o_off = o_coff;
        for (int c = 0; c < 4; ++c) p.o_cls[c] = ((long long)(c >> 1) * d->W + (c & 1)) * o_ld;
    } else {
        p.o_sn = (long long)GH * GW * o_ld; p.o_sh = (long long)GW * o_ld; p.o_sw = o_ld; p.o_off = o_coff;
    }
    p.bias = bias; p.stats = stats; p.out = (__nv_bfloat16*)out; p.n_total = ng; p.relu = d->act == URIR_ACT_RELU;
    p.gate = stats ? next_gate() : nullptr;
    { const char* e = getenv("URIR_DEEP_DEBUG"); p.debug = e ? atoi(e) : 0; }
    for (int pl = 0; pl < p.nplanes; ++pl) {
        // activation: dims (C, W, H, N); one box = one padded row: Wp columns from column 0 (the last one out of bounds = 0).
        // Stride-2 fprop: parity plane (pl / 2, pl % 2) of x = every second row / column from that origin.
        const int sub = p.nplanes > 1 ? 2 : 1;
        const int AH = op == 0 ? d->H : d->P, AW = op == 0 ? d->W : d->Q;             // extent of the A-side tensor
        const uint64_t dims[4] = {(uint64_t)kg, (uint64_t)(AW / sub), (uint64_t)(AH / sub), (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)sub * a_ld * 2, (uint64_t)sub * AW * a_ld * 2, (uint64_t)AH * AW * a_ld * 2};
        const uint32_t box[4] = {(uint32_t)DP_BK, (uint32_t)p.Wp, 1, 1};
        const size_t origin = p.nplanes > 1 ? ((size_t)(pl >> 1) * AW + (pl & 1)) * a_ld : 0;
        int rc = encode_map(&maps.a[pl], (const char*)a + ((size_t)a_coff + origin) * 2, 4, dims, strides, box, 128);
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
    dim3 grid(p.job_cta0[njobs]);
    URIR_CUDA_OK(launch_pdl(conv_deep_kernel, grid, dim3(DP_THREADS), smem, st, maps, p));
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

}  // namespace urir
