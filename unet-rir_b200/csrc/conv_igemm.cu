// tcgen05 implicit-GEMM convolution for sm_100a: one kernel family serves
//   * Conv2D fprop, stride 1 and 2        (dl_models/u_net.py:269-276, 366)
//   * Conv2D dgrad, stride 1 and 2        (tape.gradient, amp_phase_trainer.py:138)
//   * Conv2DTranspose fprop (= dgrad s2)  (dl_models/u_net.py:297-304) and its dgrad (= fprop s2)
//
// GEMM view: D[pixels (M=128 per CTA), channels_out (N)] = sum over (tap, channel_in chunk)
//            A[pixels, chunk] * B[channels_out, chunk]^T.
//   A: the NHWC activation itself. For tap (dh,dw) a 4-D TMA box {BLOCK_K ch, bw, bh, bn pixels}
//      is fetched at the shifted coordinate; out-of-range rows/cols are zero-filled by TMA, which
//      IS the TF "SAME" padding. Stride-2 fprop reads one of four parity views of x (tensor maps
//      with doubled W/H strides), so no strided gather is ever needed. Stride-2 dgrad decomposes
//      into four output-parity classes (blockIdx.z), each a dense stride-1 problem over a tap subset.
//   B: bf16 weights [tap][N][Kgemm] (K-major), one 3-D TMA box {BLOCK_K, BLOCK_N, 1} per step.
//   Both land in 128B (or 64B when Cin = 32) swizzled K-major tiles that tcgen05.mma consumes
//   directly from shared memory; the fp32 accumulator lives in TMEM.
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA
// producer, warp 5 MMA issuer. smem ring of STAGES {A,B} tiles with full/empty mbarriers.
// Epilogue: tcgen05.ld -> +bias -> per-channel sum / sum-of-squares (BatchNorm batch statistics
// or bias gradients; warp transpose-reduce, smem atomics, one global atomic per channel per CTA)
// -> bf16 -> 32-byte stores straight into the (possibly channel-sliced = concat, possibly
// parity-strided) NHWC destination.
#include <stdlib.h>
#include <atomic>
#include "urir_common.cuh"
#include "urir_tc.cuh"

namespace urir {

using namespace tc;

struct IgemmTap { short dh, dw; short map; short wtap; };

struct IgemmParams {
    int bw, bh, bn;                 // pixel box; bw*bh*bn <= 128 rows of the M tile
    int tiles_w, tiles_h, tiles_n;  // tiles over the (largest) logical output grid
    int NB;                         // batch
    int n_classes;                  // 1, or 4 output-parity classes (stride-2 dgrad)
    int cls_OW[4], cls_OH[4];       // logical output extent per class
    long long cls_off[4];           // element offset of the class origin in the output buffer
    int tap_begin[5];               // taps of class c: [tap_begin[c], tap_begin[c+1])
    long long o_sn, o_sh, o_sw;     // output element strides (batch, logical row, logical col)
    int kchunks;                    // Kgemm / BLOCK_K
    int accumulate;
    const float* bias;              // [Ngemm] or null
    float* stats;                   // [2*Ngemm] or null
    void* out;                      // bf16, or fp32 when out_f32
    int n_total;                    // Ngemm (valid output channels; may be < gridDim.y * BLOCK_N)
    int out_f32;                    // 1: fp32 output (head), columns stored individually
    int act;                        // URIR_ACT_SIGMOID only with out_f32; URIR_ACT_RELU (bf16 output, inference)
    unsigned int* gate;             // deterministic mode: CTAs commit their statistics in blockIdx order (urir_common.cuh)
    IgemmTap taps[36];
};

struct IgemmMaps { CUtensorMap a[4]; CUtensorMap b; };

template <int BLOCK_N, int BLOCK_K, int STAGES>
struct IgemmSmem {
    static constexpr int A_BYTES = 128 * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFF = TILE_BYTES;                      // full[S], empty[S], tmem_full
    static constexpr int STAT_OFF = BAR_OFF + (2 * STAGES + 1) * 8 + 8;   // + tmem ptr slot
    static constexpr int TOTAL = STAT_OFF + 4 * 2 * BLOCK_N * 4 + 1024;  // per-epilogue-warp statistics + alignment slack
};

// sum 16 per-lane values across the 32 lanes of a warp: after the call lanes with (lane&1)==0
// hold in v[0] the total of column col_of_lane(lane) = 8*b4 + 4*b3 + 2*b2 + b1.
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
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
__device__ __forceinline__ int col_of_lane(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

template <int BLOCK_N, int BLOCK_K, int STAGES>
__global__ void __launch_bounds__(192)
conv_igemm_kernel(const __grid_constant__ IgemmMaps maps, const __grid_constant__ IgemmParams p) {
    using L = IgemmSmem<BLOCK_N, BLOCK_K, STAGES>;
    constexpr uint32_t SWZ = (BLOCK_K == 64) ? SWZ_128B : SWZ_64B;
    constexpr uint32_t ROW_BYTES = BLOCK_K * 2;
    constexpr uint32_t SBO = 8 * ROW_BYTES;
    constexpr uint32_t TMEM_COLS = BLOCK_N < 32 ? 32 : BLOCK_N;
    constexpr uint32_t IDESC = make_idesc_bf16(128, BLOCK_N, 0, 0);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
    float* sstats = reinterpret_cast<float*>(smem + L::STAT_OFF);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform by construction

    // tile coordinates
    int t = blockIdx.x;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    const int tn = t;
    const int n_tile = blockIdx.y;
    const int cls = blockIdx.z;
    const int OW = p.cls_OW[cls], OH = p.cls_OH[cls];
    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
    const int tap0 = p.tap_begin[cls], ntaps = p.tap_begin[cls + 1] - tap0;
    const bool tile_live = (w0 < OW) && (h0 < OH);       // classes may be smaller than the grid

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == 4 && lane == 0) {
        prefetch_tmap(&maps.b);
        prefetch_tmap(&maps.a[0]);
    }
    for (int i = threadIdx.x; i < 4 * 2 * BLOCK_N; i += blockDim.x) sstats[i] = 0.f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    const int total_iters = tile_live ? ntaps * p.kchunks : 0;

    if (warp == 4) {
        // ===================== TMA producer =====================
        {   // warp-convergent: one elected lane issues, operands stay in uniform registers
            const uint32_t a_bytes = (uint32_t)(p.bw * p.bh * p.bn) * ROW_BYTES;
            int stage = 0; uint32_t phase = 0;
            uint8_t* a_dst = smem;
            const int b_n0 = n_tile * BLOCK_N;
            for (int ti = 0; ti < (tile_live ? ntaps : 0); ++ti) {
                const IgemmTap tp = p.taps[tap0 + ti];
                const CUtensorMap* am = &maps.a[__shfl_sync(0xffffffffu, (int)tp.map, 0)];
                const int cw = w0 + __shfl_sync(0xffffffffu, (int)tp.dw, 0), ch = h0 + __shfl_sync(0xffffffffu, (int)tp.dh, 0);
                const int wt = __shfl_sync(0xffffffffu, (int)tp.wtap, 0);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    mbar_expect_tx_elect(full_bar + stage, a_bytes + L::B_BYTES);
                    tma_load_4d_elect(am, full_bar + stage, a_dst, kc * BLOCK_K, cw, ch, n0);
                    tma_load_3d_elect(&maps.b, full_bar + stage, a_dst + L::A_BYTES, kc * BLOCK_K, b_n0, wt);
                    a_dst += L::STAGE_BYTES;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; a_dst = smem; }
                }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        int stage = 0; uint32_t phase = 0;
        const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t d_hi = (SBO >> 4) | (1u << 14) | (SWZ << 29);      // SBO | sm100 version | swizzle
        uint32_t a_lo = smem_u32(smem) >> 4;
        for (int it = 0; it < total_iters; ++it) {
            mbar_wait(full_bar + stage, phase);
            fence_after_sync();
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
                const uint64_t ad = ((uint64_t)d_hi << 32) | (a_lo + 2 * k);
                const uint64_t bd = ((uint64_t)d_hi << 32) | (a_lo + (L::A_BYTES >> 4) + 2 * k);
                umma_bf16_elect(tm0, ad, bd, IDESC, (it | k) != 0);
            }
            umma_commit_elect(empty_bar + stage);
            if (it == total_iters - 1) umma_commit_elect(tmem_full_bar);
            __syncwarp();
            a_lo += L::STAGE_BYTES >> 4;
            if (++stage == STAGES) { stage = 0; phase ^= 1; a_lo = smem_u32(smem) >> 4; }
        }
    } else if (warp < 4) {
        // ===================== epilogue =====================
        if (tile_live) {
            mbar_wait(tmem_full_bar, 0);
            fence_after_sync();
            const int row = warp * 32 + lane;
            const int iw = row % p.bw, ih = (row / p.bw) % p.bh, in = row / (p.bw * p.bh);
            const bool valid = (row < p.bw * p.bh * p.bn) && (w0 + iw < OW) && (h0 + ih < OH) && (n0 + in < p.NB);
            const long long o_elem = p.cls_off[cls] + (long long)(n0 + in) * p.o_sn + (long long)(h0 + ih) * p.o_sh +
                                     (long long)(w0 + iw) * p.o_sw + (long long)n_tile * BLOCK_N;
            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + o_elem;
            float* orow_f = reinterpret_cast<float*>(p.out) + o_elem;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
            const int n_left = p.n_total - n_tile * BLOCK_N;           // valid columns of this N tile
#pragma unroll 1
            for (int c0 = 0; c0 < BLOCK_N && c0 < n_left; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(lane_addr + c0, r);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = __uint_as_float(r[j]);
                    if (p.bias && c0 + j < n_left) v[j] += __ldg(p.bias + n_tile * BLOCK_N + c0 + j);
                    if (p.act == URIR_ACT_RELU) v[j] = fmaxf(v[j], 0.f);
                    if (!valid || c0 + j >= n_left) v[j] = 0.f;
                }
                if (valid) {
                    if (p.out_f32) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (c0 + j < n_left) {
                                float o = v[j];
                                if (p.act == URIR_ACT_SIGMOID) o = 1.f / (1.f + __expf(-o));
                                if (p.accumulate) o += orow_f[c0 + j];
                                orow_f[c0 + j] = o;
                            }
                        }
                    } else {
                        uint4 o0, o1;
                        if (p.accumulate) {
                            const uint4 e0 = *reinterpret_cast<const uint4*>(orow + c0), e1 = *reinterpret_cast<const uint4*>(orow + c0 + 8);
                            const uint32_t ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                            uint32_t oo[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) { const float2 f = unpack_bf16x2(ee[j]); oo[j] = pack_bf16x2(v[2 * j] + f.x, v[2 * j + 1] + f.y); }
                            o0 = make_uint4(oo[0], oo[1], oo[2], oo[3]); o1 = make_uint4(oo[4], oo[5], oo[6], oo[7]);
                        } else {
                            o0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                            o1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                        }
                        *reinterpret_cast<uint4*>(orow + c0) = o0;
                        *reinterpret_cast<uint4*>(orow + c0 + 8) = o1;
                    }
                }
                if (p.stats) {
                    float q[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) q[j] = v[j] * v[j];
                    const float s1 = warp_colsum16(v, lane);
                    const float s2 = warp_colsum16(q, lane);
                    if ((lane & 1) == 0) {      // one tile per CTA: every (warp, column) slot is written exactly once
                        const int col = c0 + col_of_lane(lane);
                        sstats[warp * 2 * BLOCK_N + col] = s1;
                        sstats[warp * 2 * BLOCK_N + BLOCK_N + col] = s2;
                    }
                }
            }
            fence_before_sync();
        }
    }
    __syncthreads();
    if (p.stats) {
        gate_enter(p.gate, cta_linear(), threadIdx.x == 0, 0, blockDim.x);
        if (tile_live) {
            for (int i = threadIdx.x; i < 2 * BLOCK_N; i += blockDim.x) {
                const int which = i / BLOCK_N, col = i % BLOCK_N;
                // the four epilogue warps' partial sums, added in a fixed order
                const float v = (sstats[i] + sstats[2 * BLOCK_N + i]) + (sstats[4 * BLOCK_N + i] + sstats[6 * BLOCK_N + i]);
                if (n_tile * BLOCK_N + col < p.n_total)
                    atomicAdd(p.stats + which * p.n_total + n_tile * BLOCK_N + col, v);
            }
        }
        gate_leave(p.gate, cta_linear(), cta_count(), threadIdx.x == 0, 0, blockDim.x);
    }
    if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return fail(URIR_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    // L2 promotion: a tensor that is a channel SLICE of a wider NHWC buffer (skip-concat halves: innermost extent
    // narrower than the pixel pitch) must not promote its fetches past the slice, or every box drags the other
    // half of the buffer through DRAM (ncu, round 1: 188.9 MB read for 94.4 MB of operand). Dense tensors keep 256 B.
    const uint64_t inner_bytes = dims[0] * 2;
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (rank > 1 && inner_bytes < strides_bytes[0])
        promo = inner_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : inner_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
              : inner_bytes >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    { static int force = -2; if (force == -2) { const char* e = getenv("URIR_TMA_PROMO"); force = e ? atoi(e) : -1; }
      if (force == 0) promo = CU_TENSOR_MAP_L2_PROMOTION_NONE; else if (force == 64) promo = CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
      else if (force == 128) promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; else if (force == 256) promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B; }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(URIR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]",
                    (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                    (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                    box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return URIR_OK;
}

// choose the pixel box (bw,bh,bn), bw*bh*bn <= limit, maximising useful rows per M tile; ties go
// to the widest box (longest contiguous runs for TMA and for the epilogue stores)
void choose_box(int OW, int OH, int NB, int limit, int* bw, int* bh, int* bn) {
    double best = -1; int bbw = 1, bbh = 1, bbn = 1;
    for (int w = 1; w <= OW && w <= limit; ++w)
        for (int h = 1; h <= OH && w * h <= limit; ++h) {
            int nmax = limit / (w * h); if (nmax > NB) nmax = NB;
            for (int n = 1; n <= nmax; ++n) {
                const double tiles = (double)((OW + w - 1) / w) * ((OH + h - 1) / h) * ((NB + n - 1) / n);
                const double eff = ((double)OW * OH * NB) / (tiles * limit) + 1e-6 * w + 1e-8 * h;
                if (eff > best) { best = eff; bbw = w; bbh = h; bbn = n; }
            }
        }
    *bw = bbw; *bh = bbh; *bn = bbn;
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

template <int BLOCK_N, int BLOCK_K, int STAGES>
static int launch_cfg(const IgemmMaps& maps, const IgemmParams& p, int n_tiles, cudaStream_t st) {
    using L = IgemmSmem<BLOCK_N, BLOCK_K, STAGES>;
    static std::atomic<bool> attr_set{false};    // benign if two threads both set the attribute
    auto kern = conv_igemm_kernel<BLOCK_N, BLOCK_K, STAGES>;
    if (!attr_set) {
        URIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        attr_set = true;
    }
    dim3 grid(p.tiles_w * p.tiles_h * p.tiles_n, n_tiles, p.n_classes);
    URIR_CUDA_OK(launch_pdl(kern, grid, dim3(192), L::TOTAL, st, maps, p));
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

static int launch_igemm(const IgemmMaps& maps, const IgemmParams& p, int block_n, int block_k, cudaStream_t st) {
    const int n_tiles = (p.n_total + block_n - 1) / block_n;
#define URIR_CFG(BN, BK, ST) if (block_n == BN && block_k == BK) return launch_cfg<BN, BK, ST>(maps, p, n_tiles, st);
    URIR_CFG(128, 64, 3) URIR_CFG(64, 64, 4) URIR_CFG(32, 64, 4)
    URIR_CFG(128, 32, 4) URIR_CFG(64, 32, 4) URIR_CFG(32, 32, 4)
#undef URIR_CFG
    return fail(URIR_ERR_UNSUP, "igemm: no kernel for BLOCK_N=%d BLOCK_K=%d", block_n, block_k);
}

// GEMM-N tile: channel counts that are not a multiple of 32 run a masked 32-wide tile (TMA zero-fills
// the missing weight rows); GEMM-K chunks likewise zero-fill channels past the tensor's extent.
static int pick_block_n(int n) { return (n % 128 == 0) ? 128 : (n % 64 == 0) ? 64 : 32; }
static int pick_block_k(int k) { return (k % 64 == 0) ? 64 : 32; }

bool igemm_fprop_supported(const urir_conv_desc* d) {
    if (d->x_dtype != URIR_BF16 || d->x_ld % 8 || d->x_coff % 8 || d->C % 8 || d->R * d->S > 36) return false;
    if (d->stride != 1 && d->stride != 2) return false;
    if (d->y_dtype == URIR_BF16)
        return d->K % 16 == 0 && d->y_ld % 8 == 0 && d->y_coff % 8 == 0 && (d->act == URIR_ACT_NONE || d->act == URIR_ACT_RELU);
    return d->K <= 32;                       // fp32 output (the sigmoid head): scalar stores, one masked N tile
}
bool igemm_dgrad_supported(const urir_conv_desc* d) {
    return d->x_dtype == URIR_BF16 && d->y_dtype == URIR_BF16 && d->C % 16 == 0 && d->K % 8 == 0 &&
           d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && d->y_ld % 8 == 0 && d->y_coff % 8 == 0 &&
           (d->stride == 1 || d->stride == 2) && d->R * d->S <= 36;
}

// y = conv(x): A = x (parity views when stride 2), B = w_kc [tap][K][C]
int conv_fprop_igemm(const urir_conv_desc* d, const void* x, const void* w_kc, const float* bias, void* y,
                     float* stats, cudaStream_t st) {
    URIR_CHECK_ARG(w_kc != nullptr, "fprop(tcgen05) needs w_kc");
    const int BK = pick_block_k(d->C), BN = pick_block_n(d->K);
    IgemmMaps maps; IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.out_f32 = d->y_dtype == URIR_F32; p.act = d->act;
    choose_box(d->Q, d->P, d->N, 128, &p.bw, &p.bh, &p.bn);
    p.tiles_w = cdiv(d->Q, p.bw); p.tiles_h = cdiv(d->P, p.bh); p.tiles_n = cdiv(d->N, p.bn);
    p.NB = d->N; p.n_classes = 1; p.cls_OW[0] = d->Q; p.cls_OH[0] = d->P; p.cls_off[0] = d->y_coff;
    p.o_sn = (long long)d->P * d->Q * d->y_ld; p.o_sh = (long long)d->Q * d->y_ld; p.o_sw = d->y_ld;
    p.kchunks = (d->C + BK - 1) / BK; p.accumulate = d->accumulate; p.bias = bias; p.stats = stats;
    p.out = y; p.n_total = d->K; p.gate = stats ? next_gate() : nullptr;
    const int s = d->stride;
    const char* xb = (const char*)x + (size_t)d->x_coff * 2;
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    int nt = 0;
    bool used[4] = {false, false, false, false};
    for (int r = 0; r < d->R; ++r)
        for (int q = 0; q < d->S; ++q) {
            const int th = r - d->pad_top, tw = q - d->pad_left;
            const int ph = posmod(th, s), pw = posmod(tw, s);
            IgemmTap& tp = p.taps[nt++];
            tp.dh = (short)floordiv(th, s); tp.dw = (short)floordiv(tw, s);
            tp.map = (short)(ph * s + pw); tp.wtap = (short)(r * d->S + q);
            used[tp.map] = true;
        }
    p.tap_begin[0] = 0; p.tap_begin[1] = nt;
    for (int ph = 0; ph < s; ++ph)
        for (int pw = 0; pw < s; ++pw) {
            const int mi = ph * s + pw;
            if (!used[mi]) { maps.a[mi] = maps.a[0]; continue; }
            const uint64_t dims[4] = {(uint64_t)d->C, (uint64_t)((d->W - pw + s - 1) / s), (uint64_t)((d->H - ph + s - 1) / s), (uint64_t)d->N};
            const uint64_t strides[3] = {(uint64_t)s * d->x_ld * 2, (uint64_t)s * d->W * d->x_ld * 2, (uint64_t)d->H * d->W * d->x_ld * 2};
            int rc = encode_map(&maps.a[mi], xb + ((size_t)ph * d->W + pw) * d->x_ld * 2, 4, dims, strides, box, BK * 2);
            if (rc) return rc;
        }
    if (s == 1) { maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0]; }
    else for (int mi = 0; mi < 4; ++mi) if (!used[mi]) { for (int mj = 0; mj < 4; ++mj) if (used[mj]) { maps.a[mi] = maps.a[mj]; break; } }
    {
        const uint64_t dims[3] = {(uint64_t)d->C, (uint64_t)d->K, (uint64_t)(d->R * d->S)};
        const uint64_t strides[2] = {(uint64_t)d->C * 2, (uint64_t)d->C * d->K * 2};
        const uint32_t bbox[3] = {(uint32_t)BK, (uint32_t)BN, 1};
        int rc = encode_map(&maps.b, w_kc, 3, dims, strides, bbox, BK * 2);
        if (rc) return rc;
    }
    return launch_igemm(maps, p, BN, BK, st);
}

// dx = conv_dgrad(dy): A = dy, B = w_ck [tap][C][K]; stride 2 -> four output parity classes
int conv_dgrad_igemm(const urir_conv_desc* d, const void* dy, const void* w_ck, const float* bias, void* dx,
                     float* stats, cudaStream_t st) {
    URIR_CHECK_ARG(w_ck != nullptr, "dgrad(tcgen05) needs w_ck");
    const int BK = pick_block_k(d->K), BN = pick_block_n(d->C);
    IgemmMaps maps; IgemmParams p;
    memset(&p, 0, sizeof(p));
    const int s = d->stride;
    const int LW = (d->W + s - 1) / s, LH = (d->H + s - 1) / s;        // largest class extent
    choose_box(LW, LH, d->N, 128, &p.bw, &p.bh, &p.bn);
    p.tiles_w = cdiv(LW, p.bw); p.tiles_h = cdiv(LH, p.bh); p.tiles_n = cdiv(d->N, p.bn);
    p.NB = d->N; p.n_classes = s * s;
    p.o_sn = (long long)d->H * d->W * d->x_ld; p.o_sh = (long long)s * d->W * d->x_ld; p.o_sw = (long long)s * d->x_ld;
    p.kchunks = (d->K + BK - 1) / BK; p.accumulate = d->accumulate; p.bias = bias; p.stats = stats;
    p.out = dx; p.n_total = d->C; p.gate = stats ? next_gate() : nullptr;
    int nt = 0;
    for (int pi = 0; pi < s; ++pi)
        for (int pj = 0; pj < s; ++pj) {
            const int c = pi * s + pj;
            p.tap_begin[c] = nt;
            p.cls_OH[c] = (d->H - pi + s - 1) / s; p.cls_OW[c] = (d->W - pj + s - 1) / s;
            p.cls_off[c] = (long long)d->x_coff + ((long long)pi * d->W + pj) * d->x_ld;
            for (int r = 0; r < d->R; ++r) {
                if (posmod(pi + d->pad_top - r, s) != 0) continue;
                for (int q = 0; q < d->S; ++q) {
                    if (posmod(pj + d->pad_left - q, s) != 0) continue;
                    IgemmTap& tp = p.taps[nt++];
                    tp.dh = (short)floordiv(pi + d->pad_top - r, s); tp.dw = (short)floordiv(pj + d->pad_left - q, s);
                    tp.map = 0; tp.wtap = (short)(r * d->S + q);
                }
            }
            p.tap_begin[c + 1] = nt;
        }
    {
        const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->Q, (uint64_t)d->P, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->y_ld * 2, (uint64_t)d->Q * d->y_ld * 2, (uint64_t)d->P * d->Q * d->y_ld * 2};
        const uint32_t box[4] = {(uint32_t)BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
        int rc = encode_map(&maps.a[0], (const char*)dy + (size_t)d->y_coff * 2, 4, dims, strides, box, BK * 2);
        if (rc) return rc;
        maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0];
    }
    {
        const uint64_t dims[3] = {(uint64_t)d->K, (uint64_t)d->C, (uint64_t)(d->R * d->S)};
        const uint64_t strides[2] = {(uint64_t)d->K * 2, (uint64_t)d->C * d->K * 2};
        const uint32_t bbox[3] = {(uint32_t)BK, (uint32_t)BN, 1};
        int rc = encode_map(&maps.b, w_ck, 3, dims, strides, bbox, BK * 2);
        if (rc) return rc;
    }
    return launch_igemm(maps, p, BN, BK, st);
}

}  // namespace urir
