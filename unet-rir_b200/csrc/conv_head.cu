// The output head Conv2D(F0 = 32 -> 2, 6x6, SAME) + sigmoid (dl_models/u_net.py:247-249) on tcgen05.
//
// With only two output channels a plain implicit GEMM (taps as GEMM-K, N = 2 padded to 16/32) spends one
// 44-cycle MMA per (tap, 16 channels): 72 per 128 pixels, and re-fetches the activation tile once per tap.
// Here the vertical taps stay in GEMM-K but the HORIZONTAL taps move into GEMM-N:
//     Y[p, (s, k)] = sum over (r, c) of  x[p + (r - pad_top, 0), c] * w[r, s, c, k]        N = S*2 = 12 -> 16
//     out[oh, ow, k] = sigmoid(bias[k] + sum over s of  Y[(oh, ow + s - pad_left), (s, k)])
// so one MMA covers six taps (12 per 128 pixels) and the shifted sum over s is a 5-step warp-shuffle chain
// in the epilogue (a warp = 32 consecutive columns of one output row).
// Vertical taps cost no extra loads either: ONE TMA box {32 ch, 32 cols, 16+R-1 rows} per region lands as
// rows of 64 B (64B swizzle); M-tile q (output rows 4q..4q+3) and tap r read it through a UMMA descriptor
// that simply starts (4q + r) * 2048 bytes into the box -- 1024-byte aligned, so the swizzle phase is kept.
// Warp roles (192 threads, persistent, one CTA per SM): warps 0-3 epilogue (TMEM lane quarter = warp),
// warp 4 TMA producer, warp 5 MMA issuer; 3 smem stages; two TMEM accumulator stages of 4 x 16 columns so
// the epilogue of region i overlaps the MMAs of region i+1.
#include <atomic>
#include "urir_common.cuh"
#include "urir_tc.cuh"

namespace urir {

using namespace tc;

int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, int swizzle_bytes);

constexpr int HF_RW = 32;               // region width in input columns (= one warp of output lanes)
constexpr int HF_RH = 16;               // output rows per region (4 M-tiles of 4 rows)
constexpr int HF_STAGES = 3;
constexpr int HF_ROWB = HF_RW * 64;     // bytes per region row (32 pixels x 32 ch x bf16)

struct HeadParams {
    int N, H, W, R, S, pt, pl;
    int out_w;                          // output columns per region = HF_RW - (S - 1)
    int tiles_w, tiles_h, total;
    int stage_bytes;
    const __nv_bfloat16* w_ck;          // [tap][32][2]
    const float* bias;
    float* out; int out_ld, out_coff;
    int act;
};

__global__ void __launch_bounds__(192)
head_fprop_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ HeadParams p) {
    constexpr uint32_t IDESC = make_idesc_bf16(128, 16, 0, 0);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sB = smem;                                   // R tiles of 16 rows x 64 B
    uint8_t* sA = smem + 8192;                            // HF_STAGES region boxes
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + HF_STAGES * p.stage_bytes);
    uint64_t* empty_bar = full_bar + HF_STAGES;
    uint64_t* tfull_bar = empty_bar + HF_STAGES;          // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                 // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < HF_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar + a, 1); mbar_init(tempty_bar + a, 4); }
        fence_barrier_init();
        prefetch_tmap(&xmap);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    // B_r[n = (s, k)][c] = w[r, s, c, k], rows of 64 B, 16-byte chunks swizzled with (row / 2) % 4 (64B swizzle)
    for (int idx = threadIdx.x; idx < p.R * 16 * 4; idx += blockDim.x) {
        const int ch = idx & 3, n = (idx >> 2) & 15, r = idx >> 6;
        const int s = n >> 1, k = n & 1;
        uint32_t wd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float a = 0.f, b = 0.f;
            if (s < p.S) {
                const int c = ch * 8 + 2 * e;
                a = bf2f(p.w_ck[((r * p.S + s) * 32 + c) * 2 + k]);
                b = bf2f(p.w_ck[((r * p.S + s) * 32 + c + 1) * 2 + k]);
            }
            wd[e] = pack_bf16x2(a, b);
        }
        *reinterpret_cast<uint4*>(sB + r * 1024 + n * 64 + ((ch ^ ((n >> 1) & 3)) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===================== TMA producer: one box per region =====================
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < p.total; t += gridDim.x) {
            const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, n = t / (p.tiles_w * p.tiles_h);
            mbar_wait(empty_bar + stage, phase ^ 1);
            mbar_expect_tx_elect(full_bar + stage, (uint32_t)p.stage_bytes);
            tma_load_4d_elect(&xmap, full_bar + stage, sA + stage * p.stage_bytes, 0, tw * p.out_w - p.pl, th * HF_RH - p.pt, n);
            if (++stage == HF_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t d_hi = (512u >> 4) | (1u << 14) | (SWZ_64B << 29);
        const uint32_t a0 = smem_u32(sA) >> 4, b0 = smem_u32(sB) >> 4;
        int stage = 0; uint32_t phase = 0;
        int it = 0;
        for (int t = blockIdx.x; t < p.total; t += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(tempty_bar + acc, ((it >> 1) & 1) ^ 1);
            mbar_wait(full_bar + stage, phase);
            fence_after_sync();
            const uint32_t a_st = a0 + ((stage * p.stage_bytes) >> 4);
            if (p.R == 6) {
                // the model's 6x6 head: 48 MMAs per region issued as straight-line code with immediate offsets. This warp is
                // the only issuer, and the rolled loops below cost ~84 cycles per MMA against 44 on the tensor pipe
                // (measured on the halo kernel's identical issue sequence, profiles/r02_halo_trace.txt)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            const uint64_t ad = ((uint64_t)d_hi << 32) | (a_st + (uint32_t)((((4 * q + r) * HF_ROWB) >> 4) + 2 * kk));
                            const uint64_t bd = ((uint64_t)d_hi << 32) | (b0 + (uint32_t)(((r * 1024) >> 4) + 2 * kk));
                            umma_bf16_elect(tm0 + acc * 64 + q * 16, ad, bd, IDESC, (r | kk) != 0);
                        }
                    }
                }
            } else {
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
#pragma unroll 1
                for (int r = 0; r < p.R; ++r) {
                    const uint32_t a_lo = a_st + (((4 * q + r) * HF_ROWB) >> 4);
                    const uint32_t b_lo = b0 + ((r * 1024) >> 4);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint64_t ad = ((uint64_t)d_hi << 32) | (a_lo + 2 * kk);
                        const uint64_t bd = ((uint64_t)d_hi << 32) | (b_lo + 2 * kk);
                        umma_bf16_elect(tm0 + acc * 64 + q * 16, ad, bd, IDESC, (r | kk) != 0);
                    }
                }
            }
            }
            umma_commit_elect(empty_bar + stage);
            umma_commit_elect(tfull_bar + acc);
            __syncwarp();
            if (++stage == HF_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp < 4) {
        // ===================== epilogue: shifted sum over s, bias, sigmoid, fp32 store =====================
        const float b0 = p.bias ? __ldg(p.bias) : 0.f, b1 = p.bias ? __ldg(p.bias + 1) : 0.f;
        int it = 0;
        for (int t = blockIdx.x; t < p.total; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, n = t / (p.tiles_w * p.tiles_h);
            mbar_wait(tfull_bar + acc, (it >> 1) & 1);
            fence_after_sync();
            const int ow = tw * p.out_w + lane;
            const bool col_ok = lane < p.out_w && ow < p.W;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 64 + q * 16, r);
                tmem_ld_wait();
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int s = 7; s >= 0; --s) {
                    if (s < p.S) {
                        t0 = __uint_as_float(r[2 * s]) + __shfl_down_sync(0xffffffffu, t0, 1);
                        t1 = __uint_as_float(r[2 * s + 1]) + __shfl_down_sync(0xffffffffu, t1, 1);
                    }
                }
                const int oh = th * HF_RH + 4 * q + warp;
                if (col_ok && oh < p.H) {
                    float o0 = t0 + b0, o1 = t1 + b1;
                    if (p.act == URIR_ACT_SIGMOID) { o0 = 1.f / (1.f + __expf(-o0)); o1 = 1.f / (1.f + __expf(-o1)); }
                    *reinterpret_cast<float2*>(p.out + ((size_t)(n * p.H + oh) * p.W + ow) * p.out_ld + p.out_coff) = make_float2(o0, o1);
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar + acc);
        }
    }
    __syncthreads();
    if (warp == 1) { fence_after_sync(); tmem_dealloc(tmem_base, 128); }
}

bool head_fprop_supported(const urir_conv_desc* d) {
    return d->x_dtype == URIR_BF16 && d->y_dtype == URIR_F32 && d->C == 32 && d->K == 2 && d->stride == 1 &&
           d->P == d->H && d->Q == d->W && d->R <= 6 && d->S <= 8 && d->S >= 1 && !d->accumulate &&
           d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && d->y_ld % 2 == 0 && d->y_coff % 2 == 0 &&
           (d->act == URIR_ACT_NONE || d->act == URIR_ACT_SIGMOID);
}

int head_fprop(const urir_conv_desc* d, const void* x, const void* w_ck, const float* bias, void* y, cudaStream_t st) {
    URIR_CHECK_ARG(w_ck != nullptr, "head fprop needs w_ck");
    HeadParams p; memset(&p, 0, sizeof(p));
    p.N = d->N; p.H = d->H; p.W = d->W; p.R = d->R; p.S = d->S; p.pt = d->pad_top; p.pl = d->pad_left;
    p.out_w = HF_RW - (d->S - 1);
    p.tiles_w = cdiv(d->W, p.out_w); p.tiles_h = cdiv(d->H, HF_RH); p.total = p.tiles_w * p.tiles_h * d->N;
    const int rows = HF_RH + d->R - 1;
    p.stage_bytes = rows * HF_ROWB;
    p.w_ck = (const __nv_bfloat16*)w_ck; p.bias = bias; p.out = (float*)y; p.out_ld = d->y_ld; p.out_coff = d->y_coff; p.act = d->act;
    CUtensorMap map;
    {
        const uint64_t dims[4] = {32, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
        const uint64_t strides[3] = {(uint64_t)d->x_ld * 2, (uint64_t)d->W * d->x_ld * 2, (uint64_t)d->H * d->W * d->x_ld * 2};
        const uint32_t box[4] = {32, HF_RW, (uint32_t)rows, 1};
        int rc = encode_map(&map, (const char*)x + (size_t)d->x_coff * 2, 4, dims, strides, box, 64);
        if (rc) return rc;
    }
    const int smem = 8192 + HF_STAGES * p.stage_bytes + 128 + 1024;
    static std::atomic<bool> attr_set{false};    // benign if two threads both set the attribute
    if (!attr_set) { URIR_CUDA_OK(cudaFuncSetAttribute(head_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + HF_STAGES * 21 * HF_ROWB + 128 + 1024)); attr_set = true; }
    const int grid = p.total < sm_count() ? p.total : sm_count();
    head_fprop_kernel<<<grid, 192, smem, st>>>(map, p);
    URIR_LAUNCH_OK(1);
    return URIR_OK;
}

}  // namespace urir
