// Probe (round 2): does a SWIZZLE_128B TMA box written to a shared-memory address that is 128-byte but NOT 1024-byte
// aligned land with the swizzle anchored to ABSOLUTE address bits (chunk j of the 128-byte row at address a goes to
// j ^ ((a >> 7) & 7)), i.e. the same convention tcgen05.mma reads with (profiles/r01_umma_halo_probe.txt)?
// The deep-layer kernel (conv_deep.cu) packs one {64 ch x (W+2) cols x 1 row} box per padded image row back to back,
// so rows start at arbitrary multiples of 128 bytes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_row_probe tools/tma_row_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include "../unet-rir_b200/csrc/urir_tc.cuh"
using namespace urir::tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int WP = 22, C = 64, W = 20, H = 18, N = 2;
constexpr int DUMP_ROWS = 96;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int off_rows, int h, int n, uint32_t* dump, uint32_t* base_out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < DUMP_ROWS * 32; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xDEADBEEFu;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, WP * C * 2);
        tma_load_4d(&tm, &bar, smem + off_rows * 128, 0, -1, h, n);
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DUMP_ROWS * 32; i += blockDim.x) dump[i] = reinterpret_cast<uint32_t*>(smem)[i];
    if (threadIdx.x == 0) *base_out = smem_u32(smem);
}

int main() {
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    // element (n, h, w, c) = bf16 bit pattern encoding pixel index and channel: value = (pixel << 6 | c) as uint16
    std::vector<uint16_t> host((size_t)N * H * W * C);
    for (int n = 0; n < N; ++n) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c)
        host[(((size_t)n * H + h) * W + w) * C + c] = (uint16_t)((((n * H + h) * W + w) & 1023) << 6 | c);
    void* dbuf; cudaMalloc(&dbuf, host.size() * 2); cudaMemcpy(dbuf, host.data(), host.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t gd[4] = {C, W, H, N}; cuuint64_t gs[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t bx[4] = {C, WP, 1, 1}; cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dbuf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    uint32_t *ddump, *dbase; cudaMalloc(&ddump, DUMP_ROWS * 128); cudaMalloc(&dbase, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, DUMP_ROWS * 128 + 2048);
    std::vector<uint32_t> dump(DUMP_ROWS * 32);
    int offs[] = {0, 1, 3, 5, 8, 22, 23, 45};
    for (int off : offs) {
        for (int hh : {3, -1}) {
            probe<<<1, 128, DUMP_ROWS * 128 + 2048>>>(tm, off, hh, 1, ddump, dbase);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("off %d h %d: CUDA error %s\n", off, hh, cudaGetErrorString(e)); return 1; }
            uint32_t base; cudaMemcpy(&base, dbase, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(dump.data(), ddump, DUMP_ROWS * 128, cudaMemcpyDeviceToHost);
            const uint16_t* d16 = reinterpret_cast<const uint16_t*>(dump.data());
            int bad_abs = 0, bad_rel = 0, total = 0;
            for (int row = 0; row < WP; ++row) {
                const int w = row - 1;
                const uint32_t row_addr = base + (off + row) * 128;
                for (int c = 0; c < C; ++c) {
                    uint16_t want = 0;
                    if (w >= 0 && w < W && hh >= 0) want = (uint16_t)((((1 * H + hh) * W + w) & 1023) << 6 | c);
                    const int j = c >> 3, e8 = c & 7;
                    const int j_abs = j ^ ((row_addr >> 7) & 7);          // anchored to the absolute address
                    const int j_rel = j ^ (row & 7);                       // anchored to the box start
                    const uint16_t got_abs = d16[((off + row) * 128 + j_abs * 16) / 2 + e8];
                    const uint16_t got_rel = d16[((off + row) * 128 + j_rel * 16) / 2 + e8];
                    bad_abs += got_abs != want; bad_rel += got_rel != want; ++total;
                }
            }
            // bytes outside the box must be untouched
            int touched = 0;
            for (int row = 0; row < DUMP_ROWS; ++row)
                if (row < off || row >= off + WP)
                    for (int k = 0; k < 32; ++k) touched += dump[row * 32 + k] != 0xDEADBEEFu;
            printf("dst = base + %2d rows (base %% 1024 = %u), h = %2d : absolute-address swizzle %s (%d/%d wrong), box-relative swizzle %s (%d/%d wrong), words outside box touched: %d\n",
                   off, base & 1023, hh, bad_abs ? "MISMATCH" : "ok", bad_abs, total, bad_rel ? "MISMATCH" : "ok", bad_rel, total, touched);
        }
    }
    return 0;
}
