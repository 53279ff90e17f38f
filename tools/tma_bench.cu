// Micro-benchmark: TMA (cp.async.bulk.tensor.4d) box-load issue throughput per SM as a function of the
// number of issuing warps, the box shape and the depth in flight. One CTA per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_bench tools/tma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>
#include "../unet-rir_b200/csrc/urir_tc.cuh"
using namespace urir::tc;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int DEPTH = 4;      // boxes in flight per issuing warp

__global__ void __launch_bounds__(256) bench(const __grid_constant__ CUtensorMap tm, int nwarps, int boxes_per_warp,
                                             int box_bytes, int W, int H, int bw, int bh, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[8][DEPTH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int w = 0; w < 8; ++w) for (int d = 0; d < DEPTH; ++d) mbar_init(&bars[w][d], 1); fence_barrier_init(); }
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps && lane == 0) {
        const int tiles_w = W / bw, tiles_h = H / bh;
        int t = (blockIdx.x * nwarps + warp) * 977;
        for (int i = 0; i < boxes_per_warp; ++i) {
            const int d = i % DEPTH;
            if (i >= DEPTH) mbar_wait(&bars[warp][d], ((i / DEPTH) - 1) & 1);
            mbar_expect_tx(&bars[warp][d], box_bytes);
            const int tw = t % tiles_w, th = (t / tiles_w) % tiles_h, tn = (t / (tiles_w * tiles_h)) % 64;
            tma_load_4d(&tm, &bars[warp][d], smem + ((warp * DEPTH + d) % 8) * 24576, 0, tw * bw, th * bh, tn);
            t += 1;
        }
        for (int i = boxes_per_warp; i < boxes_per_warp + DEPTH; ++i) {
            const int d = i % DEPTH;
            if (i >= DEPTH) mbar_wait(&bars[warp][d], ((i / DEPTH) - 1) & 1);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    const int N = 64, H = 144, W = 160;
    long long* dout; cudaMalloc(&dout, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 24576 + 2048);
    for (int C : {32, 64}) {
        void* buf; size_t bytes = (size_t)N * H * W * C * 2; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
        struct { int bw, bh; } shapes[] = {{16, 4}, {16, 8}, {8, 8}, {8, 16}, {10, 18}, {32, 4}};
        for (auto sh : shapes) {
            CUtensorMap tm;
            cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
            cuuint64_t gs[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
            cuuint32_t bx[4] = {(cuuint32_t)C, (cuuint32_t)sh.bw, (cuuint32_t)sh.bh, 1}, es[4] = {1, 1, 1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            const int box_bytes = C * 2 * sh.bw * sh.bh;
            for (int nw : {1, 2, 4, 8}) {
                const int per = 2048 / nw;
                bench<<<148, 256, 8 * 24576 + 2048>>>(tm, nw, per, box_bytes, W, H, sh.bw, sh.bh, dout);
                cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
                long long h = 0; cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
                printf("C=%2d box %2dx%2d (%5d B, %3d rows) warps=%d depth=%d : %7.1f cyc/box/SM  %6.1f B/cyc/SM  %s\n", C, sh.bw, sh.bh,
                       box_bytes, sh.bw * sh.bh, nw, DEPTH, (double)h / (per * nw), (double)box_bytes * per * nw / h,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
        cudaFree(buf);
    }
    return 0;
}
