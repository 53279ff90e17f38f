// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, operand major-ness,
// swizzle, LBO/SBO and number of accumulators. smem content is garbage (never read back).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../unet-rir_b200/csrc/urir_tc.cuh"
using namespace urir::tc;

struct Cfg { int N, a_mn, b_mn, a_swz, b_swz, a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step, nacc, a_stride_acc; };

template <int NACC>
__global__ void __launch_bounds__(128) bench(Cfg c, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    fence_proxy_async();
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, c.N, c.a_mn, c.b_mn);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
        uint64_t ad[4][NACC], bd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bd[k] = make_smem_desc(b0 + k * c.b_step, c.b_lbo, c.b_sbo, c.b_swz);
#pragma unroll
            for (int j = 0; j < NACC; ++j) ad[k][j] = make_smem_desc(a0 + (NACC == 1 ? c.a_stride_acc : j * c.a_stride_acc) + k * c.a_step, c.a_lbo, c.a_sbo, c.a_swz);
        }
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int j = 0; j < NACC; ++j) umma_bf16(tm + j * c.N, ad[k][j], bd[k], idesc, 1);
        }
        long long t_issue = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue - t0; }
    }
    fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) { fence_after_sync(); tmem_dealloc(tm, 512); }
}

static void run(const char* name, Cfg c, int iters = 2000) {
    long long* d; cudaMalloc(&d, 16);
    if (c.nacc == 1) { cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); bench<1><<<148, 128, 200 * 1024>>>(c, iters, d); }
    else { cudaFuncSetAttribute(bench<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); bench<3><<<148, 128, 200 * 1024>>>(c, iters, d); }
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-58s N=%3d acc=%d : %7.1f cyc/UMMA (issue %6.1f)  %s\n", name, c.N, c.nacc, (double)h[0] / (iters * 4.0 * c.nacc),
           (double)h[1] / (iters * 4.0 * c.nacc), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    // halo-style A operands (conv_halo.cu): start rows not aligned to the 8-row swizzle group, SBO = 10 rows
    for (int N : {32, 64, 128}) {
        for (int off_rows : {0, 1, 3, 10, 11}) {
            char nm[96];
            snprintf(nm, sizeof(nm), "halo K-major SW64  SBO=640  start +%d rows (A), B aligned", off_rows);
            run(nm, {N, 0, 0, SWZ_64B, SWZ_64B, 0, 640, 0, 512, 32, 32, 1, 8192 + off_rows * 64});
            snprintf(nm, sizeof(nm), "halo K-major SW128 SBO=1280 start +%d rows (A), B aligned", off_rows);
            run(nm, {N, 0, 0, SWZ_128B, SWZ_128B, 0, 1280, 0, 1024, 32, 32, 1, 16384 + off_rows * 128});
        }
    }
    for (int N : {32, 64, 128, 256}) {
        for (int nacc : {1, 3}) {
            if (N * nacc > 512) continue;
            // K-major SW128: rows 128B, SBO 1024, k-step 32B
            run("K-major SW128 (A,B)", {N, 0, 0, SWZ_128B, SWZ_128B, 0, 1024, 0, 1024, 32, 32, nacc, 16384});
            // K-major SW64: rows 64B, SBO 512
            run("K-major SW64 (A,B)", {N, 0, 0, SWZ_64B, SWZ_64B, 0, 512, 0, 512, 32, 32, nacc, 8192});
            // MN-major SW128, my layout: atoms [64 rows x 128B], LBO = 8192, SBO = 1024, k-step = 16 rows = 2048B
            run("MN-major SW128 LBO8192 SBO1024 (A,B)", {N, 1, 1, SWZ_128B, SWZ_128B, 8192, 1024, 8192, 1024, 2048, 2048, nacc, 16384});
            // MN-major SW128, cutlass order: k-group major: LBO = 1024, SBO = 2048 (M=128 -> 2 atoms per k-group)
            run("MN-major SW128 LBO1024 SBO2048 (A) / B same (N/64 atoms)", {N, 1, 1, SWZ_128B, SWZ_128B, 1024, 2048, 1024, (N / 64 > 0 ? N / 64 : 1) * 1024, 4096, (N / 64 > 0 ? N / 64 : 1) * 2048, nacc, 16384});
            // MN-major SW64, my layout: atoms [64 rows x 64B] = 4096, LBO 4096, SBO 512, k-step 1024
            run("MN-major SW64 LBO4096 SBO512 (A,B)", {N, 1, 1, SWZ_64B, SWZ_64B, 4096, 512, 4096, 512, 1024, 1024, nacc, 16384});
            // MN-major SW64 with aliased atoms (LBO 0)
            run("MN-major SW64 LBO0 (A) B LBO4096", {N, 1, 1, SWZ_64B, SWZ_64B, 0, 512, 4096, 512, 1024, 1024, nacc, 16384});
            // A MN-major SW128, B K-major SW128
            run("A MN-major SW128 (8192/1024), B K-major SW128", {N, 1, 0, SWZ_128B, SWZ_128B, 8192, 1024, 0, 1024, 2048, 32, nacc, 16384});
            // A K-major SW128, B MN-major SW128
            run("A K-major SW128, B MN-major SW128 (8192/1024)", {N, 0, 1, SWZ_128B, SWZ_128B, 0, 1024, 8192, 1024, 32, 2048, nacc, 16384});
            // MN-major SW128 with padded atom stride (LBO 8192+128): bank-conflict probe
            run("MN-major SW128 LBO8320 SBO1024 (A,B) [pad probe]", {N, 1, 1, SWZ_128B, SWZ_128B, 8320, 1024, 8320, 1024, 2048, 2048, nacc, 17000});
        }
    }
    return 0;
}
