// Probe 2: (a) K-major SW64 A operand (64 B rows) at unaligned start rows / SBO = 640 (halo reuse for C = 32);
//          (b) MN-major SW128 A operand with SBO = 1280 and start-row offsets (halo reuse for wgrad);
//          (c) MN-major SW64 A operand, 4 atoms of 32 channels overlapping at LBO = 64 B (taps stacked along M).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_halo_test2 tools/umma_halo_test2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../unet-rir_b200/csrc/urir_tc.cuh"
using namespace urir::tc;

constexpr int ROWS = 512;

// mode 0: A K-major SW64 (rows 64 B, 32 k), B K-major SW64 identity-ish, M=128 N=32 K=32 (2 steps)
// mode 1: A MN-major SW128 (rows = k index (pixels) of 128 B = 64 m), M=64x2 atoms via LBO, K=16 pixels per step, 2 steps
// mode 2: A MN-major SW64 (rows = pixels of 64 B = 32 m), M=128 = 4 atoms at LBO bytes apart
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* a_rows, int mode, int start_row, int sbo, int lbo, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;
    uint8_t* sb = smem + ROWS * 128;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 32);
    const int rowb = (mode == 1) ? 128 : 64, epr = rowb / 2;
    for (int i = threadIdx.x; i < ROWS * epr; i += blockDim.x) {
        int r = i / epr, e = i % epr;
        int chunk = e / 8;
        int sw = (rowb == 128) ? (chunk ^ (r % 8)) : (chunk ^ ((r >> 1) & 3));
        *(__nv_bfloat16*)(sa + r * rowb + sw * 16 + (e % 8) * 2) = a_rows[i];
    }
    if (mode == 0) {   // B[n][k] K-major SW64, 32 x 32: identity
        for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
            int n = i / 32, kk = i % 32; int chunk = (kk / 8) ^ ((n >> 1) & 3);
            *(__nv_bfloat16*)(sb + n * 64 + chunk * 16 + (kk % 8) * 2) = __float2bfloat16(kk == n ? 1.f : 0.f);
        }
    } else {           // B MN-major SW64: rows = k (32 of them), 64 B = 32 n; B[k][n] = (n == k % 32)
        for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
            int k = i / 32, n = i % 32; int chunk = (n / 8) ^ ((k >> 1) & 3);
            *(__nv_bfloat16*)(sb + k * 64 + chunk * 16 + (n % 8) * 2) = __float2bfloat16(n == k ? 1.f : 0.f);
        }
    }
    fence_proxy_async();
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        if (mode == 0) {
            const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
            for (int k = 0; k < 2; ++k) {
                uint64_t ad = make_smem_desc(smem_u32(sa) + start_row * 64 + k * 32, 0, sbo, SWZ_64B);
                uint64_t bd = make_smem_desc(smem_u32(sb) + k * 32, 0, 512, SWZ_64B);
                umma_bf16(tm, ad, bd, idesc, k != 0);
            }
        } else {
            const uint32_t idesc = make_idesc_bf16(128, 32, 1, 1);
            for (int k = 0; k < 2; ++k) {      // k-step = 16 rows of A = 2 groups of 8 rows at SBO
                uint64_t ad = make_smem_desc(smem_u32(sa) + start_row * rowb + k * 2 * sbo, lbo, sbo, mode == 1 ? SWZ_128B : SWZ_64B);
                uint64_t bd = make_smem_desc(smem_u32(sb) + k * 1024, 2048, 512, SWZ_64B);
                umma_bf16(tm, ad, bd, idesc, k != 0);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    const int warp = threadIdx.x / 32;
    for (int c0 = 0; c0 < 32; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 32 + c0 + j] = __uint_as_float(r[j]);
    }
    fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) { fence_after_sync(); tmem_dealloc(tm, 32); }
}

int main() {
    float* dout; cudaMalloc(&dout, 128 * 32 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    for (int mode = 0; mode < 3; ++mode) {
        const int epr = (mode == 1) ? 64 : 32;
        std::vector<__nv_bfloat16> ha(ROWS * epr); std::vector<float> fa(ROWS * epr);
        for (int i = 0; i < ROWS * epr; ++i) { fa[i] = (float)((i * 7 + (i / epr) * 3) % 17 - 8); ha[i] = __float2bfloat16(fa[i]); }
        __nv_bfloat16* da; cudaMalloc(&da, ha.size() * 2);
        cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
        const int rowb = epr * 2;
        struct Cfg { int start, sbo, lbo; };
        std::vector<Cfg> cfgs;
        if (mode == 0) for (int sbo : {512, 640, 576}) for (int st : {0, 1, 3, 10, 21}) cfgs.push_back({st, sbo, 0});
        if (mode == 1) for (int sbo : {1024, 1280}) for (int st : {0, 1, 3, 10, 21}) cfgs.push_back({st, sbo, 180 * 128});
        if (mode == 2) for (int sbo : {512, 640}) for (int st : {0, 1, 10}) for (int lbo : {64, 128, 640}) cfgs.push_back({st, sbo, lbo});
        for (auto c : cfgs) {
            probe<<<1, 128, 96 * 1024>>>(da, mode, c.start, c.sbo, c.lbo, dout);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> ho(128 * 32);
            cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            const int rpg = c.sbo / rowb;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
                float exp = 0.f;
                if (mode == 0) { int r = c.start + (m / 8) * rpg + (m % 8); exp = fa[r * 32 + n]; }          // D[m][n] = A[m][n]
                else {
                    // D[m][n] = sum_k A[m][k] B[k][n] = A[m][k = n] ; A[m][k] lives at smem row r(k), element index within row
                    int k = n; int r = c.start + (k / 8) * rpg + (k % 8);
                    if (mode == 1) { int atom = m / 64; r += atom * (c.lbo / 128); exp = fa[r * 64 + (m % 64)]; }
                    else { int atom = m / 32; r += atom * (c.lbo / 64); exp = fa[r * 32 + (m % 32)]; }
                }
                if (ho[m * 32 + n] != exp) ++bad;
            }
            printf("mode %d start_row=%2d SBO=%4d LBO=%5d : %s (%d/4096 wrong) %s\n", mode, c.start, c.sbo, c.lbo, bad ? "MISMATCH" : "ok", bad,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
        cudaFree(da);
    }
    return 0;
}
