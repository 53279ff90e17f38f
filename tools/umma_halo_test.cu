// Probe: can a K-major SW128 A operand start at a 128B-multiple (not 1024B-aligned) smem address with an SBO
// that is not a multiple of 1024 (halo-tile reuse: one TMA box per conv tile, 9 taps as descriptor offsets)?
// smem is filled with the layout TMA would write (XOR of the 16B chunk index with address bits [7,9]).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../unet-rir_b200/csrc/urir_tc.cuh"
using namespace urir::tc;

constexpr int ROWS = 256;   // smem rows of 128 B available to A

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* a_rows, int start_row, int sbo, int base_off_field, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                    // ROWS x 128 B
    uint8_t* sb = smem + ROWS * 128;       // 32 rows x 128 B  (B[n][k], k-major)
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 32);
    // A: logical smem row r, element kk -> swizzled position
    for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) {
        int r = i / 64, kk = i % 64;
        int chunk = (kk / 8) ^ (r % 8);
        *(__nv_bfloat16*)(sa + r * 128 + chunk * 16 + (kk % 8) * 2) = a_rows[i];
    }
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
        int n = i / 64, kk = i % 64;
        int chunk = (kk / 8) ^ (n % 8);
        *(__nv_bfloat16*)(sb + n * 128 + chunk * 16 + (kk % 8) * 2) = __float2bfloat16((kk % 32) == n ? 1.f : 0.f);
    }
    fence_proxy_async();
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
        for (int k = 0; k < 4; ++k) {
            uint64_t ad = make_smem_desc(smem_u32(sa) + start_row * 128 + k * 32, 0, sbo, SWZ_128B) | ((uint64_t)(base_off_field & 7) << 49);
            uint64_t bd = make_smem_desc(smem_u32(sb) + k * 32, 0, 1024, SWZ_128B);
            umma_bf16(tm, ad, bd, idesc, k != 0);
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
    std::vector<__nv_bfloat16> ha(ROWS * 64);
    std::vector<float> fa(ROWS * 64);
    for (int i = 0; i < ROWS * 64; ++i) { fa[i] = (float)((i * 7 + (i / 64) * 3) % 17 - 8); ha[i] = __float2bfloat16(fa[i]); }
    __nv_bfloat16* da; float* dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&dout, 128 * 32 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    int sbos[] = {1024, 1280, 1152, 2304};
    for (int sbo : sbos) for (int start : {0, 1, 2, 3, 5, 11}) for (int bo : {0, -1}) {
        int rows_per_group = sbo / 128;
        if (start + 15 * rows_per_group + 8 > ROWS) continue;
        int bofield = bo == 0 ? 0 : (start % 8);
        probe<<<1, 128, 64 * 1024>>>(da, start, sbo, bofield, dout);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> ho(128 * 32);
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; 
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
            int r = start + (m / 8) * rows_per_group + (m % 8);
            float exp = fa[r * 64 + n] + fa[r * 64 + n + 32];
            if (ho[m * 32 + n] != exp) ++bad;
        }
        printf("SBO=%4d start_row=%2d base_offset=%d : %s (%d/4096 wrong) %s\n", sbo, start, bofield, bad ? "MISMATCH" : "ok", bad,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
