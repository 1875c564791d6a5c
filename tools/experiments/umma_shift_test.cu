// Experiment: can a tcgen05 SWIZZLE_128B K-major A descriptor start at an arbitrary 128-byte row of a
// larger swizzled shared-memory patch (row shift = 3x3 tap offset), with SBO = 2048 picking every
// second 8-row group (16(y) x 8(x) pixel tile out of a pitch-16 patch)?  Tries base_offset = 0 and
// base_offset = (start >> 7) & 7.   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_shift_test umma_shift_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

static constexpr int PATCH_ROWS = 18 * 16 + 16;  // 18 x 16 pixels (+ slack), 128 B each
static constexpr int BN = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift_rows,
                                        int use_base_offset, int sbo_bytes) {
  extern __shared__ uint8_t raw[];
  const uint32_t raw_addr = smem_u32(raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - raw_addr);
  uint8_t* a_s = sm;                                  // PATCH_ROWS x 128 B
  uint8_t* b_s = sm + ((PATCH_ROWS * 128 + 1023) / 1024) * 1024;  // 16 x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // swizzled fill: (row r, 16-byte chunk c) -> r*128 + ((c ^ (r & 7)) << 4)
  for (int i = tid; i < PATCH_ROWS * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(a_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
  }
  for (int i = tid; i < BN * 8; i += 128) {
    int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(b_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_start = base + (uint32_t)shift_rows * 128u;
    uint64_t hi_a = ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
    if (use_base_offset) hi_a |= (uint64_t)((a_start >> 7) & 7u) << 49;
    const uint64_t hi_b = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
    const uint32_t b_start = smem_u32(b_s);
    for (int kk = 0; kk < 4; kk++) {
      uint64_t da = hi_a | (uint64_t)(((a_start + kk * 32) >> 4) & 0x3FFF);
      uint64_t db = hi_b | (uint64_t)(((b_start + kk * 32) >> 4) & 0x3FFF);
      uint32_t accum = kk != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 16; j++) D[tid * 16 + j] = __uint_as_float(v[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> hA(PATCH_ROWS * 64), hB(BN * 64);
  std::vector<float> fA(PATCH_ROWS * 64), fB(BN * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); i++) { float x = (float)(rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); i++) { float x = (float)(rand() % 13 - 6) / 4.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 16 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  size_t smem = 1024 + ((PATCH_ROWS * 128 + 1023) / 1024) * 1024 + BN * 128 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int shifts[] = {0, 1, 2, 7, 8, 16, 17, 18, 32, 33, 34};
  int sbos[] = {1024, 2048};
  for (int sbo : sbos)
    for (int sh : shifts)
      for (int ubo = 0; ubo < 2; ubo++) {
        cudaMemset(dD, 0, 128 * 16 * 4);
        k<<<1, 128, smem>>>(dA, dB, dD, sh, ubo, sbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d shift %d bo %d: CUDA error %s\n", sbo, sh, ubo, cudaGetErrorString(e)); return 1; }
        std::vector<float> hD(128 * 16);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        for (int m = 0; m < 128; m++) {
          int row = sh + (m / 8) * (sbo / 128) + (m % 8);
          for (int n = 0; n < 16; n++) {
            float ref = 0;
            for (int kk = 0; kk < 64; kk++) ref += fA[row * 64 + kk] * fB[n * 64 + kk];
            double err = fabs(ref - hD[m * 16 + n]);
            if (err > 1e-3) bad++;
            if (err > maxerr) maxerr = err;
          }
        }
        printf("sbo %4d shift_rows %2d base_offset_field %d: %s (bad %d / 2048, max err %.4f)\n", sbo, sh, ubo, bad ? "MISMATCH" : "OK", bad, maxerr);
      }
  return 0;
}
