// Experiment: tcgen05.mma issue rate (cycles per M128 x N x K16 instruction) as a function of the A
// descriptor's alignment inside a SWIZZLE_128B patch (row shift, SBO) and of N.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) k(long long* out, int shift_rows, int sbo_bytes, int BN, int iters, int a_layout, int row_bytes) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(raw + (base - smem_u32(raw)))[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_start = base + (uint32_t)shift_rows * row_bytes;
    const uint64_t hi_a = ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)a_layout << 61) | ((uint64_t)1 << 16);
    const uint64_t hi_b = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
    const uint32_t b_start = base + 49152;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
      uint64_t da = hi_a | (uint64_t)(((a_start + (it & 3) * 32 * (row_bytes >= 128)) >> 4) & 0x3FFF);
      uint64_t db = hi_b | (uint64_t)(((b_start + (it & 3) * 32) >> 4) & 0x3FFF);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  struct V { int shift, sbo, layout, rowb; const char* name; } vs[] = {
    {0, 1024, 2, 128, "128B aligned sbo1024"}, {1, 1024, 2, 128, "128B shift1 sbo1024"}, {2, 1024, 2, 128, "128B shift2 sbo1024"},
    {0, 1280, 2, 128, "128B aligned sbo1280"}, {1, 1280, 2, 128, "128B shift1 sbo1280"}, {11, 1280, 2, 128, "128B shift11 sbo1280"},
    {0, 2048, 2, 128, "128B aligned sbo2048"}, {1, 2048, 2, 128, "128B shift1 sbo2048"}, {17, 2048, 2, 128, "128B shift17 sbo2048"},
    {0, 640, 4, 64, "64B sbo640"}, {11, 640, 4, 64, "64B shift11 sbo640"}, {0, 512, 4, 64, "64B sbo512"},
    {0, 320, 6, 32, "32B sbo320"}, {11, 320, 6, 32, "32B shift11 sbo320"}, {0, 256, 6, 32, "32B sbo256"},
  };
  int Ns[] = {16, 64, 128, 256};
  for (auto& v : vs)
    for (int N : Ns) {
      k<<<1, 128, 200 * 1024>>>(d, v.shift, v.sbo, N, iters, v.layout, v.rowb);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("%-24s N=%3d: %.1f cycles / MMA\n", v.name, N, (double)c / iters);
    }
  return 0;
}
