// Experiment: 3x3 stride-1 convolution tile (16 x 8 output pixels) computed by tcgen05.mma straight
// from a TMA-loaded halo patch (18 x 10 pixels, OOB zero-filled = padding), one shifted A descriptor
// per tap.  Checks C = 64 (SWIZZLE_128B), 32 (SWIZZLE_64B), 16 (SWIZZLE_32B) and 8 (no swizzle, two
// taps per K=16 MMA through LBO) against a CPU convolution.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

static constexpr int PW = 10, PH = 18, BN = 16;
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// W packed [BN][KP] bf16 K-major, k = tap * C + c, KP = 9C rounded up to 64
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tmA, const __nv_bfloat16* Wp, float* D,
                                        int C, int KP, int ox0, int oy0, int a_layout) {
  extern __shared__ uint8_t raw[];
  const uint32_t raw_addr = smem_u32(raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - raw_addr);
  const int row_bytes = C * 2;
  uint8_t* a_s = sm;  // PH*PW rows (+ slack)
  uint8_t* b_s = sm + 32768;  // KP/64 blocks of [BN][64] swizzled 128B
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nkb = KP / 64;
  for (int i = tid; i < nkb * BN * 8; i += 128) {
    int kb = i / (BN * 8), r = (i / 8) % BN, c = i & 7;
    *reinterpret_cast<uint4*>(b_s + kb * BN * 128 + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(Wp + (size_t)r * KP + kb * 64 + c * 8);
  }
  // zero the slack after the patch (C = 8 reads one pixel past the last tap)
  for (int i = tid; i < 64; i += 128) *reinterpret_cast<uint4*>(a_s + PH * PW * row_bytes + i * 16) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
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
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar2)), "r"((uint32_t)(PH * PW * row_bytes)) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(base),
                 "l"((uint64_t)&tmA), "r"(0), "r"(ox0 - 1), "r"(oy0 - 1), "r"(0), "r"(smem_u32(&bar2)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t hi_b = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61) | ((uint64_t)1 << 16);
    const uint32_t b_start = smem_u32(b_s);
    const uint32_t sbo = (uint32_t)(PW * row_bytes);
    int first = 1;
    if (C >= 16) {
      const uint64_t hi_a = ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)a_layout << 61) | ((uint64_t)1 << 16);
      for (int tap = 0; tap < 9; tap++) {
        const int dy = tap / 3, dx = tap % 3;
        const uint32_t a_tap = base + (uint32_t)((dy * PW + dx) * row_bytes);
        for (int kk = 0; kk < C / 16; kk++) {
          const int kglob = tap * C + kk * 16;
          uint64_t da = hi_a | (uint64_t)(((a_tap + kk * 32) >> 4) & 0x3FFF);
          uint64_t db = hi_b | (uint64_t)(((b_start + (kglob / 64) * BN * 128 + (kglob % 64) * 2) >> 4) & 0x3FFF);
          uint32_t accum = !first;
          first = 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
        }
      }
    } else {
      // C = 8: rows are 16 B; K = 16 = taps (t, t+1); LBO = byte distance between the two taps' pixels
      for (int j = 0; j < 5; j++) {
        const int t0 = 2 * j, t1 = 2 * j + 1 < 9 ? 2 * j + 1 : 2 * j;  // 10th half: weights are zero
        const uint32_t a0 = base + (uint32_t)(((t0 / 3) * PW + t0 % 3) * 16);
        const uint32_t a1 = base + (uint32_t)(((t1 / 3) * PW + t1 % 3) * 16);
        const uint32_t lbo = t1 == t0 ? 16u : a1 - a0;
        const uint64_t hi_a = ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)0 << 61) | ((uint64_t)(lbo >> 4) << 16);
        const int kglob = j * 16;
        uint64_t da = hi_a | (uint64_t)((a0 >> 4) & 0x3FFF);
        uint64_t db = hi_b | (uint64_t)(((b_start + (kglob / 64) * BN * 128 + (kglob % 64) * 2) >> 4) & 0x3FFF);
        uint32_t accum = !first;
        first = 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
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
  PFN_encodeTiled enc = nullptr;
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no encode\n"); return 1; }
    enc = (PFN_encodeTiled)p;
  }
  const int H = 24, W = 12;
  int Cs[] = {64, 32, 16, 8};
  for (int C : Cs) {
    const int KP = (9 * C + 63) / 64 * 64;
    std::vector<__nv_bfloat16> hX((size_t)H * W * C), hW((size_t)BN * KP);
    std::vector<float> fX(hX.size()), fW(hW.size(), 0.f);
    srand(C);
    for (size_t i = 0; i < hX.size(); i++) { float x = (float)(rand() % 17 - 8) / 8.f; hX[i] = __float2bfloat16(x); fX[i] = x; }
    for (size_t i = 0; i < hW.size(); i++) hW[i] = __float2bfloat16(0.f);
    for (int n = 0; n < BN; n++)
      for (int kk = 0; kk < 9 * C; kk++) { float x = (float)(rand() % 13 - 6) / 4.f; hW[(size_t)n * KP + kk] = __float2bfloat16(x); fW[(size_t)n * KP + kk] = x; }
    __nv_bfloat16 *dX, *dW; float* dD;
    cudaMalloc(&dX, hX.size() * 2); cudaMalloc(&dW, hW.size() * 2); cudaMalloc(&dD, 128 * 16 * 4);
    cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, PW, PH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : C == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    int a_layout = C == 64 ? 2 : C == 32 ? 4 : C == 16 ? 6 : 0;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("C=%d encode failed %d\n", C, (int)r); continue; }
    size_t smem = 1024 + 32768 + (size_t)(KP / 64) * BN * 128 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int tiles[2][2] = {{0, 0}, {8, 16}};
    for (auto& t : tiles) {
      const int ox0 = t[0], oy0 = t[1];
      cudaMemset(dD, 0, 128 * 16 * 4);
      k<<<1, 128, smem>>>(tm, dW, dD, C, KP, ox0, oy0, a_layout);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("C=%d: CUDA error %s\n", C, cudaGetErrorString(e)); return 1; }
      std::vector<float> hD(128 * 16);
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, checked = 0; double maxerr = 0;
      for (int m = 0; m < 128; m++) {
        const int oy = oy0 + m / 8, ox = ox0 + m % 8;
        if (oy >= H || ox >= W) continue;
        for (int n = 0; n < BN; n++) {
          float ref = 0;
          for (int tap = 0; tap < 9; tap++) {
            const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
            if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
            for (int c = 0; c < C; c++) ref += fX[((size_t)iy * W + ix) * C + c] * fW[(size_t)n * KP + tap * C + c];
          }
          double err = fabs(ref - hD[m * 16 + n]);
          checked++;
          if (err > 1e-2) bad++;
          if (err > maxerr) maxerr = err;
        }
      }
      printf("C=%2d tile (ox0 %d, oy0 %d): %s (bad %d / %d, max err %.4f)\n", C, ox0, oy0, bad ? "MISMATCH" : "OK", bad, checked, maxerr);
    }
    cudaFree(dX); cudaFree(dW); cudaFree(dD);
  }
  return 0;
}
