// Memory-bound kernels of the YOLOv11 forward: stem conv, depthwise 3x3, SPPF pooling, the C2PSA
// attention core and the DFL box decode.  All activations are NHWC fp16 or bf16 (template parameter F16,
// plan->act_f16); arithmetic is fp32.
#include <math.h>
#include <algorithm>
#include <stdio.h>
#include <string.h>

#include <cuda_fp16.h>

#include "yb_internal.h"

namespace yb {

__device__ __forceinline__ float silu_half(float h) {  // SiLU(2h) = h + h*tanh(h)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float silu_acc(float x) {  // h + h*tanh(h), h = x/2 (one MUFU op)
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ float ex2_approx(float x) {   // 2^x, one MUFU op: -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Blackwell packed fp32 FMA: two independent fused multiply-adds per instruction (FFMA2).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <bool F16>
__device__ __forceinline__ void act8_to_float(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float2 t = Act16<F16>::unpack2(w[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
template <bool F16>
__device__ __forceinline__ uint4 float_to_act8(const float* f) {
  return make_uint4(Act16<F16>::pack2(f[0], f[1]), Act16<F16>::pack2(f[2], f[3]), Act16<F16>::pack2(f[4], f[5]),
                    Act16<F16>::pack2(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------
// Stem: net.p1.0 = Conv(3 -> w1, k3, s2, p1) + SiLU (nets/nn.py:161).  Reads the caller's NCHW
// image (fp32 / fp16 / bf16 / uint8), writes NHWC bf16.  K = 27 is far below the tensor-core
// ridge (AI ~ 23 flop/B), so this is a direct convolution: a CTA stages the (2*8+1) x (2*32+1) x 3
// input patch of an 8 x 32 output tile in shared memory with coalesced loads, then one thread per
// output pixel keeps its 27 taps in registers and reads the weights as shared-memory broadcasts.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_px(const T* p);
template <>
__device__ __forceinline__ float load_px<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_px<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <>
__device__ __forceinline__ float load_px<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
template <>
__device__ __forceinline__ float load_px<uint8_t>(const uint8_t* p) { return (float)__ldg(p); }

static constexpr int STEM_TW = 64, STEM_TH = 8;              // output tile per CTA (256 threads x 2 pixels)
static constexpr int STEM_IH = 2 * STEM_TH + 1;
static constexpr int STEM_HP = STEM_TW + 16 + 4;             // half-row pitch (even or odd columns), floats

// unpack one 16-byte vector of pixels into its even-column and odd-column halves
template <typename T>
__device__ __forceinline__ void unpack_eo(const uint4& v, float* ev, float* od, float scale);
template <>
__device__ __forceinline__ void unpack_eo<float>(const uint4& v, float* ev, float* od, float scale) {
  ev[0] = __uint_as_float(v.x) * scale;
  od[0] = __uint_as_float(v.y) * scale;
  ev[1] = __uint_as_float(v.z) * scale;
  od[1] = __uint_as_float(v.w) * scale;
}
template <>
__device__ __forceinline__ void unpack_eo<__half>(const uint4& v, float* ev, float* od, float scale) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float2 t = __half22float2(h[j]);
    ev[j] = t.x * scale;
    od[j] = t.y * scale;
  }
}
template <>
__device__ __forceinline__ void unpack_eo<__nv_bfloat16>(const uint4& v, float* ev, float* od, float scale) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float2 t = __bfloat1622float2(h[j]);
    ev[j] = t.x * scale;
    od[j] = t.y * scale;
  }
}
template <>
__device__ __forceinline__ void unpack_eo<uint8_t>(const uint4& v, float* ev, float* od, float scale) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 8; j++) {
    ev[j] = (float)((w[j >> 1] >> (16 * (j & 1))) & 0xFFu) * scale;
    od[j] = (float)((w[j >> 1] >> (16 * (j & 1) + 8)) & 0xFFu) * scale;
  }
}

template <typename T, bool F16>
__global__ void __launch_bounds__(256)
    stem_conv_kernel(const T* __restrict__ in, act_t* __restrict__ out,
                     const float* __restrict__ wgt, int B, int H, int W, int Ho, int Wo, int Cp,
                     int out_ld, float in_scale) {
  constexpr int EPV = 16 / (int)sizeof(T);  // elements per 16-byte vector (W % 32 == 0 keeps rows aligned)
  constexpr int HV = EPV / 2;               // even (or odd) columns per vector
  // shared: [27][Cp] weights, [Cp] bias, then the input patch split into even and odd columns
  // [3][IH][HP] each: with stride 2 a warp's taps become unit-stride, conflict-free reads
  extern __shared__ float ws[];
  pdl_prologue_done();
  pdl_wait();
  float* te = ws + 28 * Cp;
  float* to = te + 3 * STEM_IH * STEM_HP;
  const int tiles_x = (Wo + STEM_TW - 1) / STEM_TW, tiles_y = (Ho + STEM_TH - 1) / STEM_TH;
  int bid = blockIdx.x;
  const int tx = bid % tiles_x;
  bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  for (int i = threadIdx.x; i < 28 * Cp; i += blockDim.x) ws[i] = wgt[i];
  // aligned window of the input rows: starts at gxa = 2*tx*TW - EPV (a multiple of EPV)
  const int gy0 = 2 * ty * STEM_TH - 1;
  const int gxa = 2 * tx * STEM_TW - EPV;
  constexpr int NVEC = (2 * STEM_TW + 1 + EPV + EPV - 1) / EPV;  // covers [gxa, gxa + EPV + 2*TW + 1)
  for (int i = threadIdx.x; i < 3 * STEM_IH * NVEC; i += blockDim.x) {
    const int vx = i % NVEC;
    const int row = i / NVEC;  // ci * IH + iy
    const int iy = row % STEM_IH, ci = row / STEM_IH;
    const int gy = gy0 + iy, gx = gxa + vx * EPV;
    float ev[HV], od[HV];
    if ((unsigned)gy < (unsigned)H && gx >= 0 && gx < W) {
      uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * 3 + ci) * H + gy) * W + gx));
      unpack_eo<T>(v, ev, od, in_scale);
    } else {
#pragma unroll
      for (int j = 0; j < HV; j++) ev[j] = od[j] = 0.f;
    }
    float* pe = te + row * STEM_HP + vx * HV;
    float* po = to + row * STEM_HP + vx * HV;
    if (HV % 4 == 0) {
#pragma unroll
      for (int j = 0; j < HV; j += 4) {
        *reinterpret_cast<float4*>(pe + j) = make_float4(ev[j], ev[j + 1], ev[j + 2], ev[j + 3]);
        *reinterpret_cast<float4*>(po + j) = make_float4(od[j], od[j + 1], od[j + 2], od[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < HV; j++) {
        pe[j] = ev[j];
        po[j] = od[j];
      }
    }
  }
  __syncthreads();
  const float* bs = ws + 27 * Cp;
  // one thread = output pixels lx and lx + 32 of a tile row: every weight read from shared memory
  // (a broadcast) feeds two packed FFMA2s, and lanes read consecutive even / odd columns
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int ox = tx * STEM_TW + lx, oy = ty * STEM_TH + ly;
  if (ox >= Wo || oy >= Ho) return;
  float x0[27], x1[27];
#pragma unroll
  for (int ci = 0; ci < 3; ci++)
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      // input column 2*ox + kx - 1: local index EPV + 2*lx + kx - 1 -> odd[e-1], even[e], odd[e], e = HV + lx
      const float* re = te + (ci * STEM_IH + 2 * ly + ky) * STEM_HP + HV + lx;
      const float* ro = to + (ci * STEM_IH + 2 * ly + ky) * STEM_HP + HV + lx;
      const int t = (ci * 3 + ky) * 3;
      x0[t] = ro[-1];
      x0[t + 1] = re[0];
      x0[t + 2] = ro[0];
      x1[t] = ro[31];
      x1[t + 1] = re[32];
      x1[t + 2] = ro[32];
    }
  act_t* op = out + (((size_t)b * Ho + oy) * Wo + ox) * out_ld;
  const bool two = ox + 32 < Wo;
  for (int c0 = 0; c0 < Cp; c0 += 8) {
    float2 a0[4], a1[4];
#pragma unroll
    for (int j = 0; j < 4; j++) a0[j] = a1[j] = make_float2(bs[c0 + 2 * j], bs[c0 + 2 * j + 1]);
#pragma unroll
    for (int t = 0; t < 27; t++) {
      const float4 w0 = *reinterpret_cast<const float4*>(ws + t * Cp + c0);
      const float4 w1 = *reinterpret_cast<const float4*>(ws + t * Cp + c0 + 4);
      const float2 wa = make_float2(w0.x, w0.y), wb = make_float2(w0.z, w0.w);
      const float2 wc = make_float2(w1.x, w1.y), wd = make_float2(w1.z, w1.w);
      const float2 p0 = make_float2(x0[t], x0[t]), p1 = make_float2(x1[t], x1[t]);
      a0[0] = ffma2(p0, wa, a0[0]);
      a0[1] = ffma2(p0, wb, a0[1]);
      a0[2] = ffma2(p0, wc, a0[2]);
      a0[3] = ffma2(p0, wd, a0[3]);
      a1[0] = ffma2(p1, wa, a1[0]);
      a1[1] = ffma2(p1, wb, a1[1]);
      a1[2] = ffma2(p1, wc, a1[2]);
      a1[3] = ffma2(p1, wd, a1[3]);
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      o[2 * j] = silu_acc(a0[j].x);
      o[2 * j + 1] = silu_acc(a0[j].y);
    }
    *reinterpret_cast<uint4*>(op + c0) = float_to_act8<F16>(o);
    if (two) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        o[2 * j] = silu_acc(a1[j].x);
        o[2 * j + 1] = silu_acc(a1[j].y);
      }
      *reinterpret_cast<uint4*>(op + (size_t)32 * out_ld + c0) = float_to_act8<F16>(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stem on the tensor cores (default).  The 3x3x3 receptive field is a K = 27 (padded to 32) GEMM row:
// a CTA stages the input patch of an 8 x 64 output tile in shared memory as bf16 (uint8 pixels are
// exact in bf16; the 1/255 scale is folded into the weights), every warp owns one tile row and builds
// its mma.sync.m16n8k16 A fragments straight from the patch (one 16-bit shared load per element,
// addresses = pixel base + a per-thread tap offset), the weights live in registers as B fragments.
// Against the direct version this trades 27 x Cp FMAs per pixel for 16 loads + 2 x Cp/8 MMAs per
// 16 pixels; the output tile goes through a per-warp staging buffer so that global stores are 16 B
// per lane and contiguous.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float vec_elem(const uint4& v, int j);   // element j of a 16-byte vector as fp32
template <>
__device__ __forceinline__ float vec_elem<float>(const uint4& v, int j) {
  return __uint_as_float(j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w);
}
template <>
__device__ __forceinline__ float vec_elem<__half>(const uint4& v, int j) {
  const uint32_t w = (j >> 1) == 0 ? v.x : (j >> 1) == 1 ? v.y : (j >> 1) == 2 ? v.z : v.w;
  return __half2float(__ushort_as_half((unsigned short)((w >> (16 * (j & 1))) & 0xFFFFu)));
}
template <>
__device__ __forceinline__ float vec_elem<__nv_bfloat16>(const uint4& v, int j) {
  const uint32_t w = (j >> 1) == 0 ? v.x : (j >> 1) == 1 ? v.y : (j >> 1) == 2 ? v.z : v.w;
  return __uint_as_float((j & 1) ? (w & 0xFFFF0000u) : (w << 16));
}
template <>
__device__ __forceinline__ float vec_elem<uint8_t>(const uint4& v, int j) {
  const uint32_t w = (j >> 2) == 0 ? v.x : (j >> 2) == 1 ? v.y : (j >> 2) == 2 ? v.z : v.w;
  return (float)((w >> (8 * (j & 3))) & 0xFFu);
}

// Shared-memory layout of the stem patch: three "tap column" planes per input channel.  The conv has
// stride 2, so output column x reads input columns 2x-1, 2x, 2x+1; plane kx holds input column
// 2x + kx - 1 at element x, which makes the 8 output pixels of an MMA row block contiguous (16 bytes) for
// every tap - exactly one row of an ldmatrix 8x8 tile.  ldmatrix.trans of [tap][pixel] tiles yields the
// [pixel][tap] A fragments of mma.m16n8k16 (two ldmatrix.x4 per 16 pixels x 32 taps instead of 32
// 16-bit loads).  Row pitch 144 B and plane stride 459 x 16 B spread the 8 taps of a tile over the banks.
static constexpr int STEM_PWB = (STEM_TW + 8) * 2;                  // plane row pitch, bytes
static constexpr int STEM_PLANE_B = 3 * STEM_IH * STEM_PWB;         // one kx plane (3 channels), bytes
static constexpr int STEM_PART_B = 3 * STEM_PLANE_B;                // all planes of the hi (or lo) part
static constexpr int STEM_ZERO_B = 128;
#ifndef YB_STEM_MI_UNROLL
#define YB_STEM_MI_UNROLL 1
#endif
static constexpr int STEM_MI_UNROLL = YB_STEM_MI_UNROLL;                             // zero rows for the K padding (k = 27..31)

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// NT: n-tiles (8 channels each) per pass over the tile; wider stems take several passes.
// SPLIT: the input is not exactly representable in the 16-bit operand type (fp32 images; fp16 images with bf16
// operands and vice versa): input and weights are staged as hi + lo pairs and multiplied as
// hi*W_hi + hi*W_lo + lo*W_hi, which matches an fp32 convolution to ~2^-16 relative.  uint8 images (and images
// already in the operand type) are exact and use 16-bit weights like every other layer of the network.
// F16: operand / output type (Act16).  uint8 pixels with fp16 operands are staged as v * 2^-8 (exact) and the
// weights carry 256/255 instead of 1/255, which keeps them clear of the fp16 subnormal range.
template <typename T, int NT, bool SPLIT, bool F16, int MINB = 1>
__global__ void __launch_bounds__(256, MINB)
    stem_mma_kernel(const T* __restrict__ in, act_t* __restrict__ out, const float* __restrict__ wgt,
                    int B, int H, int W, int Ho, int Wo, int Cp, int out_ld, float in_scale, int total_tiles,
                    uint32_t tx_mul, uint32_t tx_shr, uint32_t ty_mul, uint32_t ty_shr, int s2d) {
  using A16 = Act16<F16>;
  constexpr int EPV = 16 / (int)sizeof(T);
  constexpr bool U8 = sizeof(T) == 1;
  constexpr float PXS = (U8 && F16) ? 1.f / 256.f : 1.f;   // pixel staging scale (a power of two)
  const float wscale = in_scale / PXS;
  // generic staging walks every 16-byte vector that overlaps patch columns [EPV-1, EPV-1 + 2*TW]
  constexpr int NVEC = (2 * STEM_TW + 1 + EPV + EPV - 1) / EPV;
  constexpr int PARTS = SPLIT ? 2 : 1;
  extern __shared__ __align__(16) uint8_t stem_smem[];
  uint8_t* zero_rows = stem_smem + PARTS * STEM_PART_B;
  act_t* stage = reinterpret_cast<act_t*>(zero_rows + STEM_ZERO_B);   // [8 warps][16 px][NT*8]
  pdl_prologue_done();
  const int tiles_x = (Wo + STEM_TW - 1) / STEM_TW, tiles_y = (Ho + STEM_TH - 1) / STEM_TH;
  // tile -> (image, tile row, tile column) by multiply-high with host-computed magic numbers: the two runtime
  // divisions (twice per tile and thread: prefetch and compute) were ~12 % of the kernel's instructions
  auto split_tile = [&](int tl, int& tx, int& ty, int& b) {
    const int q = tiles_x == 1 ? tl : (int)(__umulhi((uint32_t)tl, tx_mul) >> tx_shr);
    tx = tl - q * tiles_x;
    b = tiles_y == 1 ? q : (int)(__umulhi((uint32_t)q, ty_mul) >> ty_shr);
    ty = q - b * tiles_y;
  };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (tid < STEM_ZERO_B / 4) reinterpret_cast<uint32_t*>(zero_rows)[tid] = 0u;
  // ---- ldmatrix row address of this lane per k-step: tile (lane >> 3) = {px 0-7 | px 8-15} x {k 0-7 | k 8-15},
  // row (lane & 7) = tap k
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(stem_smem);
  uint32_t aoff[2];
#pragma unroll
  for (int ks = 0; ks < 2; ks++) {
    const int k = ks * 16 + ((lane >> 4) & 1) * 8 + (lane & 7);
    const int ci = k / 9, r9 = k - ci * 9, ky = r9 / 3, kx = r9 - ky * 3;
    aoff[ks] = k < 27 ? (uint32_t)(kx * STEM_PLANE_B + (ci * STEM_IH + 2 * warp + ky) * STEM_PWB + ((lane >> 3) & 1) * 16)
                      : (uint32_t)(PARTS * STEM_PART_B);
  }
  // ---- weights of a channel group -> B fragments (n = g per n-tile), hi + lo split.  The input scale
  // and the 1/2 of SiLU(x) = h + h*tanh(h), h = x/2, are folded into weights and bias (exact: power of 2
  // for the half; the scale is folded before the bf16 rounding exactly like the packed conv weights).
  const float* bias = wgt + 27 * Cp;
  uint32_t bhi[NT][2][2], blo[NT][2][2];
  float bia[NT][2];
  auto load_weights = [&](int c0) {
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      const int n = c0 + nt * 8 + g;
#pragma unroll
      for (int ks = 0; ks < 2; ks++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int k0 = ks * 16 + h * 8 + 2 * t;
          const float w0 = (k0 < 27 && n < Cp) ? 0.5f * (__ldg(wgt + k0 * Cp + n) * wscale) : 0.f;
          const float w1 = (k0 + 1 < 27 && n < Cp) ? 0.5f * (__ldg(wgt + (k0 + 1) * Cp + n) * wscale) : 0.f;
          const uint16_t h0 = A16::pack1(w0), h1 = A16::pack1(w1);
          bhi[nt][ks][h] = (uint32_t)h0 | ((uint32_t)h1 << 16);
          blo[nt][ks][h] = A16::pack2(w0 - A16::unpack1(h0), w1 - A16::unpack1(h1));
        }
      const int cb = c0 + nt * 8 + 2 * t;
      bia[nt][0] = cb < Cp ? 0.5f * __ldg(bias + cb) : 0.f;
      bia[nt][1] = cb + 1 < Cp ? 0.5f * __ldg(bias + cb + 1) : 0.f;
    }
  };
  const bool single = Cp <= NT * 8;   // one channel group: its fragments stay in registers for all tiles
  if (single) load_weights(0);
  act_t* wstage = stage + warp * 16 * NT * 8;
  // output store: lane -> (pixel, 16-byte channel chunk) of the warp's 16 x (NT*8) staging tile (NT = 2 / 4: 32 % NT == 0)
  const int st_px = lane / NT, st_cg = lane - st_px * NT;
  const int st_src = st_px * NT * 8 + st_cg * 8;
  const int st_off = (s2d ? 2 * st_px - (st_px & 1) : st_px) * out_ld + st_cg * 8;   // (32 / NT is even: the parity of px is the lane's)
  bool st_cok = true;

  // ---- uint8 staging: a thread owns 16-byte vectors (16 pixels of one input row) whose even / odd
  // pixels are whole 16-byte plane rows; the kx = 0 plane is the kx = 2 plane shifted by one pixel and
  // needs the byte in front of the vector.  Loads run one tile ahead (registers).
  constexpr int NV8 = STEM_TW / 8;                                  // vectors per patch row (uint8)
  constexpr int NIT = (3 * STEM_IH * NV8 + 255) / 256;
  uint4 pv[U8 ? NIT : 1];
  uint32_t pb[U8 ? NIT : 1];
  auto fetch_tile = [&](int tl) {
    int tx, ty, b;
    split_tile(tl, tx, ty, b);
    const int gy0 = 2 * ty * STEM_TH - 1, gx0 = 2 * tx * STEM_TW;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(in) + (size_t)b * 3 * (size_t)H * (size_t)W;
#pragma unroll
    for (int it = 0; it < (U8 ? NIT : 1); it++) {
      const int i = tid + it * 256;
      const int vx = i % NV8, row = i / NV8;   // row = ci * IH + iy
      const int ci = row / STEM_IH, iy = row - ci * STEM_IH;
      const int gy = gy0 + iy, gx = gx0 + vx * 16;
      pv[it] = make_uint4(0u, 0u, 0u, 0u);
      pb[it] = 0u;
      if (row < 3 * STEM_IH && (unsigned)gy < (unsigned)H && gx < W) {
        // (one 64-bit image base per tile; the offset inside a 3 x H x W image fits 32 bits)
        const uint8_t* src = img + (uint32_t)((ci * H + gy) * W + gx);
        pv[it] = __ldg(reinterpret_cast<const uint4*>(src));
        if (vx == 0 && gx > 0) pb[it] = __ldg(src - 1);   // first vector of the patch row: the byte left of the tile
      }
    }
  };
  auto store_tile_u8 = [&]() {
#pragma unroll
    for (int it = 0; it < (U8 ? NIT : 1); it++) {
      const int i = tid + it * 256;
      const int vx = i % NV8, row = i / NV8;
      // every vector but the first of a patch row takes its left neighbour from the previous lane (NV8 = 8 consecutive
      // lanes hold one patch row; a masked-out vector is zero, which is also the right value for it).  The shuffle sits
      // HERE, not behind the loads in fetch_tile: there it made every thread wait for its prefetch on the spot (ncu: 28 %
      // of the kernel's stall samples, long scoreboard) instead of one tile later.  All lanes take part, hence no break.
      const uint32_t left = __shfl_up_sync(0xffffffffu, pv[it].w >> 24, 1);
      if (row >= 3 * STEM_IH) continue;
      const uint32_t pbv = vx != 0 ? left : pb[it];
      const uint32_t w[4] = {pv[it].x, pv[it].y, pv[it].z, pv[it].w};
      // input column gx0 + 16 vx + j is patch column r = 16 vx + j + 1 (r = 2x + kx)
      uint4 p0, p1, p2;
      if constexpr (F16) {
        // bytes -> fp16 without integer conversions: 0x6400 | b is the fp16 number 1024 + b, and
        // (1024 + b) * 2^-8 - 4 = b / 256 exactly; one PRMT builds a pair, one HFMA2 rescales it
        const uint32_t sc = 0x1C001C00u, off = 0xC400C400u;   // half2(2^-8), half2(-4.0)
        uint32_t ev[4], od[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          uint32_t a, b;
          asm("prmt.b32 %0, %1, %2, 0x7270;" : "=r"(a) : "r"(w[k]), "r"(0x64646464u));   // pixels 4k, 4k+2
          asm("prmt.b32 %0, %1, %2, 0x7371;" : "=r"(b) : "r"(w[k]), "r"(0x64646464u));   // pixels 4k+1, 4k+3
          asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(ev[k]) : "r"(a), "r"(sc), "r"(off));
          asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(od[k]) : "r"(b), "r"(sc), "r"(off));
        }
        uint32_t hp;
        asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(hp) : "r"(0x64006400u | pbv), "r"(sc), "r"(off));
        uint32_t sh[4];   // odd pixels shifted by one: (prev, 1), (3, 5), (7, 9), (11, 13)
        asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(sh[0]) : "r"(hp), "r"(od[0]));
#pragma unroll
        for (int k = 1; k < 4; k++) asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(sh[k]) : "r"(od[k - 1]), "r"(od[k]));
        p0 = make_uint4(sh[0], sh[1], sh[2], sh[3]);
        p1 = make_uint4(ev[0], ev[1], ev[2], ev[3]);
        p2 = make_uint4(od[0], od[1], od[2], od[3]);
      } else {
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; j++) f[j] = (float)((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) * PXS;
      const float fp = (float)pbv * PXS;
      p1 = make_uint4(A16::pack2(f[0], f[2]), A16::pack2(f[4], f[6]), A16::pack2(f[8], f[10]), A16::pack2(f[12], f[14]));
      p2 = make_uint4(A16::pack2(f[1], f[3]), A16::pack2(f[5], f[7]), A16::pack2(f[9], f[11]), A16::pack2(f[13], f[15]));
      p0 = make_uint4(A16::pack2(fp, f[1]), A16::pack2(f[3], f[5]), A16::pack2(f[7], f[9]), A16::pack2(f[11], f[13]));
      }
      uint8_t* dst = stem_smem + row * STEM_PWB + vx * 16;
      *reinterpret_cast<uint4*>(dst) = p0;
      *reinterpret_cast<uint4*>(dst + STEM_PLANE_B) = p1;
      *reinterpret_cast<uint4*>(dst + 2 * STEM_PLANE_B) = p2;
    }
  };
  pdl_wait();
  if (U8 && (int)blockIdx.x < total_tiles) fetch_tile(blockIdx.x);
#pragma unroll 1
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    int tx, ty, b;
    split_tile(tile, tx, ty, b);
    if (tile != (int)blockIdx.x) __syncthreads();   // every warp is done with the previous patch
    if (U8) {
      store_tile_u8();
    } else {
      // ---- generic staging (bf16 / fp16 / fp32 images): element-wise plane stores, hi and lo parts
      const int gy0 = 2 * ty * STEM_TH - 1;
      const int gxa = 2 * tx * STEM_TW - EPV;   // patch column pc = EPV - 1 + r
      const T* tile_in = in + ((size_t)b * 3 * H + gy0) * (size_t)W + gxa;   // only in-bounds offsets are read
#pragma unroll 1
      for (int i = tid; i < 3 * STEM_IH * NVEC; i += 256) {
        const int vx = i % NVEC, row = i / NVEC;  // row = ci * IH + iy
        const int iy = row % STEM_IH, vxe = vx * EPV;
        const int gy = gy0 + iy, gx = gxa + vxe;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if ((unsigned)gy < (unsigned)H && gx >= 0 && gx < W)
          v = __ldg(reinterpret_cast<const uint4*>(tile_in + ((row / STEM_IH) * H + iy) * W + vxe));
        unsigned short* prow = reinterpret_cast<unsigned short*>(stem_smem + row * STEM_PWB);
#pragma unroll
        for (int j = 0; j < EPV; j++) {
          const float x = vec_elem<T>(v, j);
          const unsigned short uh = A16::pack1(x);
          const unsigned short ul = SPLIT ? A16::pack1(x - A16::unpack1(uh)) : (unsigned short)0;
          const int r = vxe + j - (EPV - 1);   // parity of r is a compile-time property of j
          if (r < 0 || r > 2 * STEM_TW) continue;
          if (r & 1) {
            prow[STEM_PLANE_B / 2 + (r >> 1)] = uh;
            if (SPLIT) prow[(STEM_PART_B + STEM_PLANE_B) / 2 + (r >> 1)] = ul;
          } else {
            if (r < 2 * STEM_TW) {
              prow[r >> 1] = uh;
              if (SPLIT) prow[STEM_PART_B / 2 + (r >> 1)] = ul;
            }
            if (r >= 2) {
              prow[STEM_PLANE_B + (r >> 1) - 1] = uh;
              if (SPLIT) prow[STEM_PART_B / 2 + STEM_PLANE_B + (r >> 1) - 1] = ul;
            }
          }
        }
      }
    }
    __syncthreads();
    // the next tile's pixels travel while this one is computed
    if (U8 && tile + (int)gridDim.x < total_tiles) fetch_tile(tile + (int)gridDim.x);
    const int oy = ty * STEM_TH + warp;
    if (oy >= Ho) continue;
#pragma unroll 1
    for (int c0 = 0; c0 < Cp; c0 += NT * 8) {
      if (!single) load_weights(c0);
      st_cok = c0 + st_cg * 8 < Cp;
      // base of this warp's output row in the tile (one 64-bit address per tile and channel pass, 32-bit steps per block)
      const int tx0 = tx * STEM_TW;
      act_t* orow0 = s2d ? out + ((((size_t)b * (Ho >> 1) + (oy >> 1)) * (Wo >> 1) + (tx0 >> 1)) * 4 + (oy & 1) * 2) * out_ld + c0
                         : out + (((size_t)b * Ho + oy) * Wo + tx0) * out_ld + c0;
      const int opitch = s2d ? 2 * out_ld : out_ld;   // elements per output pixel step along x
#pragma unroll(STEM_MI_UNROLL)
      for (int mi = 0; mi < STEM_TW / 16; mi++) {
        uint32_t afr[PARTS][2][4];
#pragma unroll
        for (int part = 0; part < PARTS; part++)
#pragma unroll
          for (int ks = 0; ks < 2; ks++)
            ldmatrix_x4_trans(afr[part][ks], smem_base + aoff[ks] + (uint32_t)(mi * 32) +
                                                 ((part && aoff[ks] < (uint32_t)STEM_PART_B) ? (uint32_t)STEM_PART_B : 0u));
        float d[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          d[nt][0] = d[nt][2] = bia[nt][0];
          d[nt][1] = d[nt][3] = bia[nt][1];
#pragma unroll
          for (int ks = 0; ks < 2; ks++) {
            if (SPLIT) {
              A16::mma16816(d[nt], afr[0][ks], blo[nt][ks][0], blo[nt][ks][1]);
              A16::mma16816(d[nt], afr[SPLIT ? 1 : 0][ks], bhi[nt][ks][0], bhi[nt][ks][1]);
            }
            A16::mma16816(d[nt], afr[0][ks], bhi[nt][ks][0], bhi[nt][ks][1]);
          }
        }
        // SiLU (accumulators hold x/2) -> bf16 -> per-warp staging [16 px][NT*8] -> 16-byte global stores
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          *reinterpret_cast<uint32_t*>(wstage + g * NT * 8 + nt * 8 + 2 * t) = A16::pack2(silu_half(d[nt][0]), silu_half(d[nt][1]));
          *reinterpret_cast<uint32_t*>(wstage + (g + 8) * NT * 8 + nt * 8 + 2 * t) = A16::pack2(silu_half(d[nt][2]), silu_half(d[nt][3]));
        }
        const int ox0 = tx * STEM_TW + mi * 16;
        __syncwarp();
        // one 64-bit base per 16-pixel block; space-to-depth destination (plan.cu): pixel (y, x) is channel block
        // (y&1)*2 + (x&1) of pixel (y/2, x/2) of a (Ho/2, Wo/2, 4*out_ld) tensor - a row of 16 pixels lands as 64-byte
        // halves of eight 128-byte pixels (+33 us on the stem against -104 us on net.p2.0; pairing the two warps of a
        // row pair through named barriers to store whole 512-byte runs was measured slower: +75 us; so was giving each warp a
        // row pair and half the columns, +57 us: 64 bytes of spills at the kernel's 64 registers)
        act_t* orow = orow0 + mi * 16 * opitch;
        if constexpr (NT == 2 || NT == 4) {
          // 16-byte chunks of the 16 x (NT*8) tile: a lane's channel group and its first pixel never change, so the
          // offsets are set up once per kernel (this store was 12 % of the kernel's instructions)
#pragma unroll
          for (int k = 0; k < 16 * NT / 32; k++) {
            const int px = st_px + k * (32 / NT);
            if (ox0 + px < Wo && st_cok)
              *reinterpret_cast<uint4*>(orow + (s2d ? st_off + k * (32 / NT) * 2 * out_ld : st_off + k * (32 / NT) * out_ld)) =
                  *reinterpret_cast<const uint4*>(wstage + st_src + k * 32 * 8);
          }
        } else {
        for (int c = lane; c < 16 * NT; c += 32) {   // 16-byte chunks of the 16 x (NT*8) tile
          const int px = c / NT, cg = c - px * NT;
          const int pxo = s2d ? 2 * px - (px & 1) : px;
          if (ox0 + px < Wo && c0 + cg * 8 < Cp)
            *reinterpret_cast<uint4*>(orow + pxo * out_ld + cg * 8) =
                *reinterpret_cast<const uint4*>(wstage + px * NT * 8 + cg * 8);
        }
        }
      }
    }
  }
}

static int stem_ctas_per_sm() {
  static const int v = getenv("YB_STEM_CTAS") ? atoi(getenv("YB_STEM_CTAS")) : 8;
  return v;
}

template <typename T, bool SPLIT, bool F16>
static int launch_stem_t(const yb_plan* p, const Op& op, const void* in, float scale, cudaStream_t st) {
  const ConvW& cw = p->convs[op.conv_index];
  const float* w = reinterpret_cast<const float*>(p->d_weights + cw.info.blob_offset);
  const Buf& db = p->bufs[op.dst.buf];
  act_t* out = reinterpret_cast<act_t*>(buf_ptr(p, op.dst.buf)) + op.dst.c_off;
  const int Cp = cpad8(op.dst.C);
  const unsigned blocks = (unsigned)(p->B * ((op.Hout + STEM_TH - 1) / STEM_TH) * ((op.Wout + STEM_TW - 1) / STEM_TW));
  static const bool direct = getenv("YB_STEM_DIRECT") != nullptr;
  if (direct) {  // CUDA-core version (cross-check)
    const size_t smem = ((size_t)28 * Cp + 2 * 3 * STEM_IH * STEM_HP) * 4;
    YB_CUDA(launch_pdl(stem_conv_kernel<T, F16>, dim3(blocks), dim3(256), smem, st, (const T*)in, out, w, p->B, p->H, p->W,
                       op.Hout, op.Wout, Cp, db.C, scale));
    return YB_OK;
  }
  // n-tiles (8 channels) per pass.  A pass writes nt * 16 bytes of every pixel: 4 n-tiles = 64-byte pieces (whole
  // sectors) wherever the channel count allows (YOLO11x: 96 channels = 3 passes of 32, not 4 passes of 24 whose
  // 48-byte pieces straddle sectors); 3 n-tiles only when that is the whole pixel (YOLO11t: 24 channels)
  const int nt = (Cp / 8) % 4 == 0 ? 4 : ((Cp / 8) % 3 == 0 ? 3 : 2);
  const size_t smem = (size_t)(SPLIT ? 2 : 1) * STEM_PART_B + STEM_ZERO_B + (size_t)8 * 16 * nt * 8 * 2;
  // persistent (grid = resident CTAs): the tap offsets and, with one channel group, the weight fragments
  // are set up once per CTA instead of once per tile
#define YB_STEM_MMA(NT, MINB)                                                                                    \
  do {                                                                                                           \
    static int occ_dev[YB_MAX_DEVICES] = {0};   /* the attribute and the occupancy are per device */           \
    int& occ = occ_dev[p->device & (YB_MAX_DEVICES - 1)];                                                        \
    if (occ == 0) {                                                                                              \
      YB_CUDA(cudaFuncSetAttribute(stem_mma_kernel<T, NT, SPLIT, F16, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   (int)(2 * STEM_PART_B + STEM_ZERO_B + 8 * 16 * 4 * 8 * 2)));                  \
      YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stem_mma_kernel<T, NT, SPLIT, F16, MINB>, 256, smem)); \
      occ = std::max(1, std::min(occ, stem_ctas_per_sm()));                                                      \
    }                                                                                                            \
    const unsigned grid = std::min(blocks, (unsigned)(p->num_sms * occ));                                        \
    YB_CUDA(launch_pdl(stem_mma_kernel<T, NT, SPLIT, F16, MINB>, dim3(grid), dim3(256), smem, st, (const T*)in, out, w, \
                       p->B, p->H, p->W, op.Hout, op.Wout, Cp, db.C, scale, (int)blocks, mx.mul, mx.shr, my.mul, my.shr, db.s2d)); \
  } while (0)
  // q = n / d for n < 2^31: mul = ceil(2^(31 + ceil_log2 d) / d), shr = ceil_log2 d - 1 (as in conv_tc.cu)
  struct Magic { uint32_t mul = 0, shr = 0; };
  auto magic = [](int d) {
    Magic m;
    if (d <= 1) return m;
    int lg = 0;
    while ((1u << lg) < (uint32_t)d) lg++;
    const int pshift = 31 + lg;
    m.mul = (uint32_t)((((unsigned long long)1 << pshift) + (unsigned)d - 1) / (unsigned)d);
    m.shr = (uint32_t)(pshift - 32);
    return m;
  };
  const Magic mx = magic((op.Wout + STEM_TW - 1) / STEM_TW), my = magic((op.Hout + STEM_TH - 1) / STEM_TH);
  static const bool minb4 = getenv("YB_STEM_MINB1") == nullptr;
  if (nt == 3) YB_STEM_MMA(3, 1);
  else if (nt == 4) YB_STEM_MMA(4, 1);
  else if (minb4) YB_STEM_MMA(2, 4);
  else YB_STEM_MMA(2, 1);
#undef YB_STEM_MMA
  return YB_OK;
}

int launch_stem(const yb_plan* p, const Op& op, const void* in, int in_dtype, cudaStream_t st) {
  int rc;
  const bool f16 = p->act_f16 != 0;
  switch (in_dtype) {
    case YB_F32: rc = f16 ? launch_stem_t<float, true, true>(p, op, in, 1.f, st) : launch_stem_t<float, true, false>(p, op, in, 1.f, st); break;
    case YB_F16: rc = f16 ? launch_stem_t<__half, false, true>(p, op, in, 1.f, st) : launch_stem_t<__half, true, false>(p, op, in, 1.f, st); break;
    case YB_BF16: rc = f16 ? launch_stem_t<__nv_bfloat16, true, true>(p, op, in, 1.f, st) : launch_stem_t<__nv_bfloat16, false, false>(p, op, in, 1.f, st); break;
    case YB_U8: rc = f16 ? launch_stem_t<uint8_t, false, true>(p, op, in, 1.f / 255.f, st) : launch_stem_t<uint8_t, false, false>(p, op, in, 1.f / 255.f, st); break;
    default:
      set_error("unsupported input dtype %d", in_dtype);
      return YB_ERR_ARG;
  }
  if (rc) return rc;
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

// ---------------------------------------------------------------------------------------------
// Depthwise 3x3, stride 1, pad 1 (+ folded BN bias, optional SiLU): head cls branches
// (nets/nn.py:248,250) and the attention positional conv `pe` on v (nn.py:109,122).
// One thread = 4 pixels along x by 8 channels (16 B each): 18 tap loads feed 4 outputs.  `gsz/gstride/goff` gather the
// source channels (v rows of the per-head [q k v] interleave); `add` accumulates into dst.
// ---------------------------------------------------------------------------------------------
// (Measured in round 2 on YOLO11x's unfused 384-channel head convs, 540 us each at 80 x 80, B = 128 = 2.3 TB/s; ncu: L1/TEX 75 %
// busy, DRAM 31 %.  A row-streaming variant - three-row register window, weights in registers, one new row of six loads per
// four outputs - cut the L1 traffic to 19 % but ran at 8 warps per SM with 255 registers (558 us), and with four channels
// per thread at 16 warps and 88 bytes of spills (547 us): no gain either way, dropped.)
static constexpr int DW_PX = 4;  // output pixels along x per thread: 18 tap loads for 4 outputs

template <bool F16>
__global__ void __launch_bounds__(256, 3)
    dwconv3x3_kernel(const act_t* __restrict__ src, int src_ld, act_t* dst, int dst_ld,
                     const float* __restrict__ wgt, int Cp, int B, int H, int W, int C, int gsz,
                     int gstride, int goff, int act, int add) {
  pdl_prologue_done();
  pdl_wait();
  // grid = (ceil(xq * cgs / 256), H, B): one 32-bit division per thread instead of 64-bit index math
  const int cgs = C >> 3;
  const int xq = (W + DW_PX - 1) / DW_PX;
  const int tix = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (tix >= xq * cgs) return;
  const int xqi = tix / cgs;
  const int cg = tix - xqi * cgs;
  const int x0 = xqi * DW_PX;
  const int y = (int)blockIdx.y;
  const int b = (int)blockIdx.z;
  const int c = cg * 8;
  const int sc = (c / gsz) * gstride + goff + (c % gsz);
  float2 acc[DW_PX][4];
  const float* bias = wgt + 9 * Cp;
#pragma unroll
  for (int p = 0; p < DW_PX; p++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[p][j] = make_float2(bias[c + 2 * j], bias[c + 2 * j + 1]);
  // All 18 tap loads are predicated (no branches), so they issue back to back and the thread waits
  // once for the whole 3 x 6 window instead of once per tap.
  uint4 v[3][DW_PX + 2];
#pragma unroll
  for (int ky = 0; ky < 3; ky++) {
    const int iy = y - 1 + ky;
    const bool row_ok = (unsigned)iy < (unsigned)H;
    // (one 64-bit address per row, then a running pointer: ncu showed ~8 instructions per load on this line)
    const act_t* pc = src + (((size_t)(b * H + (row_ok ? iy : y)) * W) + (size_t)x0) * src_ld + sc - src_ld;
#pragma unroll
    for (int col = 0; col < DW_PX + 2; col++) {
      const int ix = x0 - 1 + col;
      v[ky][col] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok && (unsigned)ix < (unsigned)W) v[ky][col] = __ldg(reinterpret_cast<const uint4*>(pc));
      pc += src_ld;
    }
  }
#pragma unroll
  for (int ky = 0; ky < 3; ky++) {
    float2 w[3][4];
#pragma unroll
    for (int kx = 0; kx < 3; kx++) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wgt + (ky * 3 + kx) * Cp + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wgt + (ky * 3 + kx) * Cp + c + 4));
      w[kx][0] = make_float2(w0.x, w0.y);
      w[kx][1] = make_float2(w0.z, w0.w);
      w[kx][2] = make_float2(w1.x, w1.y);
      w[kx][3] = make_float2(w1.z, w1.w);
    }
#pragma unroll
    for (int col = 0; col < DW_PX + 2; col++) {
      const uint32_t h2[4] = {v[ky][col].x, v[ky][col].y, v[ky][col].z, v[ky][col].w};
      float2 f[4];
#pragma unroll
      for (int j = 0; j < 4; j++) f[j] = Act16<F16>::unpack2(h2[j]);
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int p = col - kx;  // output pixel this column feeds through tap kx
        if (p >= 0 && p < DW_PX) {
#pragma unroll
          for (int j = 0; j < 4; j++) acc[p][j] = ffma2(f[j], w[kx][j], acc[p][j]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < DW_PX; p++) {
    const int x = x0 + p;
    if (x >= W) break;
    float o[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      o[2 * j] = act ? silu_acc(acc[p][j].x) : acc[p][j].x;
      o[2 * j + 1] = act ? silu_acc(acc[p][j].y) : acc[p][j].y;
    }
    act_t* dp = dst + ((size_t)(b * H + y) * W + x) * dst_ld + c;
    if (add) {
      float f[8];
      act8_to_float<F16>(*reinterpret_cast<const uint4*>(dp), f);
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] += f[j];
    }
    *reinterpret_cast<uint4*>(dp) = float_to_act8<F16>(o);
  }
}

int launch_dw(const yb_plan* p, const Op& op, cudaStream_t st) {
  const ConvW& cw = p->convs[op.conv_index];
  const float* w = reinterpret_cast<const float*>(p->d_weights + cw.info.blob_offset);
  const Buf& sb = p->bufs[op.src[0].buf];
  const Buf& db = p->bufs[op.dst.buf];
  const act_t* src = reinterpret_cast<const act_t*>(buf_ptr(p, op.src[0].buf)) + op.src[0].c_off;
  act_t* dst = reinterpret_cast<act_t*>(buf_ptr(p, op.dst.buf)) + op.dst.c_off;
  int C = op.dst.C;
  const int per_row = ((op.Wout + DW_PX - 1) / DW_PX) * (C >> 3);
  int threads = std::min(256, round_up(per_row, 32));
  if (per_row > 256) threads = round_up((per_row + (per_row + 255) / 256 - 1) / ((per_row + 255) / 256), 32);
  dim3 blocks((unsigned)((per_row + threads - 1) / threads), (unsigned)op.Hout, (unsigned)p->B);
  if (p->act_f16)
    YB_CUDA(launch_pdl(dwconv3x3_kernel<true>, blocks, dim3(threads), 0, st, src, sb.C, dst, db.C, w, cpad8(C), p->B,
                       op.Hout, op.Wout, C, op.dw_gsz, op.dw_gstride, op.dw_goff, op.act, op.dw_add));
  else
    YB_CUDA(launch_pdl(dwconv3x3_kernel<false>, blocks, dim3(threads), 0, st, src, sb.C, dst, db.C, w, cpad8(C), p->B,
                       op.Hout, op.Wout, C, op.dw_gsz, op.dw_gstride, op.dw_goff, op.act, op.dw_add));
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

// ---------------------------------------------------------------------------------------------
// SPPF pooling (nets/nn.py:88,92-94): three cascaded MaxPool2d(5,1,2).  One CTA holds an
// (image, 8-channel group) plane in shared memory and runs the cascade as separable row/column
// 5-max passes; slice 0 of the concat buffer is read once, slices 1..3 are written once.
// ---------------------------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ uint4 max_act8(const uint4& a, const uint4& b) {
  return make_uint4(Act16<F16>::max2(a.x, b.x), Act16<F16>::max2(a.y, b.y), Act16<F16>::max2(a.z, b.z),
                    Act16<F16>::max2(a.w, b.w));
}

// G: 8-channel groups per CTA (a pixel's G x 16 bytes are contiguous: full sectors, fewer and fatter CTAs)
template <int G, bool F16>
__global__ void __launch_bounds__(G >= 4 ? 512 : 256)
    sppf_pool_kernel(const act_t* __restrict__ src, act_t* dst, int ld, int H, int W,
                     int Chalf) {
  extern __shared__ uint4 pl[];  // cur[HW][G], tmp[HW][G]
  pdl_prologue_done();
  pdl_wait();
  const int HW = H * W, n = HW * G;
  uint4* cur = pl;
  uint4* tmp = pl + n;
  const int cgs = (Chalf >> 3) / G;
  const int b = blockIdx.x / cgs;
  const int cg = blockIdx.x % cgs;
  const size_t img_base = (size_t)b * HW * ld + cg * 8 * G;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    cur[i] = __ldg(reinterpret_cast<const uint4*>(src + img_base + (size_t)(i / G) * ld + (i % G) * 8));
  __syncthreads();
  for (int stage = 0; stage < 3; stage++) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int pix = i / G, y = pix / W, x = pix - y * W;
      uint4 m = cur[i];
#pragma unroll
      for (int d = -2; d <= 2; d++) {
        int xx = x + d;
        if (d != 0 && (unsigned)xx < (unsigned)W) m = max_act8<F16>(m, cur[i + d * G]);
      }
      tmp[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int pix = i / G, g = i % G, y = pix / W;
      uint4 m = tmp[i];
#pragma unroll
      for (int d = -2; d <= 2; d++) {
        int yy = y + d;
        if (d != 0 && (unsigned)yy < (unsigned)H) m = max_act8<F16>(m, tmp[i + d * W * G]);
      }
      // dst points at slice 1 of the concat buffer; slices are Chalf channels apart
      *reinterpret_cast<uint4*>(dst + img_base + (size_t)pix * ld + g * 8 + (size_t)stage * Chalf) = m;
      cur[i] = m;  // each thread rewrites only the pixels it owns; tmp is the read set
    }
    __syncthreads();
  }
}

int launch_pool(const yb_plan* p, const Op& op, cudaStream_t st) {
  const Buf& sb = p->bufs[op.src[0].buf];
  const act_t* src = reinterpret_cast<const act_t*>(buf_ptr(p, op.src[0].buf)) + op.src[0].c_off;
  act_t* dst = reinterpret_cast<act_t*>(buf_ptr(p, op.dst.buf)) + op.dst.c_off;
  int Chalf = op.src[0].C;
  int HW = op.Hin * op.Win;
  const int cg_all = Chalf >> 3;
  const int G = (cg_all % 4 == 0 && (size_t)2 * HW * 4 * 16 <= 96 * 1024) ? 4 : (cg_all % 2 == 0 ? 2 : 1);
  size_t smem = (size_t)2 * HW * G * 16;
  if (smem > 200 * 1024) {
    set_error("SPPF plane %dx%d does not fit shared memory", op.Hin, op.Win);
    return YB_ERR_UNSUPPORTED;
  }
  // src (slice 0) and dst (slices 1..3) live in the same concat buffer
  unsigned blocks = (unsigned)(p->B * (cg_all / G));
#define YB_SPPF(GG, FF)                                                                                          \
  do {                                                                                                           \
    static size_t attr_dev[YB_MAX_DEVICES] = {0};   /* per device */                                             \
    size_t& attr = attr_dev[p->device & (YB_MAX_DEVICES - 1)];                                                   \
    if (smem > 48 * 1024 && smem > attr) {                                                                       \
      YB_CUDA(cudaFuncSetAttribute(sppf_pool_kernel<GG, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr = smem;                                                                                               \
    }                                                                                                            \
    YB_CUDA(launch_pdl(sppf_pool_kernel<GG, FF>, dim3(blocks), dim3(GG >= 4 ? 512 : 256), smem, st, src, dst, sb.C, op.Hin, op.Win, \
                       Chalf));                                                                                  \
  } while (0)
  if (p->act_f16) {
    if (G == 4) YB_SPPF(4, true);
    else if (G == 2) YB_SPPF(2, true);
    else YB_SPPF(1, true);
  } else {
    if (G == 4) YB_SPPF(4, false);
    else if (G == 2) YB_SPPF(2, false);
    else YB_SPPF(1, false);
  }
#undef YB_SPPF
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

// ---------------------------------------------------------------------------------------------
// C2PSA attention core (nets/nn.py:112-122): per (image, head)
//     out[:, i] = sum_j softmax_j( (q_i . k_j) * dk^-1/2 ) v_j
// qkv rows are tokens; channels of head h are [q(32) | k(32) | v(64)] at h*128 (the reference's
// view(b, heads, 2*dk+dh, N) split).  Flash-style tensor-core kernel: a CTA owns 128 queries of one
// (image, head), each warp a 16-query strip.  K/V chunks of 128 keys are staged in shared memory as
// bf16 (rows padded so fragment reads are conflict-free); S = Q K^T and O += P V run as
// mma.sync.m16n8k16 bf16 with fp32 accumulators, the S accumulator layout doubling as the A
// fragment of the second MMA; softmax is computed online in base 2 on the fp32 S registers.
// (N = 400 keys of 96 channels is far too small a problem for a tcgen05/TMEM pipeline to pay.)
// ---------------------------------------------------------------------------------------------
static constexpr int ATT_Q = 128;   // queries per CTA (8 warps x 16)
static constexpr int ATT_KC = 128;  // keys per shared-memory chunk
static constexpr int ATT_KP = 40;   // K row pitch, bf16 (80 B)
static constexpr int ATT_VP = 72;   // V row pitch, bf16 (144 B)

template <bool F16>
__global__ void __launch_bounds__(256)
    attention_kernel(const act_t* __restrict__ qkv, int qkv_ld, act_t* __restrict__ out,
                     int out_ld, int N, int heads, float scale_log2e) {
  using A16 = Act16<F16>;
  __shared__ __align__(16) act_t Ks[ATT_KC * ATT_KP];
  __shared__ __align__(16) act_t Vs[ATT_KC * ATT_VP];
  pdl_prologue_done();
  pdl_wait();
  const int h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * ATT_Q + warp * 16;
  const act_t* base = qkv + (size_t)b * N * qkv_ld + h * 128;
  // Q fragments (A operand, 16 queries x 32 channels = 2 k-steps), rows clamped into the image
  uint32_t qa[2][4];
  {
    const act_t* r0 = base + (size_t)min(q0 + g, N - 1) * qkv_ld;
    const act_t* r1 = base + (size_t)min(q0 + g + 8, N - 1) * qkv_ld;
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
      qa[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(r0 + ks * 16 + 2 * t));
      qa[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(r1 + ks * 16 + 2 * t));
      qa[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(r0 + ks * 16 + 8 + 2 * t));
      qa[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(r1 + ks * 16 + 8 + 2 * t));
    }
  }
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const uint32_t ks_s = (uint32_t)__cvta_generic_to_shared(Ks);
  const uint32_t vs_s = (uint32_t)__cvta_generic_to_shared(Vs);

  for (int k0 = 0; k0 < N; k0 += ATT_KC) {
    __syncthreads();
    // stage K (KC x 32) and V (KC x 64): 12 granules of 8 channels per key, zero-filled past N
    // (192 threads: a thread keeps its granule and walks the keys in steps of 16 - no division per copy)
    if (tid < 192) {
      const int gr = tid % 12, key0 = tid / 12;
      const uint32_t dp0 = gr < 4 ? ks_s + (uint32_t)(key0 * ATT_KP + gr * 8) * 2u
                                  : vs_s + (uint32_t)(key0 * ATT_VP + (gr - 4) * 8) * 2u;
      const uint32_t dstep = (gr < 4 ? (uint32_t)ATT_KP : (uint32_t)ATT_VP) * 16u * 2u;
      const act_t* sp0 = base + 32 + gr * 8;
#pragma unroll
      for (int it = 0; it < ATT_KC / 16; it++) {
        const int key = key0 + 16 * it;
        const bool ok = k0 + key < N;
        const act_t* sp = sp0 + (size_t)(ok ? k0 + key : 0) * qkv_ld;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dp0 + (uint32_t)it * dstep), "l"(sp), "r"(ok ? 16u : 0u)
                     : "memory");
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    const int kmax = min(ATT_KC, N - k0);
    for (int kb = 0; kb < kmax; kb += 64) {
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        const act_t* kr = Ks + (kb + nt * 8 + g) * ATT_KP + 2 * t;
#pragma unroll
        for (int ks = 0; ks < 2; ks++)
          A16::mma16816(s[nt], qa[ks], *reinterpret_cast<const uint32_t*>(kr + ks * 16),
                         *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8));
      }
      // The scores stay unscaled: max(scale * s) = scale * max(s) for scale > 0, and p = 2^(scale * s - m) is one FFMA +
      // ex2.approx per score (exp2f's range handling and a separate multiply were a third of the kernel's instructions);
      // keys past N are masked only in the sub-block that holds the boundary.
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (kb + 64 > kmax) {   // uniform
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          const int key = kb + nt * 8 + 2 * t;
#pragma unroll
          for (int e = 0; e < 4; e++)
            if (key + (e & 1) >= kmax) s[nt][e] = -INFINITY;
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0 * scale_log2e), mn1 = fmaxf(m1, mx1 * scale_log2e);  // finite: every sub-block holds a valid key
      const float c0 = ex2_approx(m0 - mn0), c1 = ex2_approx(m1 - mn1);
      m0 = mn0;
      m1 = mn1;
      l0 *= c0;
      l1 *= c1;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        o[i][0] *= c0;
        o[i][1] *= c0;
        o[i][2] *= c1;
        o[i][3] *= c1;
      }
      uint32_t pa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        const float p0 = ex2_approx(fmaf(s[nt][0], scale_log2e, -mn0)), p1 = ex2_approx(fmaf(s[nt][1], scale_log2e, -mn0));
        const float p2 = ex2_approx(fmaf(s[nt][2], scale_log2e, -mn1)), p3 = ex2_approx(fmaf(s[nt][3], scale_log2e, -mn1));
        l0 += p0 + p1;
        l1 += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2] = A16::pack2(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = A16::pack2(p2, p3);
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        // V fragments (B operand, k = key, n = channel) through ldmatrix.trans on [key][channel] rows
        const uint32_t vrow = vs_s + (uint32_t)((kb + j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_VP +
                                                (lane >> 4) * 8) * 2u;
#pragma unroll
        for (int dp = 0; dp < 4; dp++) {
          uint32_t r0, r1, r2, r3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                       : "r"(vrow + (uint32_t)dp * 32u));
          A16::mma16816(o[2 * dp], pa[j], r0, r1);
          A16::mma16816(o[2 * dp + 1], pa[j], r2, r3);
        }
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  act_t* ob = out + (size_t)b * N * out_ld + h * 64 + 2 * t;
  if (q0 + g < N) {
    act_t* op = ob + (size_t)(q0 + g) * out_ld;
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
      *reinterpret_cast<uint32_t*>(op + nt * 8) = A16::pack2(o[nt][0] * i0, o[nt][1] * i0);
  }
  if (q0 + g + 8 < N) {
    act_t* op = ob + (size_t)(q0 + g + 8) * out_ld;
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
      *reinterpret_cast<uint32_t*>(op + nt * 8) = A16::pack2(o[nt][2] * i1, o[nt][3] * i1);
  }
}

int launch_attn(const yb_plan* p, const Op& op, cudaStream_t st) {
  const Buf& sb = p->bufs[op.src[0].buf];
  const Buf& db = p->bufs[op.dst.buf];
  const act_t* qkv = reinterpret_cast<const act_t*>(buf_ptr(p, op.src[0].buf)) + op.src[0].c_off;
  act_t* out = reinterpret_cast<act_t*>(buf_ptr(p, op.dst.buf)) + op.dst.c_off;
  if (op.dk != 32 || op.dh != 64) {
    set_error("attention kernel is specialised for dim_key 32 / dim_head 64 (got %d / %d)", op.dk, op.dh);
    return YB_ERR_UNSUPPORTED;
  }
  int N = op.Hin * op.Win;
  dim3 grid((N + ATT_Q - 1) / ATT_Q, op.heads, p->B);
  if (p->act_f16)
    YB_CUDA(launch_pdl(attention_kernel<true>, grid, dim3(256), 0, st, qkv, sb.C, out, db.C, N, op.heads,
                       op.scale * 1.4426950408889634f));
  else
    YB_CUDA(launch_pdl(attention_kernel<false>, grid, dim3(256), 0, st, qkv, sb.C, out, db.C, N, op.heads,
                       op.scale * 1.4426950408889634f));
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

// ---------------------------------------------------------------------------------------------
// Head decode (nets/nn.py:261-270 + DFL nn.py:222-225 + make_anchors utils/util.py:85-96):
// logits (B, A, 64+nc) fp32 -> out (B, 4+nc, A) fp32.  Anchor points and strides are computed
// from the anchor index, never materialised.  Tiles of 64 anchors are transposed through shared
// memory so that both the logits read and the plane-major write are coalesced.
// ---------------------------------------------------------------------------------------------
struct DecodeParams {
  int B, A, nc, ld;
  int lvl_off[3], lvl_w[3];
  float lvl_stride[3];
};

static constexpr int DEC_T = 64;

__global__ void __launch_bounds__(256)
    head_decode_kernel(const float* __restrict__ logits, float* __restrict__ out, const DecodeParams D) {
  extern __shared__ float tile[];  // [DEC_T][ldp]
  pdl_prologue_done();
  pdl_wait();
  const int no = 64 + D.nc;
  const int ldp = no | 1;  // odd row pitch: conflict-free column access
  float* dist = tile + DEC_T * ldp;  // [DEC_T][4]
  const long long row0 = (long long)blockIdx.x * DEC_T;
  const long long rows_total = (long long)D.B * D.A;
  const int tid = threadIdx.x;
  for (int i = tid; i < DEC_T * no; i += blockDim.x) {
    int r = i / no, c = i - r * no;
    long long row = row0 + r;
    tile[r * ldp + c] = row < rows_total ? __ldg(logits + row * D.ld + c) : 0.f;
  }
  __syncthreads();
  {  // DFL: softmax over 16 bins, expectation with weights 0..15 (nn.py:218-225)
    int r = tid & (DEC_T - 1), side = tid >> 6;  // 64 anchors x 4 sides = 256 threads
    const float* lp = tile + r * ldp + side * 16;
    float mx = lp[0];
#pragma unroll
    for (int j = 1; j < 16; j++) mx = fmaxf(mx, lp[j]);
    float se = 0.f, sw = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      float e = expf(lp[j] - mx);
      se += e;
      sw = fmaf((float)j, e, sw);
    }
    dist[r * 4 + side] = sw / se;
  }
  __syncthreads();
  const int r = tid & (DEC_T - 1);
  const long long row = row0 + r;
  if (row >= rows_total) return;
  const int b = (int)(row / D.A);
  const int a = (int)(row - (long long)b * D.A);
  float* ob = out + (size_t)b * (4 + D.nc) * D.A + a;
  const int part = tid >> 6;
  if (part == 0) {
    int lvl = a >= D.lvl_off[2] ? 2 : (a >= D.lvl_off[1] ? 1 : 0);
    int i = a - D.lvl_off[lvl];
    int y = i / D.lvl_w[lvl], x = i - y * D.lvl_w[lvl];
    float ax = (float)x + 0.5f, ay = (float)y + 0.5f, s = D.lvl_stride[lvl];
    float x1 = ax - dist[r * 4 + 0], y1 = ay - dist[r * 4 + 1];
    float x2 = ax + dist[r * 4 + 2], y2 = ay + dist[r * 4 + 3];
    ob[0] = (x1 + x2) / 2.f * s;
    ob[(size_t)D.A] = (y1 + y2) / 2.f * s;
    ob[(size_t)2 * D.A] = (x2 - x1) * s;
    ob[(size_t)3 * D.A] = (y2 - y1) * s;
  }
  for (int j = part; j < D.nc; j += 4) {
    float z = tile[r * ldp + 64 + j];
    ob[(size_t)(4 + j) * D.A] = 1.f / (1.f + expf(-z));
  }
}

int launch_decode(const yb_plan* p, const float* logits, float* out, cudaStream_t st) {
  DecodeParams D;
  D.B = p->B;
  D.A = p->A;
  D.nc = p->nc;
  D.ld = p->bufs[p->logits_buf].C;
  for (int i = 0; i < 3; i++) {
    D.lvl_off[i] = p->lvl_off[i];
    D.lvl_w[i] = p->lvl_w[i];
    D.lvl_stride[i] = p->lvl_stride[i];
  }
  int no = 64 + p->nc;
  size_t smem = ((size_t)DEC_T * (no | 1) + DEC_T * 4) * 4;
  static size_t attr_dev[YB_MAX_DEVICES] = {0};   // per device
  size_t& attr = attr_dev[p->device & (YB_MAX_DEVICES - 1)];
  if (smem > 48 * 1024 && smem > attr) {
    YB_CUDA(cudaFuncSetAttribute(head_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    attr = smem;
  }
  long long rows = (long long)p->B * p->A;
  unsigned blocks = (unsigned)((rows + DEC_T - 1) / DEC_T);
  YB_CUDA(launch_pdl(head_decode_kernel, dim3(blocks), dim3(256), smem, st, logits, out, D));
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb
