// Eval-time image pre-processing on the device: the step immediately in front of YOLO.forward in the
// reference.  One kernel replaces, per image,
//   Dataset.load_image   utils/dataset.py:95-103   cv2.resize(..., INTER_LINEAR) to input_size / max(h, w)
//   resize()             utils/dataset.py:292-313  letterbox with cv2.copyMakeBorder (constant 0)
//   __getitem__          utils/dataset.py:86-88    HWC -> CHW, BGR -> RGB
// and writes the (B, 3, S, S) uint8 batch the stem kernel consumes (which folds main.py:266-267's /255).
// Bit-exact with OpenCV's 8-bit INTER_LINEAR: 11-bit fixed-point coefficients from float32 fractions,
// int32 horizontal pass, vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; the x
// direction clamps the fraction at the image edge, the y direction only clamps the row index.
// HBM-bound: every source pixel is read through L1/L2 (<= 4 taps per output), every output byte is
// written once, 4 output pixels per thread so that plane stores are 4-byte wide and coalesced.
#include "yb_internal.h"

namespace yb {

struct LbGeom {
  int dh, dw, top, left;
  double ratio, pad_w, pad_h;
};

// geometry exactly as the reference's Python computes it (double arithmetic, int() truncation,
// round-half-even on values that are never ties)
__host__ __device__ inline LbGeom lb_geometry(int h, int w, int S) {
  LbGeom g;
  const int mx = h > w ? h : w;
  const double r = (double)S / (double)mx;
  g.ratio = r;
  if (r != 1.0) {
    g.dh = (int)((double)h * r);
    g.dw = (int)((double)w * r);
  } else {
    g.dh = h;
    g.dw = w;
  }
  g.pad_w = (double)(S - g.dw) / 2.0;
  g.pad_h = (double)(S - g.dh) / 2.0;
  g.top = (int)rint(g.pad_h - 0.1);
  g.left = (int)rint(g.pad_w - 0.1);
  return g;
}

// source index and 11-bit coefficients of destination index d (OpenCV resizeGeneric_, linear)
__device__ __forceinline__ void lb_coeff(int d, int dn, int sn, bool clamp_frac, int& s, int& a0, int& a1) {
  const double inv = (double)dn / (double)sn;
  const double scale = 1.0 / inv;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  s = (int)floorf(f);
  f = f - (float)s;
  if (clamp_frac) {
    if (s < 0) {
      f = 0.f;
      s = 0;
    }
    if (s >= sn - 1) {
      f = 0.f;
      s = sn - 1;
    }
  }
  a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
}

static constexpr int LB_ROWS = 8;   // output rows per block

// desc: [B][3] int64 = (device pointer of the HWC uint8 BGR image, height, width)
__global__ void __launch_bounds__(256)
    letterbox_kernel(const long long* __restrict__ desc, uint8_t* __restrict__ out, double* __restrict__ meta,
                     int S) {
  pdl_prologue_done();
  pdl_wait();
  const int b = blockIdx.z;
  const uint8_t* __restrict__ src = reinterpret_cast<const uint8_t*>(desc[3 * b]);
  const int h = (int)desc[3 * b + 1], w = (int)desc[3 * b + 2];
  // the letterbox geometry is double arithmetic (the reference's Python): once per block, not once per thread
  __shared__ LbGeom g_s;
  if (threadIdx.x == 0) g_s = lb_geometry(h, w, S);
  __syncthreads();
  const LbGeom g = g_s;
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && meta) {
    meta[3 * b] = g.ratio;
    meta[3 * b + 1] = g.pad_w;
    meta[3 * b + 2] = g.pad_h;
  }
  if (x0 >= S) return;
  // a block walks LB_ROWS output rows (the per-block geometry and launch costs are paid once per 8 rows)
  for (int y = blockIdx.y * LB_ROWS; y < min(S, (int)(blockIdx.y + 1) * LB_ROWS); y++) {
  uint32_t px[3] = {0u, 0u, 0u};   // 4 output pixels per plane (R, G, B), one byte each
  const int sy = y - g.top;
  if (sy >= 0 && sy < g.dh) {
    const bool copy = g.dh == h && g.dw == w;
    int ys = sy, b0 = 2048, b1 = 0;
    if (!copy) lb_coeff(sy, g.dh, h, false, ys, b0, b1);
    const int y0 = min(max(ys, 0), h - 1), y1 = min(max(ys + 1, 0), h - 1);
    const uint8_t* r0 = src + (size_t)y0 * w * 3;
    const uint8_t* r1 = src + (size_t)y1 * w * 3;
    const int sx0 = x0 - g.left;
    if (copy && sx0 >= 0 && sx0 + 3 < g.dw) {
      // frames that already have the model's long side (ratio 1: pad + BGR -> RGB + HWC -> CHW only): the four
      // pixels are 12 consecutive bytes - four aligned 32-bit loads and byte permutes instead of 12 byte loads
      const uint8_t* p = src + ((size_t)sy * w + sx0) * 3;
      const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
      const uint32_t* q = reinterpret_cast<const uint32_t*>(p - mis);
      const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
      const uint32_t w3 = mis ? __ldg(q + 3) : 0u;   // (a row's last pixels: q + 3 stays inside the image while mis > 0)
      const uint32_t sh = mis * 8u;
      const uint32_t d0 = __funnelshift_r(w0, w1, sh), d1 = __funnelshift_r(w1, w2, sh), d2 = __funnelshift_r(w2, w3, sh);
      // d0 = B0 G0 R0 B1 | d1 = G1 R1 B2 G2 | d2 = R2 B3 G3 R3  (little endian, lowest byte first)
      px[0] = (uint32_t)((d0 >> 16) & 0xFFu) | (uint32_t)(((d1 >> 8) & 0xFFu) << 8) | (uint32_t)((d2 & 0xFFu) << 16) |
              (uint32_t)(((d2 >> 24) & 0xFFu) << 24);
      px[1] = (uint32_t)((d0 >> 8) & 0xFFu) | (uint32_t)((d1 & 0xFFu) << 8) | (uint32_t)(((d1 >> 24) & 0xFFu) << 16) |
              (uint32_t)(((d2 >> 16) & 0xFFu) << 24);
      px[2] = (uint32_t)(d0 & 0xFFu) | (uint32_t)(((d0 >> 24) & 0xFFu) << 8) | (uint32_t)(((d1 >> 16) & 0xFFu) << 16) |
              (uint32_t)(((d2 >> 8) & 0xFFu) << 24);
    } else
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int sx = x0 + i - g.left;
      if (sx < 0 || sx >= g.dw) continue;
      if (copy) {
        const uint8_t* p = src + ((size_t)sy * w + sx) * 3;
        px[0] |= (uint32_t)p[2] << (8 * i);   // BGR -> RGB planes
        px[1] |= (uint32_t)p[1] << (8 * i);
        px[2] |= (uint32_t)p[0] << (8 * i);
      } else {
        int xs, a0, a1;
        lb_coeff(sx, g.dw, w, true, xs, a0, a1);
        const int x1 = min(xs + 1, w - 1);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const int S0 = (int)r0[xs * 3 + c] * a0 + (int)r0[x1 * 3 + c] * a1;
          const int S1 = (int)r1[xs * 3 + c] * a0 + (int)r1[x1 * 3 + c] * a1;
          int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
          v = min(max(v, 0), 255);
          px[2 - c] |= (uint32_t)v << (8 * i);
        }
      }
    }
  }
  uint8_t* ob = out + ((size_t)b * 3 * S + y) * S + x0;
  if (x0 + 3 < S && (S & 3) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) *reinterpret_cast<uint32_t*>(ob + (size_t)c * S * S) = px[c];
  } else {
    for (int i = 0; i < 4 && x0 + i < S; i++)
      for (int c = 0; c < 3; c++) ob[(size_t)c * S * S + i] = (uint8_t)(px[c] >> (8 * i));
  }
  }
}

int letterbox_run(const long long* desc, int B, int S, uint8_t* out, double* meta, cudaStream_t st) {
  if (!desc || !out || B <= 0 || S <= 0) {
    set_error("yb_letterbox: bad arguments (batch %d, size %d)", B, S);
    return YB_ERR_ARG;
  }
  const int threads = S >= 1024 ? 256 : 160;   // 4 pixels per thread: one block spans a 640-pixel row exactly
  dim3 grid((unsigned)((S + 4 * threads - 1) / (4 * threads)), (unsigned)((S + LB_ROWS - 1) / LB_ROWS), (unsigned)B);
  YB_CUDA(launch_pdl(letterbox_kernel, grid, dim3(threads), 0, st, desc, out, meta, S));
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb
