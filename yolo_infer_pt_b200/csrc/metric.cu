// Detection consumer on the device: the reference's `compute_metric` (utils/util.py:99-120), the step
// immediately behind non_max_suppression in `test()` (main.py:271-299).  The
// reference builds the IoU matrix on the GPU, then round-trips every image through numpy (argsort +
// two numpy.unique per IoU threshold).  That reduction has a closed form: per detection d the best
// class-matched label l*(d) with IoU i*(d) (ties: the larger label index, the order argsort()[::-1]
// yields), and d is a true positive at threshold t iff i*(d) >= t and no earlier detection with the same
// l* also reaches t.  One CTA per image computes it for the whole padded batch NMS produced; nothing
// leaves the device.  IoU arithmetic is fp32 in the reference's operation order, without contraction.
#include "yb_internal.h"

namespace yb {

static constexpr int MET_MAX_T = 16;

__global__ void __launch_bounds__(256)
    metric_kernel(const float* __restrict__ det, const int* __restrict__ counts, const float* __restrict__ tgt,
                  const int* __restrict__ tcounts, int max_det, int max_t, const float* __restrict__ iou_v, int n_iou,
                  uint8_t* __restrict__ correct) {
  extern __shared__ float msm[];   // [max_t][5] labels, then best IoU [max_det], best label [max_det]
  pdl_prologue_done();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(counts[b], max_det), m = min(tcounts[b], max_t);
  float* lab = msm;
  float* bi = msm + (size_t)max_t * 5;
  int* bl = reinterpret_cast<int*>(bi + max_det);
  for (int i = tid; i < m * 5; i += blockDim.x) lab[i] = tgt[(size_t)b * max_t * 5 + i];
  __syncthreads();
  for (int d = tid; d < n; d += blockDim.x) {
    const float* o = det + ((size_t)b * max_det + d) * 6;
    const float x1 = o[0], y1 = o[1], x2 = o[2], y2 = o[3], cls = o[5];
    const float area_b = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    float best = -1.f;
    int which = -1;
    for (int l = 0; l < m; l++) {
      const float* t = lab + l * 5;
      if (t[0] != cls) continue;
      const float w = fmaxf(__fsub_rn(fminf(t[3], x2), fmaxf(t[1], x1)), 0.f);
      const float h = fmaxf(__fsub_rn(fminf(t[4], y2), fmaxf(t[2], y1)), 0.f);
      const float inter = __fmul_rn(w, h);
      const float area_a = __fmul_rn(__fsub_rn(t[3], t[1]), __fsub_rn(t[4], t[2]));
      const float uni = __fadd_rn(__fsub_rn(__fadd_rn(area_a, area_b), inter), 1e-7f);
      const float iou = __fdiv_rn(inter, uni);
      if (iou >= best) {   // ties: the larger label index wins (argsort()[::-1] order of util.py:116)
        best = iou;
        which = l;
      }
    }
    bi[d] = best;
    bl[d] = which;
  }
  __syncthreads();
  for (int d = tid; d < max_det; d += blockDim.x) {
    uint8_t* c = correct + ((size_t)b * max_det + d) * n_iou;
    float mine = -1.f, earlier = -1.f;
    if (d < n && bl[d] >= 0) {
      mine = bi[d];
      const int l = bl[d];
      for (int e = 0; e < d; e++)
        if (bl[e] == l) earlier = fmaxf(earlier, bi[e]);
    }
    for (int t = 0; t < n_iou; t++) {
      const float thr = iou_v[t];
      c[t] = (mine >= thr && !(earlier >= thr)) ? 1 : 0;
    }
  }
}

int metric_run(const float* det, const int* counts, const float* tgt, const int* tcounts, int B, int max_det,
               int max_t, const float* iou_v, int n_iou, uint8_t* correct, cudaStream_t st) {
  if (!det || !counts || !tgt || !tcounts || !iou_v || !correct || B <= 0 || max_det <= 0 || max_t < 0 ||
      n_iou <= 0 || n_iou > MET_MAX_T) {
    set_error("yb_compute_metric: bad arguments (batch %d, max_det %d, max_targets %d, thresholds %d)", B, max_det,
              max_t, n_iou);
    return YB_ERR_ARG;
  }
  const size_t smem = ((size_t)max_t * 5 + (size_t)max_det * 2) * 4;
  if (smem > 200 * 1024) {
    set_error("yb_compute_metric: %d labels x %d detections do not fit shared memory", max_t, max_det);
    return YB_ERR_UNSUPPORTED;
  }
  static size_t attr_dev[YB_MAX_DEVICES] = {0};   // per-device function attribute
  int cur_dev = 0;
  YB_CUDA(cudaGetDevice(&cur_dev));
  size_t& attr = attr_dev[cur_dev & (YB_MAX_DEVICES - 1)];
  if (smem > 48 * 1024 && smem > attr) {
    YB_CUDA(cudaFuncSetAttribute(metric_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  YB_CUDA(launch_pdl(metric_kernel, dim3(B), dim3(256), smem, st, det, counts, tgt, tcounts, max_det, max_t, iou_v,
                     n_iou, correct));
  count_launch();
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb
