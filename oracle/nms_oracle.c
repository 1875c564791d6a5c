/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C restatement of the reference's non_max_suppression for ONE image
 * (t0saki/YOLO-Infer-pt utils/util.py:123-169), including the greedy step that the reference
 * delegates to torchvision.ops.nms (third-party, not under /root/reference; de-facto version
 * torchvision 0.26.0+cu128, CPU kernel nms_kernel_impl).  Pinned by tests/golden/nms_*.npz, which
 * hold outputs of the reference itself, and cross-checked against torchvision in
 * tests/golden/make_golden.py.
 *
 * Steps (line numbers are utils/util.py):
 *   130,147  candidates = every (anchor, class) with score > conf (fp32 compare), enumerated
 *            anchor-major / class-minor like nonzero() on the (n_anchors, nc) matrix
 *   144-145  box = wh2xy(cx, cy, w, h)                          (util.py:76-82)
 *   157      sort by score descending, keep the first max_nms   (ties: ascending candidate index —
 *            the reference's argsort is unstable there; parity is asserted on tie-free scores)
 *   160-161  boxes + class * max_wh in fp32
 *   162      greedy NMS: walk in score order, suppress j when inter/(area_i+area_j-inter) > iou,
 *            fp32 arithmetic one rounding per operation, compared against the double threshold
 *   163-165  first max_det kept rows, un-offset boxes
 * Build: gcc -O2 -fPIC -shared -fno-fast-math -ffp-contract=off (no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float score;
  uint32_t idx; /* anchor * nc + class */
} cand_t;

static int cmp_cand(const void* pa, const void* pb) {
  const cand_t* a = (const cand_t*)pa;
  const cand_t* b = (const cand_t*)pb;
  if (a->score > b->score) return -1;
  if (a->score < b->score) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx);
}

static inline float maxf_std(float a, float b) { return (a < b) ? b : a; } /* std::max */
static inline float minf_std(float a, float b) { return (b < a) ? b : a; } /* std::min */

/* pred: (4+nc, A) fp32 row-major; out: (max_det, 6); returns the number of rows written.
 * full_scan != 0 runs the greedy loop over every sorted box like torchvision does and truncates
 * afterwards; 0 stops once max_det boxes are kept (same result, used for big inputs). */
int nms_oracle_image(const float* pred, int nc, int A, float conf, double iou, int max_det, int max_nms,
                     float max_wh, int full_scan, float* out) {
  size_t cap = 1024, n = 0;
  cand_t* c = (cand_t*)malloc(cap * sizeof(cand_t));
  for (int a = 0; a < A; a++) {
    for (int k = 0; k < nc; k++) {
      float s = pred[(size_t)(4 + k) * A + a];
      if (s > conf) {
        if (n == cap) {
          cap *= 2;
          c = (cand_t*)realloc(c, cap * sizeof(cand_t));
        }
        c[n].score = s;
        c[n].idx = (uint32_t)a * (uint32_t)nc + (uint32_t)k;
        n++;
      }
    }
  }
  if (n == 0) {
    free(c);
    return 0;
  }
  qsort(c, n, sizeof(cand_t), cmp_cand);
  if (n > (size_t)max_nms) n = (size_t)max_nms;

  float* x1 = (float*)malloc(n * sizeof(float) * 9);
  float *y1 = x1 + n, *x2 = y1 + n, *y2 = x2 + n, *area = y2 + n;
  float *rx1 = area + n, *ry1 = rx1 + n, *rx2 = ry1 + n, *ry2 = rx2 + n;
  for (size_t i = 0; i < n; i++) {
    uint32_t a = c[i].idx / (uint32_t)nc, k = c[i].idx % (uint32_t)nc;
    float cx = pred[a], cy = pred[(size_t)A + a], w = pred[(size_t)2 * A + a], h = pred[(size_t)3 * A + a];
    volatile float hw = w / 2.0f, hh = h / 2.0f;
    rx1[i] = cx - hw;
    ry1[i] = cy - hh;
    rx2[i] = cx + hw;
    ry2[i] = cy + hh;
    volatile float off = (float)k * max_wh;
    x1[i] = rx1[i] + off;
    y1[i] = ry1[i] + off;
    x2[i] = rx2[i] + off;
    y2[i] = ry2[i] + off;
    volatile float dw = x2[i] - x1[i], dh = y2[i] - y1[i];
    area[i] = dw * dh;
  }
  unsigned char* dead = (unsigned char*)calloc(n, 1);
  int kept = 0;
  for (size_t i = 0; i < n; i++) {
    if (dead[i]) continue;
    if (kept < max_det) {
      float* o = out + (size_t)kept * 6;
      o[0] = rx1[i];
      o[1] = ry1[i];
      o[2] = rx2[i];
      o[3] = ry2[i];
      o[4] = c[i].score;
      o[5] = (float)(c[i].idx % (uint32_t)nc);
    }
    kept++;
    if (!full_scan && kept >= max_det) break;
    float ix1 = x1[i], iy1 = y1[i], ix2 = x2[i], iy2 = y2[i], ia = area[i];
    for (size_t j = i + 1; j < n; j++) {
      if (dead[j]) continue;
      float xx1 = maxf_std(ix1, x1[j]);
      float yy1 = maxf_std(iy1, y1[j]);
      float xx2 = minf_std(ix2, x2[j]);
      float yy2 = minf_std(iy2, y2[j]);
      volatile float dw = xx2 - xx1, dh = yy2 - yy1;
      float w = maxf_std(0.0f, dw);
      float h = maxf_std(0.0f, dh);
      volatile float inter = w * h;
      volatile float sum = ia + area[j];
      volatile float uni = sum - inter;
      volatile float ovr = inter / uni;
      if ((double)ovr > iou) dead[j] = 1;
    }
  }
  free(dead);
  free(x1);
  free(c);
  return kept < max_det ? kept : max_det;
}

/* Batched wrapper: pred (B, 4+nc, A); out (B, max_det, 6); counts (B). */
void nms_oracle_batch(const float* pred, int B, int nc, int A, float conf, double iou, int max_det,
                      int max_nms, float max_wh, int full_scan, float* out, int* counts) {
  for (int b = 0; b < B; b++)
    counts[b] = nms_oracle_image(pred + (size_t)b * (4 + nc) * A, nc, A, conf, iou, max_det, max_nms, max_wh,
                                 full_scan, out + (size_t)b * max_det * 6);
}
