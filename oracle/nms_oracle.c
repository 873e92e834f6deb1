/*
 * ORACLE (test infrastructure only) - greedy NMS restated from the published algorithm of
 * torchvision 0.26.0 `torchvision::nms` for CPU float tensors (torchvision/csrc/ops/cpu/nms_kernel.cpp,
 * nms_kernel_impl<float>), the third-party op the reference calls through
 * torchvision.ops.boxes.batched_nms at yolox-drone/models/core/utils_bbox.py:414-419.
 * torchvision's sources are not vendored under /root/reference; this restatement is pinned against the
 * installed binary by tests/golden/make_golden.py + tests/test_oracle_cpu.py.
 *
 * Semantics kept bit-exact: all arithmetic in IEEE binary32 without fused multiply-add (build with
 * -ffp-contract=off), areas = (x2-x1)*(y2-y1), stable descending score order (ties -> lower index first),
 * suppress iff inter / (area_i + area_j - inter) > threshold (strict; NaN never suppresses).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void merge_sort_desc(const float* s, int32_t* idx, int32_t* tmp, int lo, int hi) {
  if (hi - lo < 2) return;
  int mid = lo + (hi - lo) / 2;
  merge_sort_desc(s, idx, tmp, lo, mid);
  merge_sort_desc(s, idx, tmp, mid, hi);
  int a = lo, b = mid, o = lo;
  while (a < mid && b < hi) {
    /* stable: take from the right run only when it is strictly greater */
    if (s[idx[b]] > s[idx[a]]) tmp[o++] = idx[b++];
    else tmp[o++] = idx[a++];
  }
  while (a < mid) tmp[o++] = idx[a++];
  while (b < hi) tmp[o++] = idx[b++];
  memcpy(idx + lo, tmp + lo, (size_t)(hi - lo) * sizeof(int32_t));
}

/* boxes [n][4] xyxy, scores [n]; keep_out [n] receives kept indices in descending score order.
 * returns the number kept, or -1 on allocation failure */
int oracle_nms(const float* boxes, const float* scores, int n, float iou_threshold, int32_t* keep_out) {
  if (n <= 0) return 0;
  int32_t* order = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  int32_t* tmp = (int32_t*)malloc((size_t)n * sizeof(int32_t));
  float* areas = (float*)malloc((size_t)n * sizeof(float));
  uint8_t* suppressed = (uint8_t*)calloc((size_t)n, 1);
  if (!order || !tmp || !areas || !suppressed) { free(order); free(tmp); free(areas); free(suppressed); return -1; }
  for (int i = 0; i < n; ++i) {
    order[i] = i;
    areas[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  }
  merge_sort_desc(scores, order, tmp, 0, n);
  int kept = 0;
  for (int _i = 0; _i < n; ++_i) {
    int i = order[_i];
    if (suppressed[i]) continue;
    keep_out[kept++] = i;
    float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    float iarea = areas[i];
    for (int _j = _i + 1; _j < n; ++_j) {
      int j = order[_j];
      if (suppressed[j]) continue;
      float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
      float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
      float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
      float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
      float w = xx2 - xx1; if (!(w > 0.0f)) w = 0.0f;
      float h = yy2 - yy1; if (!(h > 0.0f)) h = 0.0f;
      float inter = w * h;
      float ovr = inter / (iarea + areas[j] - inter);
      if (ovr > iou_threshold) suppressed[j] = 1;
    }
  }
  free(order); free(tmp); free(areas); free(suppressed);
  return kept;
}
