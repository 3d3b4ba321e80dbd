/* TEST INFRASTRUCTURE ONLY (checker, never shipped on the product path).
 *
 * Scalar C restatement of the reference's 3-D IoU and greedy NMS (Wuziyi616/CFUN utils.py:50-70, 122-157) in strict
 * fp32 with no FMA contraction (build with -ffp-contract=off), used to fuzz the numpy oracle and the CUDA kernels at
 * sizes where the numpy loop is slow (10k boxes).  Order contract: scores descending, index ascending on ties.
 */
#include <stdlib.h>
#include <string.h>

static float vol(const float* b) { return ((b[3] - b[0]) * (b[4] - b[1])) * (b[5] - b[2]); }

float cfun_ref_iou(const float* a, const float* b) {
  float z1 = a[0] > b[0] ? a[0] : b[0], z2 = a[3] < b[3] ? a[3] : b[3];
  float y1 = a[1] > b[1] ? a[1] : b[1], y2 = a[4] < b[4] ? a[4] : b[4];
  float x1 = a[2] > b[2] ? a[2] : b[2], x2 = a[5] < b[5] ? a[5] : b[5];
  float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
  dx = dx > 0.f ? dx : 0.f; dy = dy > 0.f ? dy : 0.f; dz = dz > 0.f ? dz : 0.f;
  float inter = (dx * dy) * dz;
  float uni = (vol(a) + vol(b)) - inter;
  return inter / (uni + 1e-6f);
}

typedef struct { float s; int i; } item;
static int cmp(const void* pa, const void* pb) {
  const item* a = (const item*)pa; const item* b = (const item*)pb;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  return a->i - b->i;
}

/* returns the number of kept indices written to keep[] (capacity max_num) */
int cfun_ref_nms(const float* boxes, const float* scores, int n, float thr, int max_num, int* keep) {
  item* it = (item*)malloc(sizeof(item) * (size_t)(n > 0 ? n : 1));
  char* dead = (char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int cnt = 0;
  for (int i = 0; i < n; ++i) { it[i].s = scores[i]; it[i].i = i; }
  qsort(it, (size_t)n, sizeof(item), cmp);
  for (int a = 0; a < n; ++a) {
    if (dead[a]) continue;
    keep[cnt++] = it[a].i;
    if (cnt >= max_num) break;
    for (int b = a + 1; b < n; ++b)
      if (!dead[b] && cfun_ref_iou(boxes + 6 * (size_t)it[a].i, boxes + 6 * (size_t)it[b].i) > thr) dead[b] = 1;
  }
  free(it); free(dead);
  return cnt;
}
