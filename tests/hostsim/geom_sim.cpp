// Host build of aidet_b200/csrc/geom.cuh -- TEST ONLY.
// Lets the CPU test-suite check the exact FP32 arithmetic the kernels run against the
// float64 oracle without a GPU.  Not part of the product and not a fallback.
#include <cstddef>
#include "../../aidet_b200/csrc/geom.cuh"
using namespace aidet;

extern "C" void sim_riou_matrix(const float* a, int m, const float* b, int n, int fmt, int mode, float* out) {
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < n; j++) {
      float v;
      if (fmt == 5) {
        RectRow r; RectCol c; rect_prepare(a + 5 * (size_t)i, &r, nullptr); rect_prepare(b + 5 * (size_t)j, nullptr, &c);
        v = rect_overlap(r, c, mode);
      } else {
        QuadRow r; QuadCol c; quad_prepare(a + 8 * (size_t)i, &r, nullptr); quad_prepare(b + 8 * (size_t)j, nullptr, &c);
        v = quad_overlap(r, c, mode);
      }
      out[(size_t)i * n + j] = v;
    }
  }
}
