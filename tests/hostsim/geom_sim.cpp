// Host build of aidet_b200/csrc/geom.cuh -- TEST ONLY.
// Lets the CPU test-suite check the exact FP32 arithmetic the kernels run against the
// float64 oracle without a GPU.  Not part of the product and not a fallback.
#include <cstddef>
#include "../../aidet_b200/csrc/geom.cuh"
using namespace aidet;

// mirrors PairOp<K>::overlap of riou.cu
extern "C" void sim_riou_matrix(const float* a, int m, const float* b, int n, int fmt, int mode, float* out) {
  for (int i = 0; i < m; i++) {
    for (int j = 0; j < n; j++) {
      float v;
      if (fmt == 5) {
        RectRow ra, rb; RectCol ca, cb;
        rect_prepare(a + 5 * (size_t)i, &ra, &ca); rect_prepare(b + 5 * (size_t)j, &rb, &cb);
        float dx = ra.cx - rb.cx, dy = ra.cy - rb.cy, rr = ra.rad + rb.rad;
        if (dx * dx + dy * dy > rr * rr) v = 0.f;
        else v = finish_overlap(rect_inter(ra, cb), ra.area, rb.area, mode);
      } else {
        QuadRow ra, rb; QuadCol ca, cb;
        quad_prepare(a + 8 * (size_t)i, &ra, &ca); quad_prepare(b + 8 * (size_t)j, &rb, &cb);
        float dx = ra.mx - rb.mx, dy = ra.my - rb.my, rr = ra.rad + rb.rad;
        if (dx * dx + dy * dy > rr * rr) v = 0.f;
        else v = finish_overlap(quad_inter(ra, cb), ra.area, rb.area, mode);
      }
      out[(size_t)i * n + j] = v;
    }
  }
}

// mirrors riou_aligned_grad_kernel of riou.cu (theta-OBB): out (n), grad (n,10) = d ov / d (a, b)
extern "C" void sim_riou_aligned_grad(const float* a, const float* b, int n, int mode, float* out, float* grad) {
  for (int i = 0; i < n; i++)
    out[i] = rect_overlap_grad(a + 5 * (size_t)i, b + 5 * (size_t)i, mode, grad + 10 * (size_t)i, grad + 10 * (size_t)i + 5);
}

// point-OBB form: a, b (n,8); grad (n,16) = d ov / d (a corners, b corners)
extern "C" void sim_riou_aligned_grad8(const float* a, const float* b, int n, int mode, float* out, float* grad) {
  for (int i = 0; i < n; i++)
    out[i] = quad_overlap_grad(a + 8 * (size_t)i, b + 8 * (size_t)i, mode, grad + 16 * (size_t)i, grad + 16 * (size_t)i + 8);
}
