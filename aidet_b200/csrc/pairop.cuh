// pairop.cuh -- box "kinds" (theta-OBB, point-OBB, HBB), their prepared records and the pair operator shared by the
// overlap-matrix kernel (riou.cu) and the fused max-IoU assignment kernels (riou_assign.cu).
#pragma once
#include "common.cuh"
#include "geom.cuh"

namespace aidet {

// Row / Col: the prepared records as stored (staged in shared memory / read from global memory); Reg: the column record as
// a kernel keeps it in registers for a whole column tile (constructible from Col; may carry derived values).
struct RectKind {
  using Row = RectA; using Col = RectCol; using Reg = RectB;     // rows carry their per-box constants (48 B)
  static constexpr int FMT = 5;
  // inter / area_row / area_col: on a common scale (here: halves -- the edge integrals of geom.cuh come out halved)
  __device__ static __forceinline__ float inter(const Row& a, const Reg& b) { return rect_inter_half(a, b); }
  __device__ static __forceinline__ float area_row(const Row& a) { return a.harea; }
  __device__ static __forceinline__ float area_col(const Reg& b) { return b.harea; }
  __device__ static __forceinline__ void prepare(const float* p, Row* r, Col* c) { rect_prepare(p, r, c); }
  template <class T> __device__ static __forceinline__ float cx(const T& t) { return t.cx; }
  template <class T> __device__ static __forceinline__ float cy(const T& t) { return t.cy; }
};
struct QuadKind {
  using Row = QuadRow; using Col = QuadCol; using Reg = QuadReg;
  static constexpr int FMT = 8;
  __device__ static __forceinline__ float inter(const Row& a, const Reg& b) { return quad_inter(a, b); }
  __device__ static __forceinline__ float area_row(const Row& a) { return a.area; }
  __device__ static __forceinline__ float area_col(const Reg& b) { return b.area; }
  __device__ static __forceinline__ void prepare(const float* p, Row* r, Col* c) { quad_prepare(p, r, c); }
  template <class T> __device__ static __forceinline__ float cx(const T& t) { return t.mx; }
  template <class T> __device__ static __forceinline__ float cy(const T& t) { return t.my; }
};

// fmt 4: axis-aligned (x1,y1,x2,y2) boxes with the legacy +1 pixel convention of mmdet/core/bbox/geometry.py:57-86
// (bbox_overlaps) -- the same tiled kernel; this one is bound by the 4 B/pair result store, not by arithmetic.
struct HbbKind {
  using Row = HbbBox; using Col = HbbBox; using Reg = HbbBox;
  static constexpr int FMT = 4;
  __device__ static __forceinline__ void prepare(const float* p, Row* r, Col* c) {
    HbbBox b{p[0], p[1], p[2], p[3]};
    if (r) *r = b;
    if (c) *c = b;
  }
};

// Matrix-row boxes are staged as Row records (the box that is transformed), matrix-column
// boxes live in registers as Col records (the box whose frame is used).
template <class K>
struct PairOp {
  using S = typename K::Row;   // staged (matrix row)
  using R = typename K::Col;   // matrix col as stored
  using X = typename K::Reg;   // matrix col in registers
  // bounding circles meet (the early-out test of overlap())
  __device__ static __forceinline__ bool near(const S& s, const X& r) {
    float dx = K::cx(s) - K::cx(r), dy = K::cy(s) - K::cy(r), rr = s.rad + r.rad;
    return !(fmaf(dx, dx, dy * dy) > rr * rr);
  }
  __device__ static __forceinline__ float overlap(const S& s, const X& r, int mode) {
    float dx = K::cx(s) - K::cx(r), dy = K::cy(s) - K::cy(r), rr = s.rad + r.rad;
    if (fmaf(dx, dx, dy * dy) > rr * rr) return 0.0f;
    return finish_overlap(K::inter(s, r), K::area_row(s), K::area_col(r), mode);
  }
};

template <>
struct PairOp<HbbKind> {
  using S = HbbBox; using R = HbbBox; using X = HbbBox;
  __device__ static __forceinline__ float overlap(const S& s, const R& r, int mode) { return hbb_overlap(s, r, 1.0f, mode); }
};

template <class K>
__global__ void __launch_bounds__(256) riou_prepare_kernel(const float* __restrict__ boxes, int n,
                                                           typename K::Row* rows, typename K::Col* cols) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float b[K::FMT];
#pragma unroll
  for (int k = 0; k < K::FMT; k++) b[k] = boxes[(size_t)i * K::FMT + k];
  typename K::Row r; typename K::Col c;
  K::prepare(b, rows ? &r : nullptr, cols ? &c : nullptr);
  if (rows) rows[i] = r;
  if (cols) cols[i] = c;
}

// both sides of a matrix in ONE launch: threads [0, m) prepare the row records of `a`, threads [m, m + n) the column
// records of `b` (small problems are launch bound: C1 is a 2000 x 2000 matrix)
template <class K>
__global__ void __launch_bounds__(256) riou_prepare_both_kernel(const float* __restrict__ a, int m, const float* __restrict__ b,
                                                                int n, typename K::Row* rows, typename K::Col* cols) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m + n) return;
  const bool is_row = i < m;
  const float* src = is_row ? a + (size_t)i * K::FMT : b + (size_t)(i - m) * K::FMT;
  float v[K::FMT];
#pragma unroll
  for (int k = 0; k < K::FMT; k++) v[k] = src[k];
  typename K::Row r; typename K::Col c;
  K::prepare(v, is_row ? &r : nullptr, is_row ? nullptr : &c);
  if (is_row) rows[i] = r; else cols[i - m] = c;
}

// upper bound of a prepared record (row records of theta-OBBs are 48 B, column records 32 B)
static inline size_t record_bytes(int fmt) { return (fmt == 8) ? 64 : (fmt == 4 ? 16 : 48); }

}  // namespace aidet
