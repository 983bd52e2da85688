// geom.cuh -- pairwise overlap arithmetic for oriented boxes (FP32, registers only).
//
// What it replaces: the polygon IoU that AIDet reaches only through the external
// wwtool.mergebypoly_mp (mmdet/datasets/dota.py:336) and the commented
// nms_wrapper.thetaobb_nms slot (mmdet/core/post_processing/rbbox_nms.py:97);
// HBB overlap follows mmdet/ops/nms/src/nms_kernel.cu:14-22 (+1 convention).
//
// Formulation (B200-first, not a translation of Sutherland-Hodgman): clipping
// a polygon A against a box B is written as a line integral over A's edges
// (Green's theorem with Q(x,y) = clamp(x,-W,W) * 1[|y|<=H] in B's frame):
//
//     area(A ^ B) = sum over CCW edges (p -> p+d) of
//                   dy * Integral_{t0}^{t1} clamp(px + t dx, -W, W) dt
//
// where [t0,t1] is the part of the edge inside the slab |y| <= H.  Every term
// is a closed form of min/max/fma: no vertex lists, no data-dependent trip
// counts, no local memory, and no topological decisions that could flip the
// result by O(1) -- each edge integral is continuous in its inputs.
// General (8-point) quads use the same idea per triangle of a fan of B in
// affine coordinates: Q = clamp(xi, 0, 1-eta) * 1[0<=eta<=1].
//
// The header is host+device so tests/ can compile the very same arithmetic
// with g++ and check it against the float64 oracle without a GPU.  It is NOT a
// CPU fallback: the product only ever calls it from kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define AIDET_HD __host__ __device__ __forceinline__
#define AIDET_HDM __host__ __device__ __forceinline__      // member functions
#define AIDET_ALIGN16 __align__(16)
#else
#define AIDET_HD static inline
#define AIDET_HDM inline
#define AIDET_ALIGN16 alignas(16)
#endif

namespace aidet {

#if defined(__CUDA_ARCH__)
AIDET_HD float frcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }  // MUFU.RCP, <= 1 ulp
AIDET_HD float fdiv(float a, float b) { return __fdividef(a, b); }
#else
AIDET_HD float frcp(float x) { return 1.0f / x; }
AIDET_HD float fdiv(float a, float b) { return a / b; }
#endif

// MODE_IOF divides by the area of the first box (mmdet/core/bbox/geometry.py:70-71), MODE_IOF_B by the second's.
enum { MODE_IOU = 0, MODE_IOF = 1, MODE_IOF_B = 2 };

// ---------------------------------------------------------------- records
// theta-OBB, prepared once per box by the prologue kernel (32 B): the record of "B" (the box whose frame is used; a
// kernel's column box).  "A" (the box that is transformed; a row box) is stored as RectA below (48 B: + reciprocal
// edge lengths); callers that hold two Rects (aligned pairs, gradients) derive it per pair (rect_as_row).
struct AIDET_ALIGN16 Rect {     // first 16 B: all the bounding-circle test needs (one LDS.128)
  float cx, cy;                  // centre
  float rad, area;               // circumradius (slightly inflated), w*h
  float c, s;                    // cos, sin of theta
  float W, H;                    // half extents
};
using RectRow = Rect;
using RectCol = Rect;

// point-OBB (general simple quad), 64 B each.
struct AIDET_ALIGN16 QuadRow {
  float x[4], y[4];              // CCW corner offsets about the centroid (mx,my)
  float area, rad, mx, my;       // |area|, bounding radius about (mx,my)
  float ux, uy, vx, vy;          // parallelograms (rectangles given by their corners): half edge vectors, corners =
                                 // centre -+ u -+ v; ux = NaN otherwise (see quad_is_para / para_inter)
};
struct AIDET_ALIGN16 QuadCol {
  float ox, oy;                  // b0
  float e1x, e1y, e2x, e2y, e3x, e3y;   // b1-b0, b2-b0, b3-b0
  float invD1, invD2;            // 1/(e1 x e2), 1/(e2 x e3)   (0 if degenerate)
  float aD1, aD2;                // |e1 x e2|, |e2 x e3|
  float area, rad, mx, my;
};

// axis-aligned box (x1,y1,x2,y2) with the legacy +1 folded into `one`.
struct AIDET_ALIGN16 HbbBox { float x1, y1, x2, y2; };

// ------------------------------------------------- rect ^ rect (theta-OBB)

// clamp to [0, 1]: a modifier of the producing FADD / FMUL / FFMA on the device (.SAT) -- it costs no instruction, where
// a clamp built from min / max costs two on the half-rate ALU pipe.  NaN -> 0 on both sides.
#if defined(__CUDA_ARCH__)
AIDET_HD float sat(float x) { return __saturatef(x); }
#else
AIDET_HD float sat(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
#endif

// Edge integral described in the header comment for the edge p + tau * D, tau in [0, 1] (D = the whole edge vector,
// p in the frame shifted by xref, see rect_inter).  Parametrising by the edge fraction makes every clamp of a parameter
// to the edge a saturation:
//   nrDy = -1 / D.y,  hr = H / |D.y|          -> in-slab range  [t0, t1] = sat(-py / D.y -+ hr)
//   nrDx = -1 / D.x,  k1 = x1st / D.x, k2 ..  -> x crossings    c1, c2 = (x1st - px) / D.x, (x2nd - px) / D.x
//   x1st, x2nd : the box's x bounds in the order this direction meets them;  hdx = D.x / 2
// Inside [t0, t1] the clamp value is x1st before the first crossing (length l_lo), x2nd after the second (l_hi) and x
// itself (taken at the middle of the piece) in between.  Every piece length is a saturated difference of parameters in
// [0, 1], so an empty range contributes exactly 0 without a branch: 2 min/max per edge (r1: 7).
AIDET_HD float rect_edge(float px, float py, float nrDy, float hr, float nrDx, float k1, float k2, float x1st,
                         float x2nd, float hdx) {
  const float t0 = sat(fmaf(py, nrDy, -hr));
  const float t1 = sat(fmaf(py, nrDy, hr));                      // >= t0: hr >= 0 and sat is monotonic
  const float c1 = fmaf(px, nrDx, k1), c2 = fmaf(px, nrDx, k2);
  const float l_lo = sat(fminf(c1, t1) - t0);
  const float l_hi = sat(t1 - fmaxf(c2, t0));
  const float ta = t0 + l_lo, tb = t1 - l_hi;
  const float xmid = fmaf(hdx, ta + tb, px);
  return fmaf(tb - ta, xmid, fmaf(l_hi, x2nd, l_lo * x1st));
}

// The box that is transformed ("A", a matrix row), as the overlap-matrix kernels stage it (48 B): full edge lengths and
// their reciprocals are per-box constants, so a pair needs only the two reciprocals of the RELATIVE angle's cos / sin
// (MUFU.RCP occupies its issue port far longer than an FMUL: four per pair cost 6 % of the kernel).
struct AIDET_ALIGN16 RectA {
  float cx, cy, rad, harea;      // centre, circumradius (as Rect), HALF the area: the overlap is a ratio, and the edge
                                 // integrals below come out as half the intersection (half edge vectors)
  float c, s;                    // cos, sin of theta
  float W, H;                    // half extents (>= 1e-9)
  float iLu, iLv;                // 1 / 2W, 1 / 2H
  float pad0, pad1;
};

// The box whose frame is used ("B", a matrix column) as the kernels keep it in REGISTERS for a whole column tile: the
// stored record plus what depends on it alone, derived once per tile instead of once per pair.
struct RectB : Rect {
  float W2, i2W, harea;          // 2W, 1 / 2W, area / 2
  AIDET_HDM RectB() {}
  AIDET_HDM RectB(const Rect& b) : Rect(b) { W2 = b.W + b.W; i2W = 0.5f * frcp(b.W); harea = 0.5f * b.area; }
};

// HALF the intersection area of A with B, in B's frame.
AIDET_HD float rect_inter_half(const RectA& a, const RectB& b) {
  const float relx = a.cx - b.cx, rely = a.cy - b.cy;
  // centre of A and its axis direction (cos, sin of the relative angle) in B's frame.  The tiny addend (applied last, so
  // it cannot be absorbed) keeps both components non-zero -- parallel boxes would give 1/0 -- and is far below f32
  // resolution of any other value.
  const float trx = fmaf(b.c, relx, b.s * rely);
  const float ry = fmaf(b.c, rely, -b.s * relx);
  const float c = fmaf(a.c, b.c, a.s * b.s) + 1e-20f;
  const float s = fmaf(a.s, b.c, -a.c * b.s) + 1e-20f;
  // Shift x by xref ~ clamp(rx, -W, W): the contour integral of a constant times 1[|y|<=H] dy over a closed polygon is
  // 0, so the result does not depend on xref, but every term is now of the order of the SMALLER box, which keeps the
  // rounding error relative to the intersection.  (b.W * (2 sat(rx / 2W + 1/2) - 1): the clamp as a saturation.)
  const float xref = fmaf(b.W2, sat(fmaf(trx, b.i2W, 0.5f)), -b.W);
  const float rx = trx - xref;                    // B now spans [-xref - W, -xref + W] in x
  // half edge vectors hu = W (c, s), hv = H (-s, c); reciprocals of the components of the WHOLE edge vectors 2 hu, 2 hv
  const float rc = frcp(c), rs = frcp(s);
  const float hux = a.W * c, huy = a.W * s, hvx = -a.H * s, hvy = a.H * c;
  const float rux = a.iLu * rc, ruy = a.iLu * rs, rvx = -a.iLv * rs, rvy = a.iLv * rc;
  // corners p0 = r - hu - hv, p1 = p0 + 2 hu, p3 = p0 + 2 hv  (CCW: p0,p1,p2,p3)
  const float mx = rx - hux, my = ry - huy;
  const float p0x = mx - hvx, p0y = my - hvy;
  const float p3x = mx + hvx, p3y = my + hvy;
  const float p1x = fmaf(2.0f, hux, p0x), p1y = fmaf(2.0f, huy, p0y);
  // B's x bounds in the order each direction meets them, and their crossing parameters relative to -px / D.x
  const float ws_u = copysignf(b.W, hux), ws_v = copysignf(b.W, hvx);
  const float x1u = -xref - ws_u, x2u = ws_u - xref, x1v = -xref - ws_v, x2v = ws_v - xref;
  const float hr_u = b.H * fabsf(ruy), k1u = x1u * rux, k2u = x2u * rux;
  const float hr_v = b.H * fabsf(rvy), k1v = x1v * rvx, k2v = x2v * rvx;
  // edges p0->p1 (+2hu), p1->p2 (+2hv), p2->p3 == -(p3->p2, +2hu), p3->p0 == -(p0->p3, +2hv)
  const float iu = rect_edge(p0x, p0y, -ruy, hr_u, -rux, k1u, k2u, x1u, x2u, hux)
                 - rect_edge(p3x, p3y, -ruy, hr_u, -rux, k1u, k2u, x1u, x2u, hux);
  const float iv = rect_edge(p1x, p1y, -rvy, hr_v, -rvx, k1v, k2v, x1v, x2v, hvx)
                 - rect_edge(p0x, p0y, -rvy, hr_v, -rvx, k1v, k2v, x1v, x2v, hvx);
  return fmaf(huy, iu, hvy * iv);                 // half of dy over the whole edge
}

// A given as a stored Rect (NMS, aligned pairs, gradients): the row constants are derived per pair
AIDET_HD RectA rect_as_row(const Rect& r) {
  RectA a;
  a.cx = r.cx; a.cy = r.cy; a.rad = r.rad; a.harea = 0.5f * r.area; a.c = r.c; a.s = r.s;
  a.W = r.W; a.H = r.H; a.iLu = 0.5f * frcp(r.W); a.iLv = 0.5f * frcp(r.H);
  a.pad0 = a.pad1 = 0.0f;
  return a;
}
AIDET_HD float rect_inter(const Rect& a, const Rect& b) { const float h = rect_inter_half(rect_as_row(a), RectB(b)); return h + h; }

// den is a box area or a union of two: either 0 (degenerate boxes -> overlap 0) or far above
// the denormal range, so the plain MUFU.RCP (<= 1 ulp) replaces a guarded division.
AIDET_HD float finish_overlap(float inter, float area_a, float area_b, int mode) {
  inter = fminf(inter, fminf(area_a, area_b));
  const float den = (mode == MODE_IOF) ? area_a : (mode == MODE_IOF_B) ? area_b : (area_a + area_b - inter);
  // the final saturation is the lower clamp of `inter` (rounding can leave it a hair below 0) and the guard of
  // den == 0 (then inter <= 0: 0 * inf = NaN -> 0, negative * inf = -inf -> 0)
  return sat(inter * frcp(den));
}

AIDET_HD float rect_overlap(const Rect& a, const Rect& b, int mode) {
  float dx = a.cx - b.cx, dy = a.cy - b.cy, r = a.rad + b.rad;
  if (fmaf(dx, dx, dy * dy) > r * r) return 0.0f;       // disjoint bounding circles
  return finish_overlap(rect_inter(a, b), a.area, b.area, mode);
}

// prologue math (runs once per box; double keeps sin/cos at <= 0.5 ulp)
AIDET_HD void rect_prepare(const float* box5, Rect* out) {
  float w = fabsf(box5[2]), h = fabsf(box5[3]);
  double th = (double)box5[4];
  float W = 0.5f * w, H = 0.5f * h;
  out->cx = box5[0]; out->cy = box5[1];
  out->c = (float)cos(th); out->s = (float)sin(th);
  out->W = fmaxf(W, 1e-9f); out->H = fmaxf(H, 1e-9f);     // the edge vectors stay invertible (rect_inter); the area does not move
  out->area = w * h;
  out->rad = sqrtf(W * W + H * H) * 1.000001f + 1e-6f;
}
AIDET_HD void rect_prepare(const float* box5, RectA* row, Rect* col) {
  Rect r; rect_prepare(box5, &r);
  if (row) { *row = rect_as_row(r); row->iLu = 0.5f / r.W; row->iLv = 0.5f / r.H; }
  if (col) *col = r;
}
AIDET_HD void rect_prepare(const float* box5, Rect* row, Rect* col) {
  Rect r; rect_prepare(box5, &r);
  if (row) *row = r;
  if (col) *col = r;
}

// ------------------------------------- d(overlap)/d(box parameters), theta-OBB
//
// Used by the rotated IoU loss (the rotated counterpart of mmdet/models/losses/iou_loss.py:10-27, whose HBB
// form gets its gradient from autograd through bbox_overlaps).  By the transport theorem the derivative of
// the intersection area with respect to a parameter p of box A is the boundary integral, over the parts of
// A's edges that lie inside B, of the normal velocity of the edge under p:
//     translation : n                      -> sum over edges of (inside length) * n
//     w (h)       : 1/2 on the two edges perpendicular to the w (h) axis, 0 on the others
//     theta       : (p - centre) x n       -> first moment of the inside interval along the edge
// so everything reduces to clipping A's four edges against B (two slabs in B's frame) -- closed forms of
// min/max again, exact wherever the overlap is differentiable.

// [t0,t1] = part of the segment p + t d, t in [0,L], inside |x|<=W, |y|<=H (rdx, rdy = 1/d.x, 1/d.y, never 0/inf
// thanks to the 1e-20 addend of the caller).  Empty -> t1 == t0.
AIDET_HD void seg_clip(float px, float py, float rdx, float rdy, float L, float W, float H, float* t0, float* t1) {
  float xa = (-W - px) * rdx, xb = (W - px) * rdx;
  float ya = (-H - py) * rdy, yb = (H - py) * rdy;
  float lo = fmaxf(fmaxf(fminf(xa, xb), fminf(ya, yb)), 0.0f);
  float hi = fminf(fminf(fmaxf(xa, xb), fmaxf(ya, yb)), L);
  *t0 = lo; *t1 = fmaxf(hi, lo);
}

// g[0..4] = d area(A ^ B) / d (cx, cy, w, h, theta) of A   (w, h >= 0 as stored in the record)
AIDET_HD void rect_inter_grad(const Rect& a, const Rect& b, float* g) {
  const float relx = a.cx - b.cx, rely = a.cy - b.cy;
  const float rx = fmaf(b.c, relx, b.s * rely);
  const float ry = fmaf(b.c, rely, -b.s * relx);
  const float c = fmaf(a.c, b.c, a.s * b.s) + 1e-20f;       // A's axes in B's frame: u = (c, s), v = (-s, c)
  const float s = fmaf(a.s, b.c, -a.c * b.s) + 1e-20f;
  const float rc = frcp(c), rs = frcp(s);
  const float ux = a.W * c, uy = a.W * s, vx = -a.H * s, vy = a.H * c;
  const float Lu = a.W + a.W, Lv = a.H + a.H;
  float t0, t1;
  // edges x_local = +W / -W run along v from y_local = -H
  seg_clip(rx + ux - vx, ry + uy - vy, -rs, rc, Lv, b.W, b.H, &t0, &t1);
  const float len_up = t1 - t0, mom_up = len_up * (0.5f * (t0 + t1) - a.H);
  seg_clip(rx - ux - vx, ry - uy - vy, -rs, rc, Lv, b.W, b.H, &t0, &t1);
  const float len_um = t1 - t0, mom_um = len_um * (0.5f * (t0 + t1) - a.H);
  // edges y_local = +H / -H run along u from x_local = -W
  seg_clip(rx - ux + vx, ry - uy + vy, rc, rs, Lu, b.W, b.H, &t0, &t1);
  const float len_vp = t1 - t0, mom_vp = len_vp * (0.5f * (t0 + t1) - a.W);
  seg_clip(rx - ux - vx, ry - uy - vy, rc, rs, Lu, b.W, b.H, &t0, &t1);
  const float len_vm = t1 - t0, mom_vm = len_vm * (0.5f * (t0 + t1) - a.W);
  const float du = len_up - len_um, dv = len_vp - len_vm;
  const float gx = c * du - s * dv, gy = s * du + c * dv;     // in B's frame
  g[0] = b.c * gx - b.s * gy;                                 // back to the world frame
  g[1] = b.s * gx + b.c * gy;
  g[2] = 0.5f * (len_up + len_um);
  g[3] = 0.5f * (len_vp + len_vm);
  g[4] = (mom_um - mom_up) + (mom_vp - mom_vm);
}

// Overlap of two theta-OBBs and its gradient with respect to both boxes' (cx, cy, w, h, theta).
// a5 / b5: the raw parameters (sign of w, h is honoured: the area uses |w|, |h|).  Returns the overlap.
AIDET_HD float rect_overlap_grad(const float* a5, const float* b5, int mode, float* ga, float* gb) {
  Rect a, b;
  rect_prepare(a5, &a); rect_prepare(b5, &b);
#pragma unroll
  for (int k = 0; k < 5; k++) { ga[k] = 0.0f; gb[k] = 0.0f; }
  float dx = a.cx - b.cx, dy = a.cy - b.cy, r = a.rad + b.rad;
  if (fmaf(dx, dx, dy * dy) > r * r) return 0.0f;
  float inter = rect_inter(a, b);
  inter = fminf(fmaxf(inter, 0.0f), fminf(a.area, b.area));
  const float wa = a.W + a.W, ha = a.H + a.H, wb = b.W + b.W, hb = b.H + b.H;
  float ia[5], ib[5];
  rect_inter_grad(a, b, ia);
  rect_inter_grad(b, a, ib);
  // overlap = inter / den:  d = (dI * P - I * dQ) * Rr  with  (P, Q, Rr) per mode
  //   iou   : den = S - I, S = area_a + area_b  ->  (dI * S - I * dS) / den^2
  //   iof   : den = area_a                      ->  (dI * area_a - I * d area_a) / area_a^2
  const float S = a.area + b.area;
  const float den = (mode == MODE_IOF) ? a.area : (mode == MODE_IOF_B) ? b.area : (S - inter);
  if (!(den > 0.0f)) return 0.0f;
  const float rden = 1.0f / den, r2 = rden * rden;
  const float P = (mode == MODE_IOU) ? S : den;
  const float qa = (mode == MODE_IOF_B) ? 0.0f : inter;      // weight of d area_a in dQ
  const float qb = (mode == MODE_IOF) ? 0.0f : inter;        // weight of d area_b
  const float dAa[5] = {0.0f, 0.0f, ha, wa, 0.0f}, dAb[5] = {0.0f, 0.0f, hb, wb, 0.0f};
#pragma unroll
  for (int k = 0; k < 5; k++) {
    ga[k] = (ia[k] * P - qa * dAa[k]) * r2;
    gb[k] = (ib[k] * P - qb * dAb[k]) * r2;
  }
  if (a5[2] < 0.0f) ga[2] = -ga[2];
  if (a5[3] < 0.0f) ga[3] = -ga[3];
  if (b5[2] < 0.0f) gb[2] = -gb[2];
  if (b5[3] < 0.0f) gb[3] = -gb[3];
  return inter * rden;
}

// --------------------------------------------- quad ^ quad (point-OBB, general)

// mean over t in [0,1] of max(f0 + t (f1-f0), 0)
AIDET_HD float mean_pos(float f0, float f1) {
  float p0 = fmaxf(f0, 0.0f), p1 = fmaxf(f1, 0.0f);
  float df = f1 - f0;
  float rho = (fabsf(df) > 1e-20f) ? (p1 - p0) * frcp(df) : 1.0f;
  return 0.5f * (p0 + p1) * rho;
}

// Edge integral in a triangle's affine frame: Q = clamp(xi, 0, 1-eta) 1[0<=eta<=1], written in
// coordinates shifted by a constant xi_ref (which leaves the closed contour integral unchanged):
//   X = xi - xi_ref,  lo = -xi_ref,  Hi = (1 - eta) - xi_ref,  clamp(X, lo, Hi) = X - (X-Hi)+ + (lo-X)+
// (Xp,Ep,Hp) / (Xq,Eq,Hq): X, eta, Hi at the edge's end points.
AIDET_HD float tri_edge(float Xp, float Ep, float Hp, float Xq, float Eq, float Hq, float lo) {
  float de = Eq - Ep, dX = Xq - Xp, dH = Hq - Hp;
  float rde = frcp(de);
  float ta = -Ep * rde, tb = (1.0f - Ep) * rde;
  float t0 = fmaxf(fminf(ta, tb), 0.0f);
  float t1 = fminf(fmaxf(ta, tb), 1.0f);
  float X0 = fmaf(t0, dX, Xp), X1 = fmaf(t1, dX, Xp);
  float H0 = fmaf(t0, dH, Hp), H1 = fmaf(t1, dH, Hp);
  float v = 0.5f * (X0 + X1) - mean_pos(X0 - H0, X1 - H1) + mean_pos(lo - X0, lo - X1);
  v = de * (t1 - t0) * v;
  return (t1 > t0 && fabsf(de) > 1e-20f) ? v : 0.0f;
}

// a: corner offsets about its centroid (mx,my); triangle (o, o+e1, o+e2), invD = 1/(e1 x e2), aD = |e1 x e2|
AIDET_HD float quad_tri_inter(const QuadRow& a, float ox, float oy, float e1x, float e1y, float e2x, float e2y,
                              float invD, float aD) {
  float cx = a.mx - ox, cy = a.my - oy;
  float xc = (cx * e2y - cy * e2x) * invD, ec = (e1x * cy - e1y * cx) * invD;
  float xref = fminf(fmaxf(xc, 0.0f), 1.0f);
  float base = xc - xref, lo = -xref, hic = (1.0f - ec) - xref;
  float X[4], E[4], Hh[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    float dxi = (a.x[k] * e2y - a.y[k] * e2x) * invD;      // affine offsets of the corner about the centroid:
    float det = (e1x * a.y[k] - e1y * a.x[k]) * invD;      // small for a small box, so differences stay accurate
    X[k] = base + dxi; E[k] = ec + det; Hh[k] = hic - det;
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int q = (k + 1) & 3;
    s += tri_edge(X[k], E[k], Hh[k], X[q], E[q], Hh[q], lo);
  }
  return aD * s;
}


// ------------------------------------- parallelogram ^ parallelogram (8-point boxes that are rectangles)
//
// Most 8-point boxes ARE rectangles (thetaobb2pointobb / hobb2pointobb output, the DOTA txt rows, mmdet/core/rbbox/
// transforms.py:45-55,137-163).  For two parallelograms the fan of two triangles (8 triangle-edge integrals with three
// reciprocals each, ~550 instructions) is not needed: the affine map that takes B to the square [-1,1]^2 keeps A a
// parallelogram, so area(A ^ B) = |det M_B| * area(A' ^ square) and A' ^ square is the rectangle integral of rect_inter
// with one pair of reciprocals per edge DIRECTION (four in all) instead of the rotation's two.
// A box takes this path when its corners satisfy p0 + p2 == p1 + p3 up to a few float32 ulps of the coordinates; the
// parallelogram used is the least-squares one (each corner moves by a quarter of that residual).

// residual |p0 + p2 - p1 - p3| (L1) against a few float32 ulps of the coordinate magnitude (what rounding the corners of
// an exact rectangle leaves) AND against the box size (the overlap moves by ~0.15 residual / size: 4e-5 keeps it under
// 6e-6 -- boxes whose coordinates are so large that rounding alone exceeds this go through the general path)
AIDET_HD float para_tolerance(float mx, float my, float rad) {
  return fminf(4.8e-7f * (fabsf(mx) + fabsf(my) + rad), 4e-5f * rad);
}
AIDET_HD bool quad_is_para(const float* x, const float* y, float mx, float my, float rad) {
  const float res = fabsf((x[0] + x[2]) - (x[1] + x[3])) + fabsf((y[0] + y[2]) - (y[1] + y[3]));
  return res <= para_tolerance(mx, my, rad);
}

// Edge integral of rect_edge for the edge p + t d, t in [0, 1], against the square |x|,|y| <= 1 shifted by xref in x;
// rdx, rdy = 1/d.x, 1/d.y (never 0 / inf: the caller adds 1e-20 to the components).
AIDET_HD float para_edge(float px, float py, float dx, float rdx, float rdy, float xref) {
  const float sx = copysignf(1.0f, dx);
  const float x1 = -xref - sx, x2 = sx - xref;
  return rect_edge(px, py, -rdy, fabsf(rdy), -rdx, x1 * rdx, x2 * rdx, x1, x2, 0.5f * dx);
}

AIDET_HD bool quad_col_is_para(const QuadCol& b) {
  const float res = fabsf(b.e2x - b.e1x - b.e3x) + fabsf(b.e2y - b.e1y - b.e3y);
  return res <= para_tolerance(b.mx, b.my, b.rad);
}

// Column box as the kernels keep it in REGISTERS for a whole column tile: the stored record plus what the parallelogram
// path needs of it, derived once per tile instead of once per pair.
//   centre = o + (cdx, cdy), half edges uB = (e1 + e2 - e3) / 4, vB = (e3 + e2 - e1) / 4 (least squares over the corners),
//   M^-1 = [vBy -vBx; -uBy uBx] / det  (B = [-1,1]^2 in that frame); m00 is NaN when B is not a parallelogram.
struct QuadReg : QuadCol {
  float m00, m01, m10, m11, det, cdx, cdy;
  AIDET_HDM QuadReg() {}
  AIDET_HDM QuadReg(const QuadCol& b) : QuadCol(b) {
    const float ubx = 0.25f * (b.e1x + b.e2x - b.e3x), uby = 0.25f * (b.e1y + b.e2y - b.e3y);
    const float vbx = 0.25f * (b.e3x + b.e2x - b.e1x), vby = 0.25f * (b.e3y + b.e2y - b.e1y);
    det = ubx * vby - uby * vbx;                              // > 0 (CCW); 0 only for degenerate boxes (area 0 -> overlap 0)
    const float rdet = 1.0f / (det + 1e-30f);
    m00 = quad_col_is_para(b) ? vby * rdet : nanf("");
    m01 = -vbx * rdet; m10 = -uby * rdet; m11 = ubx * rdet;
    cdx = 0.25f * (b.e1x + b.e2x + b.e3x); cdy = 0.25f * (b.e1y + b.e2y + b.e3y);
  }
};

// a, b: both parallelograms (a.ux and b.m00 are not NaN).  Intersection area.
AIDET_HD float para_inter(const QuadRow& a, const QuadReg& b) {
  // centres: the exact mean of the four corners on both sides -- (a.mx, a.my) is that mean ROUNDED (the corner offsets
  // carry the remainder), B's corners are o, o + e1, o + e2, o + e3
  const float relx = (a.mx - b.ox) + (0.25f * ((a.x[0] + a.x[2]) + (a.x[1] + a.x[3])) - b.cdx);
  const float rely = (a.my - b.oy) + (0.25f * ((a.y[0] + a.y[2]) + (a.y[1] + a.y[3])) - b.cdy);
  // A's centre and half edges in B's frame
  float rx = fmaf(b.m00, relx, b.m01 * rely);
  const float ry = fmaf(b.m10, relx, b.m11 * rely);
  const float ux = fmaf(b.m00, a.ux, b.m01 * a.uy) + 1e-20f, uy = fmaf(b.m10, a.ux, b.m11 * a.uy) + 1e-20f;
  const float vx = fmaf(b.m00, a.vx, b.m01 * a.vy) + 1e-20f, vy = fmaf(b.m10, a.vx, b.m11 * a.vy) + 1e-20f;
  const float xref = fminf(fmaxf(rx, -1.0f), 1.0f);           // same shift as rect_inter: terms of the order of the smaller box
  rx -= xref;
  const float mx = rx - ux, my = ry - uy;
  const float p0x = mx - vx, p0y = my - vy;                    // corners p0 = r-u-v, p1 = r+u-v, p3 = r-u+v (CCW)
  const float p3x = mx + vx, p3y = my + vy;
  const float dux = ux + ux, duy = uy + uy, dvx = vx + vx, dvy = vy + vy;
  const float p1x = p0x + dux, p1y = p0y + duy;
  const float rux = frcp(dux), ruy = frcp(duy), rvx = frcp(dvx), rvy = frcp(dvy);
  // edges p0->p1 (+du), p1->p2 (+dv), p2->p3 == -(p3->p2, +du), p3->p0 == -(p0->p3, +dv)
  const float iu = para_edge(p0x, p0y, dux, rux, ruy, xref) - para_edge(p3x, p3y, dux, rux, ruy, xref);
  const float iv = para_edge(p1x, p1y, dvx, rvx, rvy, xref) - para_edge(p0x, p0y, dvx, rvx, rvy, xref);
  return b.det * fmaf(duy, iu, dvy * iv);
}

AIDET_HD float quad_fan_inter(const QuadRow& a, const QuadCol& b) {
  return quad_tri_inter(a, b.ox, b.oy, b.e1x, b.e1y, b.e2x, b.e2y, b.invD1, b.aD1)
       + quad_tri_inter(a, b.ox, b.oy, b.e2x, b.e2y, b.e3x, b.e3y, b.invD2, b.aD2);
}

// register-resident column (kernels): the parallelogram test of the column was made when the record was loaded
AIDET_HD float quad_inter(const QuadRow& a, const QuadReg& b) {
  if (a.ux == a.ux && b.m00 == b.m00) return para_inter(a, b);            // both parallelograms (NaN otherwise)
  return quad_fan_inter(a, b);
}
// stored column (aligned pairs, gradients, host simulation)
AIDET_HD float quad_inter(const QuadRow& a, const QuadCol& b) { return quad_inter(a, QuadReg(b)); }

AIDET_HD float quad_overlap(const QuadRow& a, const QuadCol& b, int mode) {
  float dx = a.mx - b.mx, dy = a.my - b.my, r = a.rad + b.rad;
  if (fmaf(dx, dx, dy * dy) > r * r) return 0.0f;
  return finish_overlap(quad_inter(a, b), a.area, b.area, mode);
}

AIDET_HD void quad_prepare(const float* box8, QuadRow* row, QuadCol* col) {
  float x[4], y[4];
  for (int k = 0; k < 4; k++) { x[k] = box8[2 * k]; y[k] = box8[2 * k + 1]; }
  // signed area about p0 (translation keeps the products small)
  float ax = x[1] - x[0], ay = y[1] - y[0], bx = x[2] - x[0], by = y[2] - y[0], cx = x[3] - x[0], cy = y[3] - y[0];
  float sa = 0.5f * ((ax * by - ay * bx) + (bx * cy - by * cx));
  if (sa < 0.0f) {  // make CCW: swap p1 <-> p3
    float t = x[1]; x[1] = x[3]; x[3] = t; t = y[1]; y[1] = y[3]; y[3] = t;
    t = ax; ax = cx; cx = t; t = ay; ay = cy; cy = t;
  }
  float area = fabsf(sa);
  float mx = 0.25f * (x[0] + x[1] + x[2] + x[3]), my = 0.25f * (y[0] + y[1] + y[2] + y[3]);
  float r2 = 0.0f;
  for (int k = 0; k < 4; k++) { float dx = x[k] - mx, dy = y[k] - my; r2 = fmaxf(r2, dx * dx + dy * dy); }
  float rad = sqrtf(r2) * 1.000001f + 1e-5f * (fabsf(mx) + fabsf(my)) + 1e-6f;
  if (row) {
    for (int k = 0; k < 4; k++) { row->x[k] = x[k] - mx; row->y[k] = y[k] - my; }
    row->area = area; row->rad = rad; row->mx = mx; row->my = my;
    if (quad_is_para(row->x, row->y, mx, my, rad)) {               // least-squares half edges
      row->ux = 0.25f * ((row->x[1] - row->x[0]) + (row->x[2] - row->x[3])); row->uy = 0.25f * ((row->y[1] - row->y[0]) + (row->y[2] - row->y[3]));
      row->vx = 0.25f * ((row->x[3] - row->x[0]) + (row->x[2] - row->x[1])); row->vy = 0.25f * ((row->y[3] - row->y[0]) + (row->y[2] - row->y[1]));
    } else {
      row->ux = nanf(""); row->uy = row->vx = row->vy = 0.0f;
    }
  }
  if (col) {
    col->ox = x[0]; col->oy = y[0];
    col->e1x = ax; col->e1y = ay; col->e2x = bx; col->e2y = by; col->e3x = cx; col->e3y = cy;
    float D1 = ax * by - ay * bx, D2 = bx * cy - by * cx;
    col->invD1 = (fabsf(D1) > 1e-20f) ? 1.0f / D1 : 0.0f; col->aD1 = (fabsf(D1) > 1e-20f) ? fabsf(D1) : 0.0f;
    col->invD2 = (fabsf(D2) > 1e-20f) ? 1.0f / D2 : 0.0f; col->aD2 = (fabsf(D2) > 1e-20f) ? fabsf(D2) : 0.0f;
    col->area = area; col->rad = rad; col->mx = mx; col->my = my;
  }
}

// ------------------------------------- d(overlap)/d(corner coordinates), point-OBB (convex quads)
//
// Same transport-theorem form as rect_inter_grad: moving corner a_k moves its two edges; a point at fraction t of
// edge (a_k -> a_k+1) moves with (1 - t) da_k + t da_k+1, the outward normal times the edge length is (d.y, -d.x), so
//     dI/da_k   += (d.y, -d.x) * [ (t1 - t0) - (t1^2 - t0^2) / 2 ]        [t0, t1] = part of the edge inside B
//     dI/da_k+1 += (d.y, -d.x) *   (t1^2 - t0^2) / 2
// The interval comes from the four half-planes of B (CCW, convex) -- min/max closed forms again.

// x, y: CCW corners of A; bx, by: CCW corners of convex B.  g[2k], g[2k+1] = d area(A ^ B) / d (x_k, y_k).
AIDET_HD void quad_inter_grad(const float* x, const float* y, const float* bx, const float* by, float* g) {
#pragma unroll
  for (int k = 0; k < 8; k++) g[k] = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int q = (k + 1) & 3;
    const float dx = x[q] - x[k], dy = y[q] - y[k];
    float lo = 0.0f, hi = 1.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int jq = (j + 1) & 3;
      const float ex = bx[jq] - bx[j], ey = by[jq] - by[j];
      const float f0 = ex * (y[k] - by[j]) - ey * (x[k] - bx[j]);        // cross(e, a_k - b_j): >= 0 inside
      const float f1 = ex * dy - ey * dx;                                 // its slope along the edge
      const float tc = -f0 * frcp(f1 + copysignf(1e-30f, f1));
      // f1 > 0: inside for t >= tc;  f1 < 0: inside for t <= tc;  f1 ~ 0: the whole edge is in or out with f0
      const bool flat = fabsf(f1) <= 1e-12f * (fabsf(ex) + fabsf(ey)) * (fabsf(dx) + fabsf(dy));
      if (flat) { if (f0 < 0.0f) hi = -1.0f; }
      else if (f1 > 0.0f) lo = fmaxf(lo, tc);
      else hi = fminf(hi, tc);
    }
    hi = fmaxf(hi, lo);
    const float len = hi - lo, half2 = 0.5f * len * (hi + lo);          // (t1^2 - t0^2) / 2, exactly 0 for an empty interval
    const float wk = len - half2, wq = half2;
    g[2 * k] += dy * wk;     g[2 * k + 1] -= dx * wk;
    g[2 * q] += dy * wq;     g[2 * q + 1] -= dx * wq;
  }
}

// Overlap of two convex quads (x1,y1,...,x4,y4, either orientation) and its gradient w.r.t. all 16 coordinates.
AIDET_HD float quad_overlap_grad(const float* a8, const float* b8, int mode, float* ga, float* gb) {
  QuadRow ra; QuadCol cb; QuadRow rb;
  quad_prepare(a8, &ra, (QuadCol*)nullptr);
  quad_prepare(b8, &rb, &cb);
#pragma unroll
  for (int k = 0; k < 8; k++) { ga[k] = 0.0f; gb[k] = 0.0f; }
  float ddx = ra.mx - rb.mx, ddy = ra.my - rb.my, rr = ra.rad + rb.rad;
  if (fmaf(ddx, ddx, ddy * ddy) > rr * rr) return 0.0f;
  float inter = quad_inter(ra, cb);
  inter = fminf(fmaxf(inter, 0.0f), fminf(ra.area, rb.area));
  // CCW corners about a common origin (the centroid of A keeps the products small)
  float ax[4], ay[4], bxx[4], byy[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    ax[k] = ra.x[k]; ay[k] = ra.y[k];
    bxx[k] = rb.x[k] + (rb.mx - ra.mx); byy[k] = rb.y[k] + (rb.my - ra.my);
  }
  float ia[8], ib[8];
  quad_inter_grad(ax, ay, bxx, byy, ia);
  quad_inter_grad(bxx, byy, ax, ay, ib);
  const float S = ra.area + rb.area;
  const float den = (mode == MODE_IOF) ? ra.area : (mode == MODE_IOF_B) ? rb.area : (S - inter);
  if (!(den > 0.0f)) return 0.0f;
  const float rden = 1.0f / den, r2 = rden * rden;
  const float P = (mode == MODE_IOU) ? S : den;
  const float qa = (mode == MODE_IOF_B) ? 0.0f : inter, qb = (mode == MODE_IOF) ? 0.0f : inter;
  // quad_prepare made both quads CCW by swapping corners 1 and 3 when needed: undo that for the caller's order
  const float sa = (a8[2] - a8[0]) * (a8[5] - a8[1]) - (a8[3] - a8[1]) * (a8[4] - a8[0])
                 + (a8[4] - a8[0]) * (a8[7] - a8[1]) - (a8[5] - a8[1]) * (a8[6] - a8[0]);
  const float sb = (b8[2] - b8[0]) * (b8[5] - b8[1]) - (b8[3] - b8[1]) * (b8[4] - b8[0])
                 + (b8[4] - b8[0]) * (b8[7] - b8[1]) - (b8[5] - b8[1]) * (b8[6] - b8[0]);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int kp = (k + 1) & 3, km = (k + 3) & 3;
    // d area / d corner k of a CCW polygon: ( (y_k+1 - y_k-1) / 2, (x_k-1 - x_k+1) / 2 )
    const float dAax = 0.5f * (ay[kp] - ay[km]), dAay = 0.5f * (ax[km] - ax[kp]);
    const float dAbx = 0.5f * (byy[kp] - byy[km]), dAby = 0.5f * (bxx[km] - bxx[kp]);
    const int ka = (sa < 0.0f && (k & 1)) ? (k ^ 2) : k;      // 1 <-> 3
    const int kb = (sb < 0.0f && (k & 1)) ? (k ^ 2) : k;
    ga[2 * ka] = (ia[2 * k] * P - qa * dAax) * r2;  ga[2 * ka + 1] = (ia[2 * k + 1] * P - qa * dAay) * r2;
    gb[2 * kb] = (ib[2 * k] * P - qb * dAbx) * r2;  gb[2 * kb + 1] = (ib[2 * k + 1] * P - qb * dAby) * r2;
  }
  return inter * rden;
}

// ------------------------------------------------------------- HBB (+1)

// mmdet/ops/nms/src/nms_kernel.cu:14-22 (devIoU) / nms_cpu.cpp:47-55
AIDET_HD float hbb_overlap(const HbbBox& a, const HbbBox& b, float one, int mode) {
  float left = fmaxf(a.x1, b.x1), right = fminf(a.x2, b.x2);
  float top = fmaxf(a.y1, b.y1), bottom = fminf(a.y2, b.y2);
  float w = fmaxf(right - left + one, 0.f), h = fmaxf(bottom - top + one, 0.f);
  float inter = w * h;
  float sa = (a.x2 - a.x1 + one) * (a.y2 - a.y1 + one);
  float sb = (b.x2 - b.x1 + one) * (b.y2 - b.y1 + one);
  float den = (mode == MODE_IOF) ? sa : (mode == MODE_IOF_B) ? sb : (sa + sb - inter);
  return inter / den;
}

}  // namespace aidet
