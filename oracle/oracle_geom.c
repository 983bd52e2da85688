/*
 * oracle_geom.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, float64).
 *
 * A plain-C restatement of the oriented-bounding-box arithmetic that AIDet's
 * hot path relies on.  Nothing under aidet_b200/ may import, link or call this
 * file: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker or the timed CPU
 * baseline.
 *
 * PARITY STATUS: "parity unpinned" for the rotated arithmetic.  The reference
 * tree holds no rotated IoU / polygon NMS: the only call sites are
 *   mmdet/datasets/dota.py:23,334,336  (wwtool.mergebypoly_mp / mergebyrec_mp)
 * and wwtool (github jwwangchn/wwtool, version unpinned, un-vendored) derives
 * its polygon IoU from the public DOTA_devkit `polyiou` (signed triangle-fan
 * clipping in double).  Both published algorithms are restated here
 * independently (fan + Sutherland-Hodgman) and cross-checked against each
 * other, against analytic known answers and against cv2's
 * rotatedRectangleIntersection (tests/test_oracle.py).
 *
 * PINNED parts (reference file:line each function follows):
 *   oracle_nms_hbb      mmdet/ops/nms/src/nms_cpu.cpp:5-60   (greedy, +1 areas,
 *                       `>=` on CPU; `>` mirrors nms_kernel.cu:61)
 *   oracle_hbb_overlaps mmdet/core/bbox/geometry.py:57-86    (+1 convention)
 *   box conventions     mmdet/core/rbbox/transforms.py:45-55 (cv2.boxPoints)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y; } pt;

/* ------------------------------------------------------------------ boxes */

/* (cx,cy,w,h,theta[rad]) -> 4 corners in cv2.boxPoints order
 * (mmdet/core/rbbox/transforms.py:45-55): p0 = c + R(-w/2,+h/2), p1 = c +
 * R(-w/2,-h/2), p2 = 2c - p0, p3 = 2c - p1, R = [[cos,-sin],[sin,cos]]. */
static void thetaobb_to_quad(const float* b, pt* q) {
  double cx = b[0], cy = b[1], w = fabs((double)b[2]), h = fabs((double)b[3]);
  double th = b[4];
  double c = cos(th) * 0.5, s = sin(th) * 0.5;
  q[0].x = cx - s * h - c * w; q[0].y = cy + c * h - s * w;
  q[1].x = cx + s * h - c * w; q[1].y = cy - c * h - s * w;
  q[2].x = 2 * cx - q[0].x;    q[2].y = 2 * cy - q[0].y;
  q[3].x = 2 * cx - q[1].x;    q[3].y = 2 * cy - q[1].y;
}

static void load_quad(const float* b, int fmt, pt* q) {
  if (fmt == 5) { thetaobb_to_quad(b, q); return; }
  for (int i = 0; i < 4; i++) { q[i].x = b[2 * i]; q[i].y = b[2 * i + 1]; }
}

void oracle_thetaobb2pointobb(const float* boxes, int n, double* out8) {
  for (int i = 0; i < n; i++) {
    pt q[4]; thetaobb_to_quad(boxes + 5 * i, q);
    for (int k = 0; k < 4; k++) { out8[8 * i + 2 * k] = q[k].x; out8[8 * i + 2 * k + 1] = q[k].y; }
  }
}

static double signed_area(const pt* p, int n) {
  double a = 0;
  for (int i = 0; i < n; i++) {
    int j = (i + 1 == n) ? 0 : i + 1;
    a += p[i].x * p[j].y - p[i].y * p[j].x;
  }
  return 0.5 * a;
}

static void make_ccw(pt* p, int n) {
  if (signed_area(p, n) < 0)
    for (int i = 0, j = n - 1; i < j; i++, j--) { pt t = p[i]; p[i] = p[j]; p[j] = t; }
}

/* ------------------------------------------- Sutherland-Hodgman + shoelace */

static double cross3(pt o, pt a, pt b) {
  return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y);
}

/* keep the part of poly `in` on the left of (or on) the directed line a->b */
static int clip_left(const pt* in, int n, pt a, pt b, pt* out) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    pt p = in[i], q = in[(i + 1 == n) ? 0 : i + 1];
    double dp = cross3(a, b, p), dq = cross3(a, b, q);
    if (dp >= 0) out[m++] = p;
    if ((dp >= 0) != (dq >= 0)) {
      double t = dp / (dp - dq);
      pt r = { p.x + t * (q.x - p.x), p.y + t * (q.y - p.y) };
      out[m++] = r;
    }
  }
  return m;
}

static double inter_area_sh(const pt* qa, const pt* qb) {
  pt a[4], b[4], buf0[16], buf1[16];
  memcpy(a, qa, sizeof a); memcpy(b, qb, sizeof b);
  make_ccw(a, 4); make_ccw(b, 4);
  memcpy(buf0, a, sizeof a);
  int n = 4; pt *cur = buf0, *nxt = buf1;
  for (int e = 0; e < 4 && n > 0; e++) {
    n = clip_left(cur, n, b[e], b[(e + 1) & 3], nxt);
    pt* t = cur; cur = nxt; nxt = t;
  }
  if (n < 3) return 0.0;
  return fabs(signed_area(cur, n));
}

/* ----------------- DOTA_devkit polyiou lineage: signed triangle-fan clipping
 * Restated from the published algorithm (sum over edge pairs of the signed
 * area of tri(O,a_i,a_i+1) ^ tri(O,b_j,b_j+1), eps = 1e-8 sign function). */

#define FAN_EPS 1e-8
static int sgn(double d) { return (d > FAN_EPS) - (d < -FAN_EPS); }

static int line_cross(pt a, pt b, pt c, pt d, pt* p) {
  double s1 = cross3(a, b, c), s2 = cross3(a, b, d);
  if (sgn(s1) == 0 && sgn(s2) == 0) return 2;
  if (sgn(s2 - s1) == 0) return 0;
  p->x = (c.x * s2 - d.x * s1) / (s2 - s1);
  p->y = (c.y * s2 - d.y * s1) / (s2 - s1);
  return 1;
}

static int pt_eq(pt a, pt b) { return sgn(a.x - b.x) == 0 && sgn(a.y - b.y) == 0; }

static void poly_cut(pt* p, int* n, pt a, pt b) {
  pt pp[24]; int m = 0;
  p[*n] = p[0];
  for (int i = 0; i < *n; i++) {
    if (sgn(cross3(a, b, p[i])) > 0) pp[m++] = p[i];
    if (sgn(cross3(a, b, p[i])) != sgn(cross3(a, b, p[i + 1]))) {
      pt r = p[i];
      if (line_cross(a, b, p[i], p[i + 1], &r)) pp[m++] = r;
    }
  }
  int k = 0;
  for (int i = 0; i < m; i++)
    if (!i || !pt_eq(pp[i], pp[i - 1])) p[k++] = pp[i];
  while (k > 1 && pt_eq(p[k - 1], p[0])) k--;
  *n = k;
}

static double tri_inter_signed(pt a, pt b, pt c, pt d) {
  pt o = {0, 0};
  int s1 = sgn(cross3(o, a, b)), s2 = sgn(cross3(o, c, d));
  if (s1 == 0 || s2 == 0) return 0.0;
  if (s1 == -1) { pt t = a; a = b; b = t; }
  if (s2 == -1) { pt t = c; c = d; d = t; }
  pt p[24] = { o, a, b }; int n = 3;
  poly_cut(p, &n, o, c);
  poly_cut(p, &n, c, d);
  poly_cut(p, &n, d, o);
  double res = fabs(signed_area(p, n));
  return (s1 * s2 == -1) ? -res : res;
}

static double inter_area_fan(const pt* qa, const pt* qb) {
  pt a[5], b[5];
  memcpy(a, qa, 4 * sizeof(pt)); memcpy(b, qb, 4 * sizeof(pt));
  make_ccw(a, 4); make_ccw(b, 4);
  /* the fan is taken about the coordinate origin; shift to a local origin so
   * the eps test keeps its meaning at scene-scale coordinates */
  pt o = a[0];
  for (int i = 0; i < 4; i++) { a[i].x -= o.x; a[i].y -= o.y; b[i].x -= o.x; b[i].y -= o.y; }
  a[4] = a[0]; b[4] = b[0];
  double res = 0;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) res += tri_inter_signed(a[i], a[i + 1], b[j], b[j + 1]);
  return res;
}

/* ------------------------------------------------------------- pair IoU */

/* algo: 0 = Sutherland-Hodgman, 1 = triangle fan.  mode: 0 = iou, 1 = iof
 * (inter / area_a, mmdet/core/bbox/geometry.py:70-71 analogue, no +1). */
static double pair_overlap(const pt* qa, const pt* qb, int mode, int algo) {
  double aa = fabs(signed_area(qa, 4)), ab = fabs(signed_area(qb, 4));
  double inter = algo ? inter_area_fan(qa, qb) : inter_area_sh(qa, qb);
  if (inter < 0) inter = 0;
  double den = mode ? aa : (aa + ab - inter);
  if (!(den > 0)) return 0.0;
  return inter / den;
}

double oracle_riou_pair(const float* a, const float* b, int fmt, int mode, int algo) {
  pt qa[4], qb[4]; load_quad(a, fmt, qa); load_quad(b, fmt, qb);
  return pair_overlap(qa, qb, mode, algo);
}

/* theta-OBB pair with DOUBLE parameters (cx,cy,w,h,theta): lets tests/ take central finite differences of the
 * overlap (step 1e-6 px) as the checker for the analytic gradient of the rotated IoU loss
 * (rotated counterpart of mmdet/models/losses/iou_loss.py:10-27).  mode 2 = inter / area_b. */
static void thetaobb_to_quad_d(const double* b, pt* q) {
  double cx = b[0], cy = b[1], w = fabs(b[2]), h = fabs(b[3]);
  double c = cos(b[4]) * 0.5, s = sin(b[4]) * 0.5;
  q[0].x = cx - s * h - c * w; q[0].y = cy + c * h - s * w;
  q[1].x = cx + s * h - c * w; q[1].y = cy - c * h - s * w;
  q[2].x = 2 * cx - q[0].x;    q[2].y = 2 * cy - q[0].y;
  q[3].x = 2 * cx - q[1].x;    q[3].y = 2 * cy - q[1].y;
}

double oracle_riou_pair_d(const double* a, const double* b, int mode, int algo) {
  pt qa[4], qb[4]; thetaobb_to_quad_d(a, qa); thetaobb_to_quad_d(b, qb);
  if (mode == 2) return pair_overlap(qb, qa, 1, algo);
  return pair_overlap(qa, qb, mode, algo);
}

/* central differences of oracle_riou_pair_d: grad (n,10) = d ov / d (a params, b params) */
void oracle_riou_aligned_grad_fd(const double* a, const double* b, int n, int mode, double step, double* ov,
                                 double* grad) {
#pragma omp parallel for
  for (int i = 0; i < n; i++) {
    double p[10];
    for (int k = 0; k < 5; k++) { p[k] = a[5 * (size_t)i + k]; p[5 + k] = b[5 * (size_t)i + k]; }
    ov[i] = oracle_riou_pair_d(p, p + 5, mode, 0);
    for (int k = 0; k < 10; k++) {
      double keep = p[k];
      p[k] = keep + step; double fp = oracle_riou_pair_d(p, p + 5, mode, 0);
      p[k] = keep - step; double fm = oracle_riou_pair_d(p, p + 5, mode, 0);
      p[k] = keep;
      grad[10 * (size_t)i + k] = (fp - fm) / (2 * step);
    }
  }
}

/* the same for point-OBBs with DOUBLE corner coordinates: grad (n,16) */
static double riou_pair_d8(const double* a, const double* b, int mode) {
  pt qa[4], qb[4];
  for (int i = 0; i < 4; i++) { qa[i].x = a[2 * i]; qa[i].y = a[2 * i + 1]; qb[i].x = b[2 * i]; qb[i].y = b[2 * i + 1]; }
  if (mode == 2) return pair_overlap(qb, qa, 1, 0);
  return pair_overlap(qa, qb, mode, 0);
}

void oracle_riou_aligned_grad_fd8(const double* a, const double* b, int n, int mode, double step, double* ov,
                                  double* grad) {
#pragma omp parallel for
  for (int i = 0; i < n; i++) {
    double p[16];
    for (int k = 0; k < 8; k++) { p[k] = a[8 * (size_t)i + k]; p[8 + k] = b[8 * (size_t)i + k]; }
    ov[i] = riou_pair_d8(p, p + 8, mode);
    for (int k = 0; k < 16; k++) {
      double keep = p[k];
      p[k] = keep + step; double fp = riou_pair_d8(p, p + 8, mode);
      p[k] = keep - step; double fm = riou_pair_d8(p, p + 8, mode);
      p[k] = keep;
      grad[16 * (size_t)i + k] = (fp - fm) / (2 * step);
    }
  }
}

void oracle_riou_matrix(const float* a, int m, const float* b, int n, int fmt, int mode, int algo,
                        double* out) {
  pt* qb = (pt*)malloc(sizeof(pt) * 4 * (size_t)(n > 0 ? n : 1));
  for (int j = 0; j < n; j++) load_quad(b + (size_t)fmt * j, fmt, qb + 4 * (size_t)j);
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < m; i++) {
    pt qa[4]; load_quad(a + (size_t)fmt * i, fmt, qa);
    for (int j = 0; j < n; j++) out[(size_t)i * n + j] = pair_overlap(qa, qb + 4 * (size_t)j, mode, algo);
  }
  free(qb);
}

void oracle_riou_aligned(const float* a, const float* b, int n, int fmt, int mode, int algo, double* out) {
#pragma omp parallel for
  for (int i = 0; i < n; i++) out[i] = oracle_riou_pair(a + (size_t)fmt * i, b + (size_t)fmt * i, fmt, mode, algo);
}

/* ----------------------------------------------------- axis-aligned boxes */

static double hbb_overlap(const float* a, const float* b, int mode, double one) {
  double xx1 = fmax(a[0], b[0]), yy1 = fmax(a[1], b[1]);
  double xx2 = fmin(a[2], b[2]), yy2 = fmin(a[3], b[3]);
  double w = fmax(0.0, xx2 - xx1 + one), h = fmax(0.0, yy2 - yy1 + one);
  double inter = w * h;
  double aa = ((double)a[2] - a[0] + one) * ((double)a[3] - a[1] + one);
  double ab = ((double)b[2] - b[0] + one) * ((double)b[3] - b[1] + one);
  double den = mode ? aa : (aa + ab - inter);
  return inter / den;
}

/* mmdet/core/bbox/geometry.py:72-86 */
void oracle_hbb_overlaps(const float* a, int m, const float* b, int n, int mode, int plus_one, double* out) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) out[(size_t)i * n + j] = hbb_overlap(a + 4 * i, b + 4 * j, mode, plus_one ? 1.0 : 0.0);
}

/* ------------------------------------------------------------------- NMS */

typedef struct { float score; int group; int idx; } skey;
static int skey_cmp(const void* pa, const void* pb) {
  const skey* a = (const skey*)pa; const skey* b = (const skey*)pb;
  if (a->group != b->group) return (a->group < b->group) ? -1 : 1;
  if (a->score != b->score) return (a->score > b->score) ? -1 : 1;   /* descending */
  return (a->idx < b->idx) ? -1 : (a->idx > b->idx);                  /* stable on ties */
}

static double pair_any(const float* boxes, int fmt, int i, int j, int plus_one, const pt* quads) {
  if (fmt == 4) return hbb_overlap(boxes + 4 * (size_t)i, boxes + 4 * (size_t)j, 0, plus_one ? 1.0 : 0.0);
  return pair_overlap(quads + 4 * (size_t)i, quads + 4 * (size_t)j, 0, 0);
}

/* Greedy NMS, batched over groups (group = image x class).
 *  - order: group asc, score desc, original index asc on ties (a stable sort,
 *    see SURVEY 8c "score ties")
 *  - suppress j (later in order, same group) when ovr >= thr (cmp_ge=1,
 *    nms_cpu.cpp:56) or ovr > thr (cmp_ge=0, nms_kernel.cu:61 and the
 *    DOTA_devkit merge `ovr <= thresh keeps`)
 *  - keep_out: ascending original index (nms_cpu.cpp:59, nms_kernel.cu:135-138)
 *  - thr: n_thr == 1 -> shared; else one per group id.
 *  - near_out (optional): number of evaluated pairs with |ovr - thr| <= margin
 * fmt: 4 = HBB (x1,y1,x2,y2; plus_one selects the legacy +1), 5 = theta-OBB,
 * 8 = point-OBB.  Returns the number kept. */
int oracle_nms(const float* boxes, int fmt, const float* scores, const int* groups, int n,
               const double* thr, int n_thr, int cmp_ge, int plus_one, double margin,
               int64_t* keep_out, int64_t* near_out) {
  if (n <= 0) { if (near_out) *near_out = 0; return 0; }
  skey* keys = (skey*)malloc(sizeof(skey) * n);
  for (int i = 0; i < n; i++) { keys[i].score = scores[i]; keys[i].group = groups ? groups[i] : 0; keys[i].idx = i; }
  qsort(keys, n, sizeof(skey), skey_cmp);
  pt* quads = NULL;
  if (fmt != 4) {
    quads = (pt*)malloc(sizeof(pt) * 4 * (size_t)n);
    for (int i = 0; i < n; i++) load_quad(boxes + (size_t)fmt * i, fmt, quads + 4 * (size_t)i);
  }
  uint8_t* sup = (uint8_t*)calloc(n, 1);
  int64_t near = 0;
  for (int p = 0; p < n; p++) {
    int i = keys[p].idx;
    if (sup[i]) continue;
    double t = thr[n_thr == 1 ? 0 : keys[p].group];
    for (int q = p + 1; q < n && keys[q].group == keys[p].group; q++) {
      int j = keys[q].idx;
      if (sup[j]) continue;
      double ovr = pair_any(boxes, fmt, i, j, plus_one, quads);
      if (fabs(ovr - t) <= margin) near++;
      if (cmp_ge ? (ovr >= t) : (ovr > t)) sup[j] = 1;
    }
  }
  int k = 0;
  for (int i = 0; i < n; i++) if (!sup[i]) keep_out[k++] = i;
  if (near_out) *near_out = near;
  free(sup); free(keys); free(quads);
  return k;
}

/* Verify a candidate keep set against the greedy definition with a tolerance
 * band: a kept box must have no kept predecessor with ovr > thr + margin
 * (>= for cmp_ge), a dropped box must have one with ovr > thr - margin.
 * Returns the number of violations; near_out counts pairs inside the band
 * (those are the pairs BASELINE.json excludes and asks to be reported). */
int oracle_nms_verify(const float* boxes, int fmt, const float* scores, const int* groups, int n,
                      const double* thr, int n_thr, int cmp_ge, int plus_one, double margin,
                      const int64_t* keep, int n_keep, int64_t* near_out) {
  if (n <= 0) { if (near_out) *near_out = 0; return n_keep != 0; }
  skey* keys = (skey*)malloc(sizeof(skey) * n);
  for (int i = 0; i < n; i++) { keys[i].score = scores[i]; keys[i].group = groups ? groups[i] : 0; keys[i].idx = i; }
  qsort(keys, n, sizeof(skey), skey_cmp);
  pt* quads = NULL;
  if (fmt != 4) {
    quads = (pt*)malloc(sizeof(pt) * 4 * (size_t)n);
    for (int i = 0; i < n; i++) load_quad(boxes + (size_t)fmt * i, fmt, quads + 4 * (size_t)i);
  }
  uint8_t* kept = (uint8_t*)calloc(n, 1);
  int bad = 0;
  for (int k = 0; k < n_keep; k++) {
    if (keep[k] < 0 || keep[k] >= n || kept[keep[k]] || (k && keep[k] <= keep[k - 1])) bad++;
    else kept[keep[k]] = 1;
  }
  int64_t near = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+:bad, near)
  for (int q = 0; q < n; q++) {
    int j = keys[q].idx;
    double t = thr[n_thr == 1 ? 0 : keys[q].group];
    int hard = 0, soft = 0;
    for (int p = q - 1; p >= 0 && keys[p].group == keys[q].group; p--) {
      int i = keys[p].idx;
      if (!kept[i]) continue;
      double ovr = pair_any(boxes, fmt, i, j, plus_one, quads);
      if (fabs(ovr - t) <= margin) near++;
      if (ovr > t + margin) hard = 1;
      if (ovr > t - margin) soft = 1;
      (void)cmp_ge;
    }
    if (kept[j] && hard) bad++;
    if (!kept[j] && !soft) bad++;
  }
  if (near_out) *near_out = near;
  free(kept); free(keys); free(quads);
  return bad;
}
