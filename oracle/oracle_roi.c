/*
 * oracle_roi.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, float64 arithmetic).
 *
 * RoIAlign forward / backward restated from the reference kernels, plus the
 * rotated generalisation the build adds.  Never linked into the product.
 *
 * Reference lines followed:
 *   bilinear taps + border rules  mmdet/ops/roi_align/src/roi_align_kernel.cu:17-62,143-185
 *   v1 (legacy +1) fwd / bwd      mmdet/ops/roi_align/src/roi_align_kernel.cu:64-117,187-254
 *   v2 (aligned) fwd / bwd        mmdet/ops/roi_align/src/roi_align_kernel_v2.cu:62-128,179-263
 * Rotated rule (new; reduces to v1/v2 at theta = 0): a sample at
 * (xx, yy) in the RoI frame, measured from the RoI centre, is read at
 *   X = ctr_x + xx cos(theta) - yy sin(theta),  Y = ctr_y + xx sin(theta) + yy cos(theta)
 * so the sampled quadrilateral equals thetaobb2pointobb of the RoI
 * (mmdet/core/rbbox/transforms.py:45-55).
 *
 * Pinned against torchvision.ops.roi_align on CPU (the implementation the
 * reference itself offers, mmdet/ops/roi_align/roi_align.py:138-141) in
 * tests/test_oracle.py and tests/golden/.
 *
 * Layout: features NHWC float32 per level, output (K, ph, pw, C) float64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* variant: 0 = v1 legacy (+1, roi_align_kernel.cu), 1 = v2 aligned=false,
 *          2 = v2 aligned=true (roi_align_kernel_v2.cu) */
typedef struct {
  int batch;
  double ctr_x, ctr_y, roi_w, roi_h, cs, sn;
  double start_w, start_h;   /* axis-aligned only: exact reference arithmetic */
  int axis;
} roi_geom;

static roi_geom decode_roi(const float* r, int roi_fmt, double scale, int variant) {
  roi_geom g; memset(&g, 0, sizeof g);
  g.batch = (int)r[0];
  if (roi_fmt == 5) {          /* [b, x1, y1, x2, y2] */
    g.axis = 1; g.cs = 1; g.sn = 0;
    double sw, sh, ew, eh;
    if (variant == 0) {        /* roi_align_kernel.cu:79-86 */
      sw = r[1] * scale; sh = r[2] * scale;
      ew = ((double)r[3] + 1) * scale; eh = ((double)r[4] + 1) * scale;
      g.roi_w = fmax(ew - sw, 0.0); g.roi_h = fmax(eh - sh, 0.0);
    } else {                   /* roi_align_kernel_v2.cu:79-90 */
      double off = (variant == 2) ? 0.5 : 0.0;
      sw = r[1] * scale - off; sh = r[2] * scale - off;
      ew = r[3] * scale - off; eh = r[4] * scale - off;
      g.roi_w = ew - sw; g.roi_h = eh - sh;
      if (variant == 1) { g.roi_w = fmax(g.roi_w, 1.0); g.roi_h = fmax(g.roi_h, 1.0); }
    }
    g.start_w = sw; g.start_h = sh;
  } else {                     /* [b, cx, cy, w, h, theta] */
    g.axis = 0; g.cs = cos((double)r[5]); g.sn = sin((double)r[5]);
    if (variant == 0) {
      g.ctr_x = ((double)r[1] + 0.5) * scale; g.ctr_y = ((double)r[2] + 0.5) * scale;
      g.roi_w = fmax(((double)r[3] + 1) * scale, 0.0); g.roi_h = fmax(((double)r[4] + 1) * scale, 0.0);
    } else {
      double off = (variant == 2) ? 0.5 : 0.0;
      g.ctr_x = r[1] * scale - off; g.ctr_y = r[2] * scale - off;
      g.roi_w = r[3] * scale; g.roi_h = r[4] * scale;
      if (variant == 1) { g.roi_w = fmax(g.roi_w, 1.0); g.roi_h = fmax(g.roi_h, 1.0); }
    }
  }
  return g;
}

/* Error-bound mode for the parity tests (NOT part of the restated algorithm): every accepted tap gets weight 1, so a
 * forward over |features| (backward over |grad_out|) yields  sum_taps |x_t| / count  per element -- the quantity a
 * float32 implementation's weight errors (a few ulp of the sample coordinates each) are multiplied by. */
static int g_unit_weights = 0;
void oracle_roi_set_unit_weights(int on) { g_unit_weights = on; }

/* roi_align_kernel.cu:143-185: tap indices + weights; x_low = -1 marks a
 * rejected sample. */
static void taps(int H, int W, double y, double x, double* w, int* yl, int* yh, int* xl, int* xh) {
  if (y < -1.0 || y > H || x < -1.0 || x > W) {
    w[0] = w[1] = w[2] = w[3] = 0; *xl = *xh = *yl = *yh = -1; return;
  }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  *yl = (int)y; *xl = (int)x;
  if (*yl >= H - 1) { *yh = *yl = H - 1; y = (double)*yl; } else *yh = *yl + 1;
  if (*xl >= W - 1) { *xh = *xl = W - 1; x = (double)*xl; } else *xh = *xl + 1;
  double ly = y - *yl, lx = x - *xl, hy = 1. - ly, hx = 1. - lx;
  w[0] = hy * hx; w[1] = hy * lx; w[2] = ly * hx; w[3] = ly * lx;
  if (g_unit_weights) w[0] = w[1] = w[2] = w[3] = 1.0;
}

static void sample_xy(const roi_geom* g, double bin_w, double bin_h, int ph_, int pw_, int iy, int ix,
                      int gh, int gw, double* X, double* Y) {
  if (g->axis) {
    *Y = g->start_h + ph_ * bin_h + (iy + .5) * bin_h / gh;
    *X = g->start_w + pw_ * bin_w + (ix + .5) * bin_w / gw;
  } else {
    double yy = -0.5 * g->roi_h + ph_ * bin_h + (iy + .5) * bin_h / gh;
    double xx = -0.5 * g->roi_w + pw_ * bin_w + (ix + .5) * bin_w / gw;
    *X = g->ctr_x + xx * g->cs - yy * g->sn;
    *Y = g->ctr_y + xx * g->sn + yy * g->cs;
  }
}

/* One level.  feat: (N,H,W,C) f32.  rois: (K, roi_fmt) f32.  out: (K,ph,pw,C) f64.
 * touched (optional, N*H*W bytes): set to 1 for every pixel read by a tap with
 * non-rejected sample -- the distinct-pixel set U_l of SURVEY 8d. */
void oracle_roi_align_fwd(const float* feat, int N, int H, int W, int C, const float* rois, int roi_fmt, int K,
                          double scale, int ph, int pw, int sample_num, int variant, double* out,
                          uint8_t* touched) {
  (void)N;
#pragma omp parallel for schedule(dynamic, 4)
  for (int k = 0; k < K; k++) {
    roi_geom g = decode_roi(rois + (size_t)roi_fmt * k, roi_fmt, scale, variant);
    double bin_h = g.roi_h / ph, bin_w = g.roi_w / pw;
    int gh = sample_num > 0 ? sample_num : (int)ceil(g.roi_h / ph);
    int gw = sample_num > 0 ? sample_num : (int)ceil(g.roi_w / pw);
    double count = (variant == 0) ? (double)(gh * gw) : (double)((gh * gw > 1) ? gh * gw : 1);
    const float* fb = feat + (size_t)g.batch * H * W * C;
    for (int p = 0; p < ph; p++)
      for (int q = 0; q < pw; q++) {
        double* o = out + (((size_t)k * ph + p) * pw + q) * C;
        for (int c = 0; c < C; c++) o[c] = 0;
        for (int iy = 0; iy < gh; iy++)
          for (int ix = 0; ix < gw; ix++) {
            double X, Y, w[4]; int yl, yh, xl, xh;
            sample_xy(&g, bin_w, bin_h, p, q, iy, ix, gh, gw, &X, &Y);
            taps(H, W, Y, X, w, &yl, &yh, &xl, &xh);
            if (xl < 0) continue;
            const float* t0 = fb + ((size_t)yl * W + xl) * C; const float* t1 = fb + ((size_t)yl * W + xh) * C;
            const float* t2 = fb + ((size_t)yh * W + xl) * C; const float* t3 = fb + ((size_t)yh * W + xh) * C;
            for (int c = 0; c < C; c++) o[c] += w[0] * t0[c] + w[1] * t1[c] + w[2] * t2[c] + w[3] * t3[c];
            if (touched) {
              uint8_t* tb = touched + (size_t)g.batch * H * W;
              tb[(size_t)yl * W + xl] = 1; tb[(size_t)yl * W + xh] = 1;
              tb[(size_t)yh * W + xl] = 1; tb[(size_t)yh * W + xh] = 1;
            }
          }
        for (int c = 0; c < C; c++) o[c] /= count;
      }
  }
}

/* grad_out: (K,ph,pw,C) f32 -> grad_feat (N,H,W,C) f64, accumulated (caller zeroes). */
void oracle_roi_align_bwd(const float* grad_out, int N, int H, int W, int C, const float* rois, int roi_fmt, int K,
                          double scale, int ph, int pw, int sample_num, int variant, double* grad_feat) {
  (void)N;
  for (int k = 0; k < K; k++) {
    roi_geom g = decode_roi(rois + (size_t)roi_fmt * k, roi_fmt, scale, variant);
    double bin_h = g.roi_h / ph, bin_w = g.roi_w / pw;
    int gh = sample_num > 0 ? sample_num : (int)ceil(g.roi_h / ph);
    int gw = sample_num > 0 ? sample_num : (int)ceil(g.roi_w / pw);
    double count = (variant == 0) ? (double)(gh * gw) : (double)((gh * gw > 1) ? gh * gw : 1);
    double* fb = grad_feat + (size_t)g.batch * H * W * C;
    for (int p = 0; p < ph; p++)
      for (int q = 0; q < pw; q++) {
        const float* go = grad_out + (((size_t)k * ph + p) * pw + q) * C;
        for (int iy = 0; iy < gh; iy++)
          for (int ix = 0; ix < gw; ix++) {
            double X, Y, w[4]; int yl, yh, xl, xh;
            sample_xy(&g, bin_w, bin_h, p, q, iy, ix, gh, gw, &X, &Y);
            taps(H, W, Y, X, w, &yl, &yh, &xl, &xh);
            if (xl < 0) continue;
            double* t0 = fb + ((size_t)yl * W + xl) * C; double* t1 = fb + ((size_t)yl * W + xh) * C;
            double* t2 = fb + ((size_t)yh * W + xl) * C; double* t3 = fb + ((size_t)yh * W + xh) * C;
            for (int c = 0; c < C; c++) {
              double gv = go[c];
              t0[c] += gv * w[0] / count; t1[c] += gv * w[1] / count;
              t2[c] += gv * w[2] / count; t3[c] += gv * w[3] / count;
            }
          }
      }
  }
}

/* mmdet/models/roi_extractors/single_level.py:54-73 map_roi_levels:
 * scale = sqrt((x2-x1+1)(y2-y1+1)); lvl = floor(log2(scale/finest + 1e-6)) clamped. */
void oracle_map_roi_levels(const float* rois5, int K, double finest_scale, int num_levels, int* lvl_out) {
  for (int k = 0; k < K; k++) {
    const float* r = rois5 + 5 * (size_t)k;
    double s = sqrt(((double)r[3] - r[1] + 1) * ((double)r[4] - r[2] + 1));
    double l = floor(log2(s / finest_scale + 1e-6));
    if (l < 0) l = 0;
    if (l > num_levels - 1) l = num_levels - 1;
    lvl_out[k] = (int)l;
  }
}
