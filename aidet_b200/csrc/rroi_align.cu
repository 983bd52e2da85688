// rroi_align.cu -- rotated / axis-aligned RoIAlign forward + backward, multi-level, NHWC, sm_100a.
//
// Replaces (reference):
//   mmdet/ops/roi_align/src/roi_align_kernel.cu:64-141,187-283      v1 (legacy +1) fwd / bwd
//   mmdet/ops/roi_align/src/roi_align_kernel_v2.cu:62-128,179-348   v2 (aligned) fwd / bwd
//   mmdet/models/roi_extractors/single_level.py:89-107              per-level gather/launch/scatter
// The reference maps one thread to one NCHW output element: 16 scattered 4-byte loads per
// element and every thread recomputes the sample geometry.  Here:
//   - features are channels-last (N,H,W,C): a bilinear tap is a contiguous C*4-byte run,
//     read with 128-bit loads by consecutive lanes (lanes = channel quads)
//   - one CTA per RoI; the sample geometry (tap offsets + weights per sample point, border
//     rules of roi_align_kernel.cu:17-62) is computed ONCE per RoI into shared memory and
//     broadcast to the channel lanes; the default forward (sample_num 1/2) merges the taps
//     of a bin by pixel first (tap list, see rroi_align_fwd_taplist_kernel)
//   - all FPN levels run in one launch (per-RoI level id), rotated and axis-aligned RoIs
//     share the kernel (theta = 0 reproduces v1/v2)
//   - backward: gather form (taps bucketed per pixel, every gradient pixel written once, no
//     atomics on the maps) for fixed sampling grids; scatter form with 128-bit vector
//     reductions (red.global.add.v4.f32) for adaptive grids
// Bound: HBM bandwidth (~2 flop per byte read); the forward additionally by load latency.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include "common.cuh"

namespace aidet {

constexpr int kMaxLevels = 8;
constexpr int kTable = 256;        // sample points staged per pass

struct RoiLevels {
  const float* feat[kMaxLevels];
  float* grad[kMaxLevels];
  int H[kMaxLevels];
  int W[kMaxLevels];
  float scale[kMaxLevels];
};

struct RoiGeom {
  float ox, oy;        // origin: RoI start (axis-aligned) or centre (rotated), feature px
  float xb, yb;        // first-sample base offset in the RoI frame (0 or -roi/2)
  float bin_w, bin_h;
  float cs, sn;
  int gh, gw;
  int H, W;
  int batch, level;
  float inv_count;
};

struct __align__(16) SampleTap { int off[4]; float w[4]; };

// variant: 0 = v1 legacy (+1), 1 = v2 aligned=false, 2 = v2 aligned=true
__device__ __forceinline__ void decode_roi(const float* __restrict__ r, int roi_fmt, float scale, int variant, int ph,
                                           int pw, int sample_num, RoiGeom& g) {
  float roi_w, roi_h;
  if (roi_fmt == 5) {                       // [b, x1, y1, x2, y2]
    float sw, sh, ew, eh;
    if (variant == 0) {                     // roi_align_kernel.cu:79-86
      sw = r[1] * scale; sh = r[2] * scale;
      ew = (r[3] + 1.f) * scale; eh = (r[4] + 1.f) * scale;
      roi_w = fmaxf(ew - sw, 0.f); roi_h = fmaxf(eh - sh, 0.f);
    } else {                                // roi_align_kernel_v2.cu:79-90
      float off = (variant == 2) ? 0.5f : 0.f;
      sw = r[1] * scale - off; sh = r[2] * scale - off;
      ew = r[3] * scale - off; eh = r[4] * scale - off;
      roi_w = ew - sw; roi_h = eh - sh;
      if (variant == 1) { roi_w = fmaxf(roi_w, 1.f); roi_h = fmaxf(roi_h, 1.f); }
    }
    g.ox = sw; g.oy = sh; g.xb = 0.f; g.yb = 0.f; g.cs = 1.f; g.sn = 0.f;
  } else {                                  // [b, cx, cy, w, h, theta]
    if (variant == 0) {
      g.ox = (r[1] + 0.5f) * scale; g.oy = (r[2] + 0.5f) * scale;
      roi_w = fmaxf((r[3] + 1.f) * scale, 0.f); roi_h = fmaxf((r[4] + 1.f) * scale, 0.f);
    } else {
      float off = (variant == 2) ? 0.5f : 0.f;
      g.ox = r[1] * scale - off; g.oy = r[2] * scale - off;
      roi_w = r[3] * scale; roi_h = r[4] * scale;
      if (variant == 1) { roi_w = fmaxf(roi_w, 1.f); roi_h = fmaxf(roi_h, 1.f); }
    }
    float sn, cs;                           // full-range sincosf (<= 2 ulp), not the fast intrinsic
    sincosf(r[5], &sn, &cs);
    g.cs = cs; g.sn = sn; g.xb = -0.5f * roi_w; g.yb = -0.5f * roi_h;
  }
  g.bin_w = roi_w / (float)pw; g.bin_h = roi_h / (float)ph;
  g.gh = sample_num > 0 ? sample_num : (int)ceilf(roi_h / (float)ph);
  g.gw = sample_num > 0 ? sample_num : (int)ceilf(roi_w / (float)pw);
  if (g.gh < 0) g.gh = 0;
  if (g.gw < 0) g.gw = 0;
  int cnt = g.gh * g.gw;
  // v1 divides by gh*gw (0/0 -> NaN for an empty grid, roi_align_kernel.cu:114); v2 by max(.,1)
  g.inv_count = (variant == 0) ? (1.0f / (float)cnt) : (1.0f / (float)max(cnt, 1));
}

// bilinear_interpolate(_gradient) border rules, roi_align_kernel.cu:17-62,143-185
struct SampleCell { int yl, xl, yh, xh; float ly, lx; bool ok; };

__device__ __forceinline__ SampleCell sample_cell(const RoiGeom& g, float x, float y) {
  SampleCell c;
  c.ok = !(y < -1.0f || y > (float)g.H || x < -1.0f || x > (float)g.W);
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= g.H - 1) { yh = yl = g.H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= g.W - 1) { xh = xl = g.W - 1; x = (float)xl; } else xh = xl + 1;
  c.yl = yl; c.xl = xl; c.yh = yh; c.xh = xh; c.ly = y - yl; c.lx = x - xl;
  return c;
}

__device__ __forceinline__ void sample_point(const RoiGeom& g, int q, int pw, float& x, float& y) {
  const int S = g.gh * g.gw;
  const int bin = q / S, rem = q - bin * S;
  const int iy = rem / g.gw, ix = rem - iy * g.gw;
  const int p_h = bin / pw, p_w = bin - p_h * pw;
  float yy = g.yb + p_h * g.bin_h + (iy + .5f) * g.bin_h / (float)g.gh;
  float xx = g.xb + p_w * g.bin_w + (ix + .5f) * g.bin_w / (float)g.gw;
  x = g.ox + xx * g.cs - yy * g.sn;
  y = g.oy + xx * g.sn + yy * g.cs;
}

__device__ __forceinline__ SampleTap cell_taps(const RoiGeom& g, const SampleCell& c) {
  SampleTap t;
  if (!c.ok) {
    t.off[0] = -1; t.off[1] = t.off[2] = t.off[3] = 0;
    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
    return t;
  }
  const float hy = 1.f - c.ly, hx = 1.f - c.lx;
  t.off[0] = c.yl * g.W + c.xl; t.off[1] = c.yl * g.W + c.xh; t.off[2] = c.yh * g.W + c.xl; t.off[3] = c.yh * g.W + c.xh;
  t.w[0] = hy * hx; t.w[1] = hy * c.lx; t.w[2] = c.ly * hx; t.w[3] = c.ly * c.lx;
  return t;
}

__device__ __forceinline__ SampleTap make_sample(const RoiGeom& g, int q, int pw) {
  float x, y;
  sample_point(g, q, pw, x, y);
  return cell_taps(g, sample_cell(g, x, y));
}

template <int VEC> struct VecT;
template <> struct VecT<4> { using T = float4; };
template <> struct VecT<1> { using T = float; };

__device__ __forceinline__ float4 vfma(float w, float4 v, float4 a) {
  a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
  return a;
}
__device__ __forceinline__ float vfma(float w, float v, float a) { return fmaf(w, v, a); }
__device__ __forceinline__ float4 vscale(float4 v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }
__device__ __forceinline__ float vscale(float v, float s) { return v * s; }
__device__ __forceinline__ void vzero(float4& v) { v = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(float& v) { v = 0.f; }
__device__ __forceinline__ float4 vldg(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float vldg(const float* p) { return __ldg(p); }
__device__ __forceinline__ void vred(float4* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void vred(float* p, float v) { atomicAdd(p, v); }

// One CTA per RoI.  Threads = (channel lane cl < CL) x (bin slot); bins are dealt round-robin
// to the slots, the sample table is filled cooperatively kTable points at a time.
template <int VEC, bool BWD>
__global__ void __launch_bounds__(256)
rroi_align_kernel(RoiLevels lv, int n_levels, int N, int C, const float* __restrict__ rois, int roi_fmt,
                  const int* __restrict__ roi_level, int ph, int pw, int sample_num, int variant,
                  float* __restrict__ io, int CL) {
  using V = typename VecT<VEC>::T;
  __shared__ RoiGeom g;
  __shared__ SampleTap table[kTable];
  const int k = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) {
    const float* r = rois + (size_t)k * roi_fmt;
    int lvl = roi_level ? roi_level[k] : 0;
    lvl = min(max(lvl, 0), n_levels - 1);
    g.level = lvl; g.H = lv.H[lvl]; g.W = lv.W[lvl];
    g.batch = (int)r[0];
    decode_roi(r, roi_fmt, lv.scale[lvl], variant, ph, pw, sample_num, g);
  }
  __syncthreads();
  const int nbins = ph * pw;
  const int S = g.gh * g.gw;
  const int nch = C / VEC;
  const int nslots = 256 / CL, slot = tid / CL, cl = tid % CL;
  const bool batch_ok = g.batch >= 0 && g.batch < N;
  V* iok = reinterpret_cast<V*>(io) + (size_t)k * nbins * nch;       // out (fwd) or grad_out (bwd)
  if (S <= 0 || !batch_ok) {
    if (!BWD) {     // empty sampling grid: v1 computes 0/0 (roi_align_kernel.cu:114), v2 writes 0
      const float fill = (variant == 0 && batch_ok) ? __int_as_float(0x7fc00000) : 0.f;
      float* o = reinterpret_cast<float*>(iok);
      for (int i = tid; i < nbins * nch * VEC; i += 256) o[i] = fill;
    }
    return;
  }
  const long long Q = (long long)nbins * S;
  const size_t plane = (size_t)g.batch * g.H * g.W;
  const V* feat = reinterpret_cast<const V*>(lv.feat[g.level]) + plane * nch;
  V* grad = BWD ? reinterpret_cast<V*>(lv.grad[g.level]) + plane * nch : nullptr;
  const float inv_count = g.inv_count;

  for (int cc = cl; cc < ((nch + CL - 1) / CL) * CL; cc += CL) {
    const bool ch_ok = cc < nch;
    V acc; vzero(acc);
    V gv; vzero(gv);
    int gv_bin = -1;
    for (long long q0 = 0; q0 < Q; q0 += kTable) {
      const int nq = (int)min((long long)kTable, Q - q0);
      __syncthreads();
      for (int e = tid; e < nq; e += 256) table[e] = make_sample(g, (int)(q0 + e), pw);
      __syncthreads();
      const int b_lo = (int)(q0 / S), b_hi = (int)((q0 + nq - 1) / S);
      int bin = b_lo + ((slot - b_lo % nslots) + nslots) % nslots;      // first bin >= b_lo owned by this slot
      for (; bin <= b_hi; bin += nslots) {
        const int e_lo = (int)max((long long)bin * S - q0, 0LL);
        const int e_hi = (int)min((long long)(bin + 1) * S - q0, (long long)nq);
        if (ch_ok) {
          if (BWD) {
            if (gv_bin != bin) { gv = vscale(iok[(size_t)bin * nch + cc], inv_count); gv_bin = bin; }
#pragma unroll 2
            for (int e = e_lo; e < e_hi; ++e) {
              const SampleTap t = table[e];
              if (t.off[0] < 0) continue;
              vred(grad + (size_t)t.off[0] * nch + cc, vscale(gv, t.w[0]));
              vred(grad + (size_t)t.off[1] * nch + cc, vscale(gv, t.w[1]));
              vred(grad + (size_t)t.off[2] * nch + cc, vscale(gv, t.w[2]));
              vred(grad + (size_t)t.off[3] * nch + cc, vscale(gv, t.w[3]));
            }
          } else {
#pragma unroll 4
            for (int e = e_lo; e < e_hi; ++e) {
              const SampleTap t = table[e];
              if (t.off[0] < 0) continue;
              V v0 = vldg(feat + (size_t)t.off[0] * nch + cc);
              V v1 = vldg(feat + (size_t)t.off[1] * nch + cc);
              V v2 = vldg(feat + (size_t)t.off[2] * nch + cc);
              V v3 = vldg(feat + (size_t)t.off[3] * nch + cc);
              acc = vfma(t.w[0], v0, acc); acc = vfma(t.w[1], v1, acc);
              acc = vfma(t.w[2], v2, acc); acc = vfma(t.w[3], v3, acc);
            }
            if ((long long)(bin + 1) * S <= q0 + nq) {                 // bin complete
              iok[(size_t)bin * nch + cc] = vscale(acc, inv_count);
              vzero(acc);
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Fast path (VEC = 4, S = gh*gw <= kTable): the shape every AIDet config uses (7x7 or 14x14
// bins, sample_num 2).  Same mapping as above -- one CTA per RoI, lanes = channel quads,
// slots = bins -- but the table holds float4-index offsets (tap pixel * C/4) and weights that
// already include 1/count, a whole number of bins per table pass, and 32-bit arithmetic only:
// per sample a thread issues 2 LDS.128, 4 IMAD.WIDE, 4 LDG.128 (or 4 RED.128) and 16 FFMA.
// Rejected samples keep offset -1: their loads are predicated off (no divergence), so a
// non-finite feature value can never leak through a zero weight.
template <bool BWD>
__global__ void __launch_bounds__(256)
rroi_align_fast_kernel(RoiLevels lv, int n_levels, int N, int C, const float* __restrict__ rois, int roi_fmt,
                       const int* __restrict__ roi_level, int ph, int pw, int sample_num, int variant,
                       float* __restrict__ io, int CL, int nslots, int groups_per_roi, int bins_per_group) {
  __shared__ RoiGeom g;
  // BWD: tap offsets in bytes from the image plane (x = -1: rejected) + tap weights, 1/count folded in.
  // FWD: ONE 16-byte record per sample (the forward is bound by the L1 data pipe, ncu r1b: 80-85 % busy,
  // a fifth of it these table reads): x = byte offset of tap 0 | dx | dy << 1 (0xffffffff: rejected) where
  // dx, dy in {0, 1} say whether the right / lower neighbour is a different pixel (border clamp), y = hy / count,
  // z = ly / count, w = lx; the four weights are rebuilt in registers (5 instructions).
  __shared__ __align__(16) int4 toff[kTable];
  __shared__ __align__(16) float4 tw[BWD ? kTable : 1];
  __shared__ int n_bad;                            // rejected samples in the current table pass
  const int k = blockIdx.x / groups_per_roi;       // the RoI's bins are split over groups_per_roi small CTAs
  const int grp = blockIdx.x - k * groups_per_roi;
  const int tid = threadIdx.x;
  if (tid == 0) {
    const float* r = rois + (size_t)k * roi_fmt;
    int lvl = roi_level ? roi_level[k] : 0;
    lvl = min(max(lvl, 0), n_levels - 1);
    g.level = lvl; g.H = lv.H[lvl]; g.W = lv.W[lvl];
    g.batch = (int)r[0];
    decode_roi(r, roi_fmt, lv.scale[lvl], variant, ph, pw, sample_num, g);
    n_bad = 0;
  }
  __syncthreads();
  const int nbins = ph * pw;
  const int bin_begin = grp * bins_per_group, bin_end = min(nbins, bin_begin + bins_per_group);
  const int S = g.gh * g.gw;
  const int nch = C >> 2;
  const int slot = tid / CL, cl = tid - slot * CL;
  const bool batch_ok = g.batch >= 0 && g.batch < N;
  float4* iok = reinterpret_cast<float4*>(io) + (size_t)k * nbins * nch;       // out (fwd) or grad_out (bwd)
  if (S <= 0 || !batch_ok) {
    if (!BWD) {     // empty sampling grid: v1 computes 0/0 (roi_align_kernel.cu:114), v2 writes 0
      const float fill = (variant == 0 && batch_ok) ? __int_as_float(0x7fc00000) : 0.f;
      float* o = reinterpret_cast<float*>(iok);
      for (int i = bin_begin * C + tid; i < bin_end * C; i += blockDim.x) o[i] = fill;
    }
    return;
  }
  // image plane of this RoI, as a byte pointer; the empty asm keeps it in registers (otherwise the
  // compiler re-derives it from the parameter table inside the tap loop)
  const size_t plane = (size_t)g.batch * g.H * g.W * C;
  const char* feat = BWD ? nullptr : reinterpret_cast<const char*>(lv.feat[g.level] + plane);
  char* grad = BWD ? reinterpret_cast<char*>(lv.grad[g.level] + plane) : nullptr;
  asm volatile("" : "+l"(feat), "+l"(grad));
  const float inv_count = g.inv_count;
  const int bins_per_pass = kTable / S;                       // >= 1 (host guarantees S <= kTable)
  const unsigned pix_bytes = (unsigned)C * 4u, row_bytes = pix_bytes * (unsigned)g.W;

  for (int bin0 = bin_begin; bin0 < bin_end; bin0 += bins_per_pass) {
    const int nb = min(bins_per_pass, bin_end - bin0);
    const int nq = nb * S;
    if (bin0 != bin_begin) { __syncthreads(); if (tid == 0) n_bad = 0; __syncthreads(); }
    for (int e = tid; e < nq; e += blockDim.x) {
      float sx, sy;
      sample_point(g, bin0 * S + e, pw, sx, sy);
      const SampleCell c = sample_cell(g, sx, sy);
      const bool ok = c.ok;
      if (!ok) atomicAdd(&n_bad, 1);
      if (BWD) {
        const SampleTap t = cell_taps(g, c);
        toff[e] = ok ? make_int4(t.off[0] * pix_bytes, t.off[1] * pix_bytes, t.off[2] * pix_bytes, t.off[3] * pix_bytes)
                     : make_int4(-1, 0, 0, 0);
        tw[e] = make_float4(t.w[0] * inv_count, t.w[1] * inv_count, t.w[2] * inv_count, t.w[3] * inv_count);
      } else {
        const unsigned dx = c.xh != c.xl, dy = c.yh != c.yl;
        toff[e] = ok ? make_int4((int)((unsigned)(c.yl * g.W + c.xl) * pix_bytes | dx | (dy << 1)),
                                 __float_as_int((1.f - c.ly) * inv_count), __float_as_int(c.ly * inv_count),
                                 __float_as_int(c.lx))
                     : make_int4(-1, 0, 0, 0);
      }
    }
    __syncthreads();
    if (slot >= nslots) continue;
    const bool all_ok = n_bad == 0;                                       // CTA-uniform
    for (int cc = cl; cc < nch; cc += CL) {
      const unsigned cbyte = (unsigned)cc * 16u;
      for (int b = slot; b < nb; b += nslots) {
        const int bin = bin0 + b;
        const int e0 = b * S;
        if (BWD) {
          const float4 gv = __ldcs(iok + (size_t)bin * nch + cc);
          char* gc = grad + cbyte;
#pragma unroll 2
          for (int e = e0; e < e0 + S; ++e) {
            const int4 o = toff[e];
            if (o.x < 0) continue;                                      // warp-uniform
            const float4 w = tw[e];
            vred(reinterpret_cast<float4*>(gc + (unsigned)o.x), vscale(gv, w.x));
            vred(reinterpret_cast<float4*>(gc + (unsigned)o.y), vscale(gv, w.y));
            vred(reinterpret_cast<float4*>(gc + (unsigned)o.z), vscale(gv, w.z));
            vred(reinterpret_cast<float4*>(gc + (unsigned)o.w), vscale(gv, w.w));
          }
        } else {
          const char* fc = feat + cbyte;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          // one sample: 4 loads issued first (addresses from the packed record), then weights + 16 FFMA
#define AIDET_TAP_ADDR(R, O1, O2, O3)                                                   \
          const unsigned R##0 = (unsigned)R.x & ~3u;                                      \
          const unsigned O1 = R##0 + (((unsigned)R.x & 1u) ? pix_bytes : 0u);             \
          const unsigned O2 = R##0 + (((unsigned)R.x & 2u) ? row_bytes : 0u);             \
          const unsigned O3 = O2 + (O1 - R##0);
#define AIDET_TAP_FMA(R, V0, V1, V2, V3)                                                \
          {                                                                             \
            const float hyi = __int_as_float(R.y), lyi = __int_as_float(R.z), lx = __int_as_float(R.w), hx = 1.f - lx; \
            acc = vfma(hyi * hx, V0, acc); acc = vfma(hyi * lx, V1, acc);                   \
            acc = vfma(lyi * hx, V2, acc); acc = vfma(lyi * lx, V3, acc);                   \
          }
          if (all_ok) {                   // no rejected sample in this pass: straight-line loads, 2 samples in flight
            int e = e0;
#pragma unroll 1
            for (; e + 2 <= e0 + S; e += 2) {
              const int4 ra = toff[e], rb = toff[e + 1];
              AIDET_TAP_ADDR(ra, oa1, oa2, oa3)
              AIDET_TAP_ADDR(rb, ob1, ob2, ob3)
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(fc + ra0));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(fc + oa1));
              const float4 a2 = __ldg(reinterpret_cast<const float4*>(fc + oa2));
              const float4 a3 = __ldg(reinterpret_cast<const float4*>(fc + oa3));
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(fc + rb0));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(fc + ob1));
              const float4 b2 = __ldg(reinterpret_cast<const float4*>(fc + ob2));
              const float4 b3 = __ldg(reinterpret_cast<const float4*>(fc + ob3));
              AIDET_TAP_FMA(ra, a0, a1, a2, a3)
              AIDET_TAP_FMA(rb, b0, b1, b2, b3)
            }
            for (; e < e0 + S; ++e) {
              const int4 ra = toff[e];
              AIDET_TAP_ADDR(ra, oa1, oa2, oa3)
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(fc + ra0));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(fc + oa1));
              const float4 a2 = __ldg(reinterpret_cast<const float4*>(fc + oa2));
              const float4 a3 = __ldg(reinterpret_cast<const float4*>(fc + oa3));
              AIDET_TAP_FMA(ra, a0, a1, a2, a3)
            }
          } else {                        // RoI hangs over the image border: rejected samples are skipped
            for (int e = e0; e < e0 + S; ++e) {
              const int4 ra = toff[e];
              if (ra.x == -1) continue;                                    // warp-uniform; rejected taps are never read
              AIDET_TAP_ADDR(ra, oa1, oa2, oa3)
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(fc + ra0));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(fc + oa1));
              const float4 a2 = __ldg(reinterpret_cast<const float4*>(fc + oa2));
              const float4 a3 = __ldg(reinterpret_cast<const float4*>(fc + oa3));
              AIDET_TAP_FMA(ra, a0, a1, a2, a3)
            }
          }
#undef AIDET_TAP_ADDR
#undef AIDET_TAP_FMA
          __stcs(iok + (size_t)bin * nch + cc, acc);
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------
// Forward with merged taps ("tap list").  ncu on the kernel above: DRAM traffic equals the algorithmic bytes,
// but every tap of every sample is a 512-byte warp load = 4 L1 wavefronts at ~2 cycles each
// (B300_MICROARCH.md: 2.07 cyc/wavefront inside one LDG), i.e. ~64 B/clk/SM: 3.3 GB of tap loads on C3 take
// 0.177 ms whatever DRAM does -- the kernel sits exactly on that L1 bound.  At the FPN level a RoI is mapped
// to, a bin is 2-4 feature pixels and its 2x2 samples are 1-2 px apart, so the 16 taps of a bin touch ~9
// distinct pixels (57 % on C3).  Here one CTA owns a whole RoI (<= kTLBins bins):
//   build  : T = 4*G*G lanes per bin compute one tap each; taps of a bin that hit the same pixel are merged
//            with __match_any_sync (weight = sum in tap order, deterministic) into a per-bin list of
//            (byte offset, weight) entries in shared memory -- two barriers per CTA in all;
//   consume: a warp owns one (bin, 32-channel-quad chunk) task at a time and walks the bin's list kChunk
//            entries per step: two broadcast LDS.128 per four entries, predicated LDG.128s all issued before
//            the FFMAs; consecutive warps take the chunks of the same bin, so both halves of a pixel's
//            1 KB are fetched together.
// sample_num 1 and 2 (every AIDet config: 2).
constexpr int kTLBins = 64;          // bins per CTA

// One tap (corner j of sample smp of `bin`) of a G x G sampling grid: pixel index in the plane (-1: rejected
// sample) and weight -- the lane-per-tap form of make_sample_fixed.
template <int G>
__device__ __forceinline__ int sample_tap_fixed(const RoiGeom& g, int bin, int smp, int j, int pw, float& w) {
  const int iy = smp / G, ix = smp - iy * G;
  const int p_h = bin / pw, p_w = bin - p_h * pw;
  const float yy = g.yb + p_h * g.bin_h + (iy + .5f) * g.bin_h / (float)G;
  const float xx = g.xb + p_w * g.bin_w + (ix + .5f) * g.bin_w / (float)G;
  const float x = g.ox + xx * g.cs - yy * g.sn;
  const float y = g.oy + xx * g.sn + yy * g.cs;
  const SampleCell c = sample_cell(g, x, y);
  w = ((j & 2) ? c.ly : 1.f - c.ly) * ((j & 1) ? c.lx : 1.f - c.lx);
  return c.ok ? ((j & 2) ? c.yh : c.yl) * g.W + ((j & 1) ? c.xh : c.xl) : -1;
}

// Sum of w over the lanes of mask m (a __match_any_sync group), in ascending lane order, for every lane at once.
__device__ __forceinline__ float group_sum_ordered(unsigned m, float w) {
  const int n = __reduce_max_sync(0xffffffffu, __popc(m));           // warp uniform trip count (typically 1-4)
  float ws = 0.f;
  for (int it = 0; it < n; ++it) {
    const int src = m ? __ffs(m) - 1 : 0;
    const float wj = __shfl_sync(0xffffffffu, w, src);
    if (m) ws += wj;
    m &= m - 1;
  }
  return ws;
}

// NP pairs of tap-list entries (two per 16-byte shared-memory word): all loads first, then the FFMAs, no branches --
// so the loads of a bin are in flight together (8 at a time: a warp keeps 4 KB in flight).
template <int NP>
__device__ __forceinline__ float4 taplist_consume(const uint4* __restrict__ e, const char* __restrict__ fc, float4 acc) {
  constexpr int A = NP > 4 ? 4 : NP;
  uint4 p[A];
  float4 v[2 * A];
#pragma unroll
  for (int j = 0; j < A; ++j) p[j] = e[j];
#pragma unroll
  for (int j = 0; j < A; ++j) {
    v[2 * j] = __ldg(reinterpret_cast<const float4*>(fc + p[j].x));
    v[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(fc + p[j].z));
  }
#pragma unroll
  for (int j = 0; j < A; ++j) {
    acc = vfma(__uint_as_float(p[j].y), v[2 * j], acc);
    acc = vfma(__uint_as_float(p[j].w), v[2 * j + 1], acc);
  }
  if constexpr (NP > 4) acc = taplist_consume<NP - 4>(e + 4, fc, acc);
  return acc;
}

// The same for TWO 32-quad channel chunks per warp (lane loads quad cc and quad cc + 32 of every pixel, i.e. the warp
// reads a 256-channel pixel's whole 1 KB back to back): 2 pairs = 8 loads in flight.
template <int NP>
__device__ __forceinline__ void taplist_consume2(const uint4* __restrict__ e, const char* __restrict__ fc, float4& acc0,
                                                 float4& acc1) {
  constexpr int A = NP > 2 ? 2 : NP;
  uint4 p[A];
  float4 v[4 * A];
#pragma unroll
  for (int j = 0; j < A; ++j) p[j] = e[j];
#pragma unroll
  for (int j = 0; j < A; ++j) {
    v[4 * j] = __ldg(reinterpret_cast<const float4*>(fc + p[j].x));
    v[4 * j + 1] = __ldg(reinterpret_cast<const float4*>(fc + p[j].x + 512u));
    v[4 * j + 2] = __ldg(reinterpret_cast<const float4*>(fc + p[j].z));
    v[4 * j + 3] = __ldg(reinterpret_cast<const float4*>(fc + p[j].z + 512u));
  }
#pragma unroll
  for (int j = 0; j < A; ++j) {
    acc0 = vfma(__uint_as_float(p[j].y), v[4 * j], acc0);
    acc1 = vfma(__uint_as_float(p[j].y), v[4 * j + 1], acc1);
    acc0 = vfma(__uint_as_float(p[j].w), v[4 * j + 2], acc0);
    acc1 = vfma(__uint_as_float(p[j].w), v[4 * j + 3], acc1);
  }
  if constexpr (NP > 2) taplist_consume2<NP - 2>(e + 2, fc, acc0, acc1);
}

template <int G, int OCC, bool PAIR>
__global__ void __launch_bounds__(256, OCC)
rroi_align_fwd_taplist_kernel(RoiLevels lv, int n_levels, int N, int C, const float* __restrict__ rois, int roi_fmt,
                              const int* __restrict__ roi_level, int ph, int pw, int variant,
                              float* __restrict__ out, int groups_per_roi, int bins_per_group, int chunk_shift) {
  constexpr int S = G * G, T = 4 * S, BPW = 32 / T;
  __shared__ __align__(16) uint2 ent[kTLBins * T];     // per bin: T slots, the first cnt[b] hold merged entries
  __shared__ int cnt[kTLBins];                         // entries per bin, padded to an even number
  const int k = blockIdx.x / groups_per_roi;
  const int grp = blockIdx.x - k * groups_per_roi;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  RoiGeom g;                                           // every thread decodes the RoI (uniform): no broadcast barrier
  {
    const float* r = rois + (size_t)k * roi_fmt;
    int lvl = roi_level ? __ldg(roi_level + k) : 0;
    lvl = min(max(lvl, 0), n_levels - 1);
    g.level = lvl; g.H = lv.H[lvl]; g.W = lv.W[lvl];
    float rr[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) rr[i] = (i < roi_fmt) ? __ldg(r + i) : 0.f;
    g.batch = (int)rr[0];
    decode_roi(rr, roi_fmt, lv.scale[lvl], variant, ph, pw, G, g);
  }
  const int nbins = ph * pw;
  const int bin_begin = grp * bins_per_group, bin_end = min(nbins, bin_begin + bins_per_group);
  const int nb = bin_end - bin_begin;
  const int nch = C >> 2;
  float4* outk = reinterpret_cast<float4*>(out) + ((size_t)k * nbins + bin_begin) * nch;
  if (!(g.batch >= 0 && g.batch < N)) {
    float* o = reinterpret_cast<float*>(outk);
    for (int i = tid; i < nb * C; i += blockDim.x) o[i] = 0.f;
    return;
  }
  const float inv_count = g.inv_count;
  const unsigned pix_bytes = (unsigned)C * 4u;
  const int seg = lane & ~(T - 1), i_tap = lane & (T - 1);
  for (int p = warp; p * BPW < nb; p += nwarps) {                    // ---- build
    const int b = p * BPW + lane / T;                                // bin of this lane, local to the CTA
    const bool valid = b < nb;                                       // (whole T-lane segments agree)
    int pix = -1; float w = 0.f;
    if (valid) pix = sample_tap_fixed<G>(g, bin_begin + b, i_tap >> 2, i_tap & 3, pw, w);
    const bool tap_ok = pix >= 0;
    // equal keys <=> same pixel of the same bin; rejected / idle lanes get unique keys
    const unsigned key = tap_ok ? (((unsigned)pix << 3) | (unsigned)(lane / T)) : (0x80000000u | (unsigned)lane);
    const unsigned m = __match_any_sync(0xffffffffu, key);
    const bool f = tap_ok && (lane == __ffs(m) - 1);
    const float ws = group_sum_ordered(m, w) * inv_count;            // duplicates summed in tap order
    const unsigned fb = (__ballot_sync(0xffffffffu, f) >> seg) & ((T == 32) ? 0xffffffffu : ((1u << T) - 1u));
    const int c = __popc(fb), within = __popc(fb & ((1u << i_tap) - 1u));
    if (valid && i_tap == 0) cnt[b] = (c + 1) & ~1;
    if (f) {
      const unsigned off = (unsigned)pix * pix_bytes;
      ent[b * T + within] = make_uint2(off, __float_as_uint(ws));
      // odd count: one zero-weight copy of the last entry (an L1 hit) so that the consumer needs no predicates
      if (within == c - 1 && (c & 1)) ent[b * T + c] = make_uint2(off, 0u);
    }
  }
  __syncthreads();
  // ---- consume: warp -> fixed channel chunk (or chunk pair), bins strided (nwarps and the chunk count are powers of two)
  if constexpr (PAIR) {
    // C % 256 == 0: a warp owns TWO chunks of a bin, so it reads each pixel's 2 x 512 bytes back to back (C3: 0.145 vs
    // 0.154 ms with one chunk per warp) -- fewer, larger bursts per DRAM page.
    const int cs = chunk_shift - 1;                               // pairs of chunks per bin = 2^cs
    const int cc = (warp & ((1 << cs) - 1)) * 64 + lane;
    const size_t plane2 = (size_t)g.batch * g.H * g.W * C;
    const char* fc2 = reinterpret_cast<const char*>(lv.feat[g.level] + plane2) + (unsigned)cc * 16u;
    float4* dst2 = outk + cc;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = warp >> cs; b < nb; b += nwarps >> cs) {
      const int c = cnt[b];
      const uint4* e = reinterpret_cast<const uint4*>(ent + b * T);
      float4 a0 = z, a1 = z;
      switch (c >> 1) {
        case 1: taplist_consume2<1>(e, fc2, a0, a1); break;
        case 2: taplist_consume2<2>(e, fc2, a0, a1); break;
        case 3: taplist_consume2<3>(e, fc2, a0, a1); break;
        case 4: taplist_consume2<4>(e, fc2, a0, a1); break;
        case 5: taplist_consume2<5>(e, fc2, a0, a1); break;
        case 6: taplist_consume2<6>(e, fc2, a0, a1); break;
        case 7: taplist_consume2<7>(e, fc2, a0, a1); break;
        case 8: taplist_consume2<8>(e, fc2, a0, a1); break;
        default: break;
      }
      __stcs(dst2 + (size_t)b * nch, a0);
      __stcs(dst2 + (size_t)b * nch + 32, a1);
    }
  } else {
  const int ch = warp & ((1 << chunk_shift) - 1), cc = ch * 32 + lane;
  if (cc >= nch) return;
  const size_t plane = (size_t)g.batch * g.H * g.W * C;
  const char* fc = reinterpret_cast<const char*>(lv.feat[g.level] + plane) + (unsigned)cc * 16u;
  float4* dst = outk + cc;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // (An L2 prefetch of the warp's next bin was measured and is slower: 0.165 vs 0.154 ms -- the extra requests compete
  //  with the demand loads.)
  for (int b = warp >> chunk_shift; b < nb; b += nwarps >> chunk_shift) {
    const int c = cnt[b];
    const uint4* e = reinterpret_cast<const uint4*>(ent + b * T);     // two entries per 16 bytes
    float4 acc = zero;
    switch (c >> 1) {                                     // warp-uniform; every case is branch-free straight-line code
      case 1: acc = taplist_consume<1>(e, fc, acc); break;
      case 2: acc = taplist_consume<2>(e, fc, acc); break;
      case 3: acc = taplist_consume<3>(e, fc, acc); break;
      case 4: acc = taplist_consume<4>(e, fc, acc); break;
      case 5: acc = taplist_consume<5>(e, fc, acc); break;
      case 6: acc = taplist_consume<6>(e, fc, acc); break;
      case 7: acc = taplist_consume<7>(e, fc, acc); break;
      case 8: acc = taplist_consume<8>(e, fc, acc); break;
      default: break;                                     // c == 0: every sample of the bin was rejected
    }
    __stcs(dst + (size_t)b * nch, acc);
  }
  }
}

// ---------------------------------------------------------------------------------------
// Gather (output-stationary) backward.  The scatter kernel above is bound by the SM -> L2
// reduction path (ncu: l1tex throughput 84 %, lts red sectors 52 % of peak) and needs the
// gradient maps zero-filled first.  Here every feature pixel is written exactly once:
//   1. tap_gen   one thread per sample point: 4 x (pixel id, weight / count); tap id = position
//   2. bucket    the taps by pixel, either
//        counting : per-pixel tap counts by atomics -> exclusive scan -> scatter (fast; order of a
//                   pixel's taps, hence the float summation order, depends on atomic timing, like the
//                   reference's atomicAdd backward), or
//        radix    : stable radix sort of (pixel id -> tap id) -- bit-reproducible gradients
//   3. gather    one warp per pixel: sum_i w_i * grad_out[row_i, :] with 128-bit loads, one
//                coalesced C*4-byte store per pixel (zeros where nothing taps) -- the zero-fill
//                of roi_align.py:63-64 / roi_align_kernel_v2.cu:325-326 is fused away.
// CUB (scan / radix sort) is the only library code.
struct GatherLevels {
  float* grad[kMaxLevels];
  int H[kMaxLevels];
  int W[kMaxLevels];
  float scale[kMaxLevels];
  unsigned first[kMaxLevels + 1];       // first global pixel id of each level; [n_levels] = total
};

template <bool COUNT, bool IDS>
__global__ void __launch_bounds__(256)
rroi_tap_gen_kernel(GatherLevels lv, int n_levels, int N, const float* __restrict__ rois, int roi_fmt,
                    const int* __restrict__ roi_level, int ph, int pw, int sample_num, int variant,
                    uint4* __restrict__ keys, uint4* __restrict__ ids, float4* __restrict__ wts,
                    unsigned* __restrict__ counts) {
  __shared__ RoiGeom g;
  const int k = blockIdx.x;
  if (threadIdx.x == 0) {
    const float* r = rois + (size_t)k * roi_fmt;
    int lvl = roi_level ? roi_level[k] : 0;
    lvl = min(max(lvl, 0), n_levels - 1);
    g.level = lvl; g.H = lv.H[lvl]; g.W = lv.W[lvl];
    g.batch = (int)r[0];
    decode_roi(r, roi_fmt, lv.scale[lvl], variant, ph, pw, sample_num, g);
  }
  __syncthreads();
  const int nq = ph * pw * sample_num * sample_num;         // gh = gw = sample_num on this path
  const unsigned none = lv.first[n_levels];                 // sentinel: behind every pixel
  const bool batch_ok = g.batch >= 0 && g.batch < N;
  const unsigned base = lv.first[g.level] + (unsigned)g.batch * (unsigned)(g.H * g.W);
  const float inv_count = g.inv_count;
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    const SampleTap t = make_sample(g, q, pw);
    const bool ok = batch_ok && t.off[0] >= 0;
    const size_t e = (size_t)k * nq + q;
    const uint4 key = ok ? make_uint4(base + t.off[0], base + t.off[1], base + t.off[2], base + t.off[3])
                         : make_uint4(none, none, none, none);
    keys[e] = key;
    wts[e] = make_float4(t.w[0] * inv_count, t.w[1] * inv_count, t.w[2] * inv_count, t.w[3] * inv_count);
    if (COUNT && ok) {
      atomicAdd(counts + key.x, 1u); atomicAdd(counts + key.y, 1u); atomicAdd(counts + key.z, 1u); atomicAdd(counts + key.w, 1u);
    }
    if (IDS) {
      const unsigned id = (unsigned)(e * 4);
      ids[e] = make_uint4(id, id + 1, id + 2, id + 3);
    }
  }
}

// Same table with the taps of each bin MERGED by pixel (sample_num 1 or 2: T = 4, 16 taps per bin live in T
// consecutive lanes).  At the FPN level a RoI is mapped to, the 2x2 samples of a bin are 1-2 px apart, so its 16
// taps touch ~9 distinct pixels (57 % on config C3); all taps of a bin read the SAME grad_out row, so duplicates
// collapse into one tap whose weight is the sum of theirs (in tap order: deterministic).  Dropped duplicates get
// the `none` key like rejected samples.  Fewer taps to bucket, and the gather kernel -- bound by the L1 data
// pipe, one 512-byte warp load per tap -- moves 43 % fewer rows.
template <int G, bool IDS>
__global__ void __launch_bounds__(256)
rroi_tap_gen_merged_kernel(GatherLevels lv, int n_levels, int N, const float* __restrict__ rois, int roi_fmt,
                           const int* __restrict__ roi_level, int ph, int pw, int variant,
                           unsigned* __restrict__ keys, unsigned* __restrict__ ids, float* __restrict__ wts,
                           unsigned* __restrict__ counts) {
  constexpr int S = G * G, T = 4 * S, BPW = 32 / T;
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  RoiGeom g;                                        // every thread decodes the RoI (uniform)
  {
    const float* r = rois + (size_t)k * roi_fmt;
    int lvl = roi_level ? __ldg(roi_level + k) : 0;
    lvl = min(max(lvl, 0), n_levels - 1);
    g.level = lvl; g.H = lv.H[lvl]; g.W = lv.W[lvl];
    float rr[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) rr[i] = (i < roi_fmt) ? __ldg(r + i) : 0.f;
    g.batch = (int)rr[0];
    decode_roi(rr, roi_fmt, lv.scale[lvl], variant, ph, pw, G, g);
  }
  const int nbins = ph * pw;
  const unsigned none = lv.first[n_levels];                 // sentinel: behind every pixel
  const bool batch_ok = g.batch >= 0 && g.batch < N;
  const unsigned base = lv.first[g.level] + (unsigned)g.batch * (unsigned)(g.H * g.W);
  const float inv_count = g.inv_count;
  const int i_tap = lane & (T - 1);
  for (int bw = warp * BPW; bw < nbins; bw += nwarps * BPW) {       // warp uniform
    const int b = bw + lane / T;
    const bool valid = b < nbins;
    int off = -1; float w = 0.f;
    if (valid && batch_ok) {
      off = sample_tap_fixed<G>(g, b, i_tap >> 2, i_tap & 3, pw, w);
      w = (off >= 0) ? w * inv_count : 0.f;
    }
    const bool tap_ok = off >= 0;
    // equal keys <=> same pixel of the same bin (32-bit MATCH: the pixel index of a plane is < 2^26, checked on the
    // host); rejected / idle lanes get unique keys
    const unsigned mkey = tap_ok ? (((unsigned)off << 5) | (unsigned)(lane / T)) : (0x80000000u | (unsigned)lane);
    const unsigned m = __match_any_sync(0xffffffffu, mkey);
    const bool f = tap_ok && (lane == __ffs(m) - 1);
    const float ws = group_sum_ordered(m, w);                       // duplicates summed in tap order
    if (valid) {
      const size_t i = ((size_t)k * nbins + b) * T + i_tap;
      const unsigned key = f ? base + (unsigned)off : none;
      keys[i] = key;
      wts[i] = ws;
      if (f) atomicAdd(counts + key, 1u);
      if (IDS) ids[i] = (unsigned)i;
    }
  }
}

// counting variant: slot of tap i inside its pixel's segment = begin + (remaining count - 1)
__global__ void __launch_bounds__(256)
rroi_tap_scatter_kernel(const unsigned* __restrict__ keys, const float* __restrict__ wts, unsigned n_taps, unsigned n_pix,
                        unsigned taps_per_row, const unsigned* __restrict__ seg_begin, unsigned* __restrict__ counts,
                        uint2* __restrict__ sorted) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_taps) return;
  const unsigned key = keys[i];
  if (key >= n_pix) return;
  const unsigned pos = seg_begin[key] + atomicSub(counts + key, 1u) - 1u;
  sorted[pos] = make_uint2(i / taps_per_row, __float_as_uint(wts[i]));
}

// radix variant: the stable sort already put tap ids in (pixel, tap id) order; rejected taps sort last,
// so position i IS the slot and seg_begin (from the same counts + scan) indexes it.
__global__ void __launch_bounds__(256)
rroi_tap_apply_kernel(const unsigned* __restrict__ ids, const float* __restrict__ wts, unsigned n_taps,
                      unsigned taps_per_row, uint2* __restrict__ sorted) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_taps) return;
  const unsigned id = ids[i];
  sorted[i] = make_uint2(id / taps_per_row, __float_as_uint(wts[id]));
}

// One warp per 4 consecutive pixels.  seg_begin is a prefix array (n_pix + 1 entries), so the taps of
// the warp's pixels are ONE contiguous run of (row, weight) pairs: a single small load for the 5 bounds,
// one coalesced load per 32 taps (broadcast by shuffles), then four grad_out rows (8 x LDG.128 per lane
// at C = 256) in flight; neighbouring pixels share most rows -> L1 hits.  The kernel is latency bound
// (ncu: no memory pipe above 45 %), so CTAs are only kGatherWarps warps: a CTA slot frees as soon as its
// few warps are done, which keeps the resident-warp count up although the tap count per pixel varies.
constexpr int kGatherWarps = 1;

// M taps of one pixel (M = 1..4, compile time): all loads issued before the first FMA, no predicates, no FMA for absent taps
// (r1 ran every step as four predicated taps: at 2.6 merged taps per pixel a third of its FFMAs multiplied by zero).
template <int M>
__device__ __forceinline__ void gather_taps(const float4* __restrict__ grad_out, int nch, int cc, bool one, bool two, uint2 mine,
                                            int rel, float4& a0, float4& a1) {
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 v[M], u[M]; float w[M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const unsigned row = __shfl_sync(0xffffffffu, mine.x, rel + i);
    w[i] = __uint_as_float(__shfl_sync(0xffffffffu, mine.y, rel + i));
    const float4* r = grad_out + (size_t)row * nch + cc;
    v[i] = one ? __ldg(r) : zero;
    u[i] = two ? __ldg(r + 32) : zero;
  }
#pragma unroll
  for (int i = 0; i < M; ++i) { a0 = vfma(w[i], v[i], a0); a1 = vfma(w[i], u[i], a1); }
}

#ifndef AIDET_ROI_GATHER_OCC
#define AIDET_ROI_GATHER_OCC 26
#endif
template <int PX>
__global__ void __launch_bounds__(kGatherWarps * 32, AIDET_ROI_GATHER_OCC)
rroi_gather_kernel(GatherLevels lv, int n_levels, int C, const float4* __restrict__ grad_out,
                   const uint2* __restrict__ sorted, const unsigned* __restrict__ seg_begin) {
  const unsigned n_pix = lv.first[n_levels];
  const unsigned p0 = (blockIdx.x * (unsigned)kGatherWarps + (threadIdx.x >> 5)) * (unsigned)PX;
  const int lane = threadIdx.x & 31;
  if (p0 >= n_pix) return;
  const int np = (int)min((unsigned)PX, n_pix - p0);
  unsigned bnd_l = 0;
  if (lane <= np) bnd_l = __ldg(seg_begin + p0 + lane);
  const unsigned run_begin = __shfl_sync(0xffffffffu, bnd_l, 0), run_end = __shfl_sync(0xffffffffu, bnd_l, np);
  const int nch = C >> 2;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // level of the warp's first pixel; its pixels are consecutive, so the level changes at most at lv.first[l0 + 1]
  int l0 = 0;
#pragma unroll
  for (int q = 1; q < kMaxLevels; ++q) if (q < n_levels && p0 >= lv.first[q]) l0 = q;
  const unsigned l0_end = lv.first[l0 + 1];
  for (int cb = 0; cb < nch; cb += 64) {            // every lane runs the loops: the shuffles need the full warp
    const int cc = cb + lane;
    const bool one = cc < nch, two = cc + 32 < nch;
    unsigned chunk = run_begin;
    uint2 mine = make_uint2(0u, 0u);
    if (chunk + lane < run_end) mine = __ldg(sorted + chunk + lane);
#pragma unroll 1
    for (int j = 0; j < np; ++j) {
      const unsigned pix = p0 + j;
      int l = l0;
      if (pix >= l0_end) {                          // rare: the warp's pixels straddle two levels
#pragma unroll
        for (int q = 1; q < kMaxLevels; ++q) if (q < n_levels && pix >= lv.first[q]) l = q;
      }
      float4* dst = reinterpret_cast<float4*>(lv.grad[l]) + (size_t)(pix - lv.first[l]) * nch;
      float4 a0 = zero, a1 = zero;
      unsigned t = __shfl_sync(0xffffffffu, bnd_l, j);
      const unsigned te = __shfl_sync(0xffffffffu, bnd_l, j + 1);
      while (t < te) {
        if (t >= chunk + 32u) {                     // next 32 taps of the run
          chunk = t;
          mine = make_uint2(0u, 0u);
          if (chunk + lane < run_end) mine = __ldg(sorted + chunk + lane);
        }
        const int rel = (int)(t - chunk);
        const int m = (int)min(min(4u, te - t), 32u - (unsigned)rel);      // warp uniform
        switch (m) {
          case 1: gather_taps<1>(grad_out, nch, cc, one, two, mine, rel, a0, a1); break;
          case 2: gather_taps<2>(grad_out, nch, cc, one, two, mine, rel, a0, a1); break;
          case 3: gather_taps<3>(grad_out, nch, cc, one, two, mine, rel, a0, a1); break;
          default: gather_taps<4>(grad_out, nch, cc, one, two, mine, rel, a0, a1); break;
        }
        t += (unsigned)m;
      }
      if (one) __stcs(dst + cc, a0);
      if (two) __stcs(dst + cc + 32, a1);
    }
  }
}

struct GatherLayout { size_t keys_in, keys_out, ids_in, ids_out, wts, sorted, seg_begin, seg_end, cub, total; size_t cub_bytes; };

static int gather_cub_bytes(size_t n_taps, size_t n_pix, int end_bit, size_t* bytes) {
  size_t a = 0, b = 0;
  AIDET_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned*)nullptr, (unsigned*)nullptr,
                                             (const unsigned*)nullptr, (unsigned*)nullptr, (int)n_taps, 0, end_bit,
                                             (cudaStream_t)0));
  AIDET_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b, (const unsigned*)nullptr, (unsigned*)nullptr, (int)(n_pix + 1),
                                           (cudaStream_t)0));
  *bytes = a > b ? a : b;
  return AIDET_OK;
}

static GatherLayout gather_layout(size_t n_taps, size_t n_pix, size_t cub_bytes) {
  GatherLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  L.keys_in = take(n_taps * 4); L.keys_out = take(n_taps * 4);
  L.ids_in = take(n_taps * 4); L.ids_out = take(n_taps * 4);
  L.wts = take(n_taps * 4);
  L.sorted = take(n_taps * 8);
  L.seg_begin = take((n_pix + 1) * 4); L.seg_end = take((n_pix + 1) * 4);
  L.cub_bytes = cub_bytes; L.cub = take(cub_bytes);
  L.total = off + 256;
  return L;
}

static int bits_for(unsigned long long n) { int b = 1; while ((1ULL << b) <= n) ++b; return b; }

static int check_common(const int* H, const int* W, const float* scale, int n_levels, int N, int C, const float* rois,
                        int roi_fmt, const int* roi_level, int K, int ph, int pw, int sample_num, int variant) {
  AIDET_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels, "rroi_align: n_levels must be in [1,%d], got %d", kMaxLevels, n_levels);
  AIDET_REQUIRE(H && W && scale, "rroi_align: null level tables");
  AIDET_REQUIRE(N >= 1 && C >= 1 && K >= 0 && ph >= 1 && pw >= 1, "rroi_align: bad sizes N=%d C=%d K=%d ph=%d pw=%d", N, C, K, ph, pw);
  AIDET_REQUIRE(roi_fmt == 5 || roi_fmt == 6, "rroi_align: wrong roi size %d (expected 5 or 6)", roi_fmt);
  AIDET_REQUIRE(variant >= 0 && variant <= 2, "rroi_align: bad variant %d", variant);
  AIDET_REQUIRE(sample_num >= 0, "rroi_align: negative sample_num");
  AIDET_REQUIRE(K == 0 || rois, "rroi_align: null rois");
  AIDET_REQUIRE(n_levels == 1 || roi_level || K == 0, "rroi_align: roi_level required with several levels");
  for (int l = 0; l < n_levels; l++) AIDET_REQUIRE(H[l] >= 1 && W[l] >= 1, "rroi_align: bad level %d size", l);
  return AIDET_OK;
}

// Forward default = the tap-list kernel (merged taps).  History on C3: one load per tap 0.176 ms (on the L1 wavefront
// bound); first merged version (r1b) 0.307 ms (merge prologue per bin row + 35 SASS instructions per entry); tap list
// with branch-free per-count consumers 0.154 ms: 61.7 M warp instructions instead of 105.7 M, 57.6 M L1 sectors
// instead of 98.2 M, now limited by load latency / DRAM efficiency of scattered 1 KB reads (profiles/r1e).
// Kernel-variant switches are COMPILE-TIME constants (the library keeps no environment-dependent state); a tuning build
// overrides them with -D (scripts/build_variants.sh).  Defaults = the variants measured fastest on config C3.
#ifndef AIDET_ROI_FWD_UNMERGED
#define AIDET_ROI_FWD_UNMERGED 0      // 1: per-sample forward (rroi_align_fast_kernel) instead of the tap-list kernel
#endif
#ifndef AIDET_ROI_BWD_UNMERGED
#define AIDET_ROI_BWD_UNMERGED 0      // 1: one tap per sample corner in the gather backward, no per-bin merge
#endif
#ifndef AIDET_ROI_GATHER_PX
#define AIDET_ROI_GATHER_PX 8         // pixels per gather warp: 4, 8 or 16
#endif
#ifndef AIDET_ROI_TL_OCC
#define AIDET_ROI_TL_OCC 5            // resident CTAs per SM of the one-chunk tap-list forward: 4, 5 or 6
#endif
#ifndef AIDET_ROI_TL_PAIR
#define AIDET_ROI_TL_PAIR 1           // two channel chunks per warp when C % 256 == 0
#endif
constexpr bool g_roi_fwd_unmerged = AIDET_ROI_FWD_UNMERGED != 0;
constexpr bool g_roi_bwd_unmerged = AIDET_ROI_BWD_UNMERGED != 0;
constexpr int g_gather_px = (AIDET_ROI_GATHER_PX == 4 || AIDET_ROI_GATHER_PX == 16) ? AIDET_ROI_GATHER_PX : 8;
constexpr int g_tl_occ = AIDET_ROI_TL_OCC;
constexpr bool g_tl_pair = AIDET_ROI_TL_PAIR != 0;

static int lanes_for(int nch) { int cl = 1; while (cl < nch && cl < 256) cl <<= 1; return cl; }

// Upper bound of the sampling-grid size over all RoIs is data dependent (sample_num = 0 means
// ceil(roi/bin) points); the fast kernel needs S <= kTable, which sample_num > 0 guarantees.
template <bool BWD>
static int launch(RoiLevels& lv, int n_levels, int N, int C, const float* rois, int roi_fmt, const int* roi_level, int K,
                  int ph, int pw, int sample_num, int variant, float* io, bool vec4, cudaStream_t s) {
  if (K == 0) return AIDET_OK;
  ProfScope prof(BWD ? PROF_ROI_BWD : PROF_ROI_FWD, s);
  long long max_plane = 0;
  for (int l = 0; l < n_levels; l++) max_plane = max(max_plane, (long long)lv.H[l] * lv.W[l]);
  const bool fast = vec4 && sample_num > 0 && sample_num * sample_num <= kTable && max_plane * C * 4 < 0xffffffffLL;
  if (fast && !BWD && sample_num <= 2 && !g_roi_fwd_unmerged && max_plane < (1LL << 28) && C <= 1024) {
    const int nbins = ph * pw;
    const int groups_per_roi = ceil_div(nbins, kTLBins);
    const int bins_per_group = ceil_div(nbins, groups_per_roi);
    int chunk_shift = 0;                                  // 2^chunk_shift warps share a bin: 32 channel quads each
    while ((32 << chunk_shift) < C / 4) ++chunk_shift;    // C <= 1024 -> <= 8 chunks = the CTA's 8 warps
    const bool pair = g_tl_pair && chunk_shift >= 1 && C % 256 == 0;   // a warp reads both 512-byte halves of a pixel
#define AIDET_LAUNCH_TAPLIST(G_, O_, P_)                                                                       \
  rroi_align_fwd_taplist_kernel<G_, O_, P_><<<K * groups_per_roi, 256, 0, s>>>(lv, n_levels, N, C, rois, roi_fmt, roi_level, \
                                                                              ph, pw, variant, io, groups_per_roi, bins_per_group, chunk_shift)
    if (sample_num == 2) {
      if (pair) AIDET_LAUNCH_TAPLIST(2, 4, true);
      else if (g_tl_occ == 4) AIDET_LAUNCH_TAPLIST(2, 4, false);
      else if (g_tl_occ == 6) AIDET_LAUNCH_TAPLIST(2, 6, false);
      else AIDET_LAUNCH_TAPLIST(2, 5, false);
    } else {
      if (pair) AIDET_LAUNCH_TAPLIST(1, 4, true); else AIDET_LAUNCH_TAPLIST(1, 5, false);
    }
#undef AIDET_LAUNCH_TAPLIST
  } else if (fast) {
    // The kernel is latency bound (ncu: no memory pipe above 65 %), so the grid is made of many small
    // CTAs -- one bin row of one RoI, channel lanes only -- that come and go independently: a CTA slot
    // frees as soon as its one or two warps are done and only they wait at the table barrier.
    const int nch = C / 4;
    const int CL = lanes_for(nch);                      // power of two >= nch, <= 256
    const int threads = max(CL, 32);
    const int nslots = threads / CL;                    // > 1 only when C < 128
    const int groups_per_roi = ph, bins_per_group = pw;
    rroi_align_fast_kernel<BWD><<<K * groups_per_roi, threads, 0, s>>>(lv, n_levels, N, C, rois, roi_fmt, roi_level, ph, pw,
                                                                       sample_num, variant, io, CL, nslots, groups_per_roi,
                                                                       bins_per_group);
  } else if (vec4) {
    rroi_align_kernel<4, BWD><<<K, 256, 0, s>>>(lv, n_levels, N, C, rois, roi_fmt, roi_level, ph, pw, sample_num,
                                                variant, io, lanes_for(C / 4));
  } else {
    rroi_align_kernel<1, BWD><<<K, 256, 0, s>>>(lv, n_levels, N, C, rois, roi_fmt, roi_level, ph, pw, sample_num,
                                                variant, io, lanes_for(C));
  }
  count_launch(1);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

}  // namespace aidet

using namespace aidet;

extern "C" {

int aidet_rroi_align_fwd_f32(const float* const* feat_host, const int* H_host, const int* W_host,
                             const float* scale_host, int n_levels, int N, int C, const float* rois, int roi_fmt,
                             const int* roi_level, int K, int ph, int pw, int sample_num, int variant, float* out,
                             int device, void* stream) {
  if (int rc = check_common(H_host, W_host, scale_host, n_levels, N, C, rois, roi_fmt, roi_level, K, ph, pw, sample_num, variant)) return rc;
  AIDET_REQUIRE(feat_host && (out || K == 0), "aidet_rroi_align_fwd_f32: null pointer");
  RoiLevels lv{};
  bool vec4 = (C % 4 == 0) && (((uintptr_t)out & 15) == 0);
  for (int l = 0; l < n_levels; l++) {
    AIDET_REQUIRE(feat_host[l], "aidet_rroi_align_fwd_f32: null feature pointer at level %d", l);
    lv.feat[l] = feat_host[l]; lv.grad[l] = nullptr; lv.H[l] = H_host[l]; lv.W[l] = W_host[l]; lv.scale[l] = scale_host[l];
    vec4 = vec4 && (((uintptr_t)feat_host[l] & 15) == 0);
  }
  if (int rc = set_device(device)) return rc;
  return launch<false>(lv, n_levels, N, C, rois, roi_fmt, roi_level, K, ph, pw, sample_num, variant, out, vec4, (cudaStream_t)stream);
}

int aidet_rroi_align_bwd_f32(const float* grad_out, float* const* grad_feat_host, const int* H_host, const int* W_host,
                             const float* scale_host, int n_levels, int N, int C, const float* rois, int roi_fmt,
                             const int* roi_level, int K, int ph, int pw, int sample_num, int variant, int device,
                             void* stream) {
  if (int rc = check_common(H_host, W_host, scale_host, n_levels, N, C, rois, roi_fmt, roi_level, K, ph, pw, sample_num, variant)) return rc;
  AIDET_REQUIRE(grad_feat_host && (grad_out || K == 0), "aidet_rroi_align_bwd_f32: null pointer");
  RoiLevels lv{};
  bool vec4 = (C % 4 == 0) && (((uintptr_t)grad_out & 15) == 0);
  for (int l = 0; l < n_levels; l++) {
    AIDET_REQUIRE(grad_feat_host[l], "aidet_rroi_align_bwd_f32: null gradient pointer at level %d", l);
    lv.feat[l] = nullptr; lv.grad[l] = grad_feat_host[l]; lv.H[l] = H_host[l]; lv.W[l] = W_host[l]; lv.scale[l] = scale_host[l];
    vec4 = vec4 && (((uintptr_t)grad_feat_host[l] & 15) == 0);
  }
  if (int rc = set_device(device)) return rc;
  return launch<true>(lv, n_levels, N, C, rois, roi_fmt, roi_level, K, ph, pw, sample_num, variant,
                      const_cast<float*>(grad_out), vec4, (cudaStream_t)stream);
}

/* The gather path needs sample_num > 0 (fixed tap count), C % 4 == 0 and 16 B aligned pointers. */
size_t aidet_rroi_align_bwd_workspace_bytes(const int* H_host, const int* W_host, int n_levels, int N, int K, int ph,
                                            int pw, int sample_num) {
  if (!H_host || !W_host || n_levels < 1 || n_levels > kMaxLevels || sample_num <= 0 || K <= 0) return 0;
  unsigned long long n_pix = 0;
  for (int l = 0; l < n_levels; l++) n_pix += (unsigned long long)N * H_host[l] * W_host[l];
  const unsigned long long n_taps = 4ULL * K * ph * pw * sample_num * sample_num;
  if (n_pix >= 0x7fffffffULL || n_taps >= 0x7fffffffULL) return 0;
  size_t cub_bytes = 0;
  if (gather_cub_bytes((size_t)n_taps, (size_t)n_pix, bits_for(n_pix), &cub_bytes) != AIDET_OK) return 0;
  return gather_layout((size_t)n_taps, (size_t)n_pix, cub_bytes).total;
}

int aidet_rroi_align_bwd_gather_f32(const float* grad_out, float* const* grad_feat_host, const int* H_host,
                                    const int* W_host, const float* scale_host, int n_levels, int N, int C,
                                    const float* rois, int roi_fmt, const int* roi_level, int K, int ph, int pw,
                                    int sample_num, int variant, int deterministic, void* workspace, size_t ws_bytes,
                                    int device, void* stream) {
  if (int rc = check_common(H_host, W_host, scale_host, n_levels, N, C, rois, roi_fmt, roi_level, K, ph, pw, sample_num, variant)) return rc;
  AIDET_REQUIRE(grad_feat_host && (grad_out || K == 0), "aidet_rroi_align_bwd_gather_f32: null pointer");
  AIDET_REQUIRE(sample_num > 0, "aidet_rroi_align_bwd_gather_f32: needs sample_num > 0 (adaptive grids use aidet_rroi_align_bwd_f32)");
  AIDET_REQUIRE(C % 4 == 0 && (((uintptr_t)grad_out & 15) == 0), "aidet_rroi_align_bwd_gather_f32: needs C %% 4 == 0 and 16 B aligned grad_out");
  GatherLevels lv{};
  unsigned long long n_pix = 0;
  for (int l = 0; l < n_levels; l++) {
    AIDET_REQUIRE(grad_feat_host[l] && (((uintptr_t)grad_feat_host[l] & 15) == 0), "aidet_rroi_align_bwd_gather_f32: bad gradient pointer at level %d", l);
    lv.grad[l] = grad_feat_host[l]; lv.H[l] = H_host[l]; lv.W[l] = W_host[l]; lv.scale[l] = scale_host[l];
    lv.first[l] = (unsigned)n_pix;
    n_pix += (unsigned long long)N * H_host[l] * W_host[l];
  }
  const unsigned long long n_taps = 4ULL * (unsigned long long)K * ph * pw * sample_num * sample_num;
  AIDET_REQUIRE(n_pix < 0x7fffffffULL && n_taps < 0x7fffffffULL, "aidet_rroi_align_bwd_gather_f32: problem too large");
  lv.first[n_levels] = (unsigned)n_pix;
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (K == 0) {        // nothing taps anything: the gradient is all zeros
    for (int l = 0; l < n_levels; l++)
      AIDET_CUDA(cudaMemsetAsync(grad_feat_host[l], 0, (size_t)N * H_host[l] * W_host[l] * C * sizeof(float), s));
    return AIDET_OK;
  }
  AIDET_REQUIRE(workspace && (((uintptr_t)workspace & 255) == 0), "aidet_rroi_align_bwd_gather_f32: workspace must be 256 B aligned");
  const int end_bit = bits_for(n_pix);
  size_t cub_bytes = 0;
  if (int rc = gather_cub_bytes((size_t)n_taps, (size_t)n_pix, end_bit, &cub_bytes)) return rc;
  const GatherLayout L = gather_layout((size_t)n_taps, (size_t)n_pix, cub_bytes);
  if (ws_bytes < L.total) { set_error("aidet_rroi_align_bwd_gather_f32: workspace %zu < %zu", ws_bytes, L.total); return AIDET_EWORKSPACE; }
  char* ws = (char*)workspace;
  unsigned* keys_in = (unsigned*)(ws + L.keys_in); unsigned* keys_out = (unsigned*)(ws + L.keys_out);
  unsigned* ids_in = (unsigned*)(ws + L.ids_in); unsigned* ids_out = (unsigned*)(ws + L.ids_out);
  float* wts = (float*)(ws + L.wts);
  uint2* sorted = (uint2*)(ws + L.sorted);
  unsigned* seg_begin = (unsigned*)(ws + L.seg_begin); unsigned* seg_end = (unsigned*)(ws + L.seg_end);
  const unsigned tpr = 4u * sample_num * sample_num;
  const unsigned tap_blocks = (unsigned)((n_taps + 255) / 256);
  ProfScope prof(PROF_ROI_BWD, s);
  size_t cb = L.cub_bytes;
  unsigned* counts = seg_end;                                      // (n_pix + 1) words
  AIDET_CUDA(cudaMemsetAsync(counts, 0, (size_t)(n_pix + 1) * 4, s));
  long long max_plane = 0;
  for (int l = 0; l < n_levels; l++) max_plane = max(max_plane, (long long)H_host[l] * W_host[l]);
  if (sample_num <= 2 && !g_roi_bwd_unmerged && max_plane < (1LL << 26)) {
#define AIDET_TAPGEN(G_, IDS_)                                                                                   \
  rroi_tap_gen_merged_kernel<G_, IDS_><<<K, 128, 0, s>>>(lv, n_levels, N, rois, roi_fmt, roi_level, ph, pw, variant, \
                                                         keys_in, ids_in, wts, counts)
    if (sample_num == 2) { if (deterministic) AIDET_TAPGEN(2, true); else AIDET_TAPGEN(2, false); }
    else                 { if (deterministic) AIDET_TAPGEN(1, true); else AIDET_TAPGEN(1, false); }
#undef AIDET_TAPGEN
  } else if (deterministic)
    rroi_tap_gen_kernel<true, true><<<K, 256, 0, s>>>(lv, n_levels, N, rois, roi_fmt, roi_level, ph, pw, sample_num, variant,
                                                      (uint4*)keys_in, (uint4*)ids_in, (float4*)wts, counts);
  else
    rroi_tap_gen_kernel<true, false><<<K, 256, 0, s>>>(lv, n_levels, N, rois, roi_fmt, roi_level, ph, pw, sample_num, variant,
                                                       (uint4*)keys_in, nullptr, (float4*)wts, counts);
  AIDET_CUDA(cub::DeviceScan::ExclusiveSum(ws + L.cub, cb, counts, seg_begin, (int)(n_pix + 1), s));
  if (deterministic) {
    cb = L.cub_bytes;
    AIDET_CUDA(cub::DeviceRadixSort::SortPairs(ws + L.cub, cb, keys_in, keys_out, ids_in, ids_out, (int)n_taps, 0, end_bit, s));
    rroi_tap_apply_kernel<<<tap_blocks, 256, 0, s>>>(ids_out, wts, (unsigned)n_taps, tpr, sorted);
  } else {
    rroi_tap_scatter_kernel<<<tap_blocks, 256, 0, s>>>(keys_in, wts, (unsigned)n_taps, (unsigned)n_pix, tpr, seg_begin,
                                                       counts, sorted);
  }
  const unsigned px = (unsigned)g_gather_px * kGatherWarps;
#define AIDET_GATHER(PX_)                                                                           \
  rroi_gather_kernel<PX_><<<(unsigned)((n_pix + px - 1) / px), kGatherWarps * 32, 0, s>>>(            \
      lv, n_levels, C, (const float4*)grad_out, sorted, seg_begin)
  if (g_gather_px == 8) AIDET_GATHER(8); else if (g_gather_px == 16) AIDET_GATHER(16); else AIDET_GATHER(4);
#undef AIDET_GATHER
  count_launch(3);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

}  // extern "C"
