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
//   - one CTA per RoI; the sample geometry (4 tap offsets + 4 weights per sample point,
//     border rules of roi_align_kernel.cu:17-62) is computed ONCE per RoI into shared
//     memory and broadcast to the channel lanes
//   - all FPN levels run in one launch (per-RoI level id), rotated and axis-aligned RoIs
//     share the kernel (theta = 0 reproduces v1/v2)
//   - backward scatters with 128-bit vector reductions (red.global.add.v4.f32)
// Bound: HBM/L2 bandwidth (no reuse of arithmetic; ~2 flop per byte read).
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
    double sn, cs;                          // once per RoI: keep sin/cos at <= 0.5 ulp
    sincos((double)r[5], &sn, &cs);
    g.cs = (float)cs; g.sn = (float)sn; g.xb = -0.5f * roi_w; g.yb = -0.5f * roi_h;
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
__device__ __forceinline__ SampleTap make_sample(const RoiGeom& g, int q, int pw) {
  const int S = g.gh * g.gw;
  const int bin = q / S, rem = q - bin * S;
  const int iy = rem / g.gw, ix = rem - iy * g.gw;
  const int p_h = bin / pw, p_w = bin - p_h * pw;
  float yy = g.yb + p_h * g.bin_h + (iy + .5f) * g.bin_h / (float)g.gh;
  float xx = g.xb + p_w * g.bin_w + (ix + .5f) * g.bin_w / (float)g.gw;
  float x = g.ox + xx * g.cs - yy * g.sn;
  float y = g.oy + xx * g.sn + yy * g.cs;
  SampleTap t;
  if (y < -1.0f || y > (float)g.H || x < -1.0f || x > (float)g.W) {
    t.off[0] = -1; t.off[1] = t.off[2] = t.off[3] = 0;
    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
    return t;
  }
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= g.H - 1) { yh = yl = g.H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= g.W - 1) { xh = xl = g.W - 1; x = (float)xl; } else xh = xl + 1;
  float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
  t.off[0] = yl * g.W + xl; t.off[1] = yl * g.W + xh; t.off[2] = yh * g.W + xl; t.off[3] = yh * g.W + xh;
  t.w[0] = hy * hx; t.w[1] = hy * lx; t.w[2] = ly * hx; t.w[3] = ly * lx;
  return t;
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

static int lanes_for(int nch) { int cl = 1; while (cl < nch && cl < 256) cl <<= 1; return cl; }

template <bool BWD>
static int launch(RoiLevels& lv, int n_levels, int N, int C, const float* rois, int roi_fmt, const int* roi_level, int K,
                  int ph, int pw, int sample_num, int variant, float* io, bool vec4, cudaStream_t s) {
  if (K == 0) return AIDET_OK;
  ProfScope prof(BWD ? PROF_ROI_BWD : PROF_ROI_FWD, s);
  if (vec4) {
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

}  // extern "C"
