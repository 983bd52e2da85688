// soft_nms.cu -- Soft-NMS on the device, batched over groups (one CTA per group), sm_100a.
//
// Replaces (reference): soft_nms_cpu_kernel, mmdet/ops/nms/src/nms_cpu.cpp:70-201, which is the ONLY
// implementation the reference has -- nms_wrapper.soft_nms (mmdet/ops/nms/nms_wrapper.py:63-118) copies
// CUDA tensors to the host, runs the single-threaded loop and copies the result back (:92-94,110-114).
//
// The algorithm is sequential in the selected box (N dependent steps), so a group runs on ONE CTA and
// the parallelism is inside a step and across groups (classes x images run concurrently).  Every step
// restates the reference's loop body exactly, including its in-place array discipline, so that the
// result rows -- [x1, y1, x2, y2, decayed score, original index], in selection order -- are the same:
//   1. arg-max of the scores at positions [i, n) -- FIRST maximum (`max_score < scores[pos]`, :111-117)
//   2. swap position i with it (:119-135)
//   3. decay every later score (:146-176): ovr with the +1 convention, weight = 1 - ovr if ovr > thr
//      (linear), exp(-ovr^2 / sigma) (gaussian) or 0/1 (hard); IEEE single precision, no FMA
//      contraction, same operation order as the C++ (float instantiation)
//   4. drop boxes whose score fell below min_score.  The reference does it while walking the array:
//      a dead box is overwritten by the LAST box and n shrinks (:179-188).  Walking up from i+1 this
//      means: the k-th dead position below the new n receives the k-th surviving box counted from
//      the end -- a two-pointer partition, done here with one block-wide prefix sum.
// Bound: latency (7 block barriers per step); a 2000-box group takes about as long as the host loop,
// but nothing leaves the device and all groups of a call run at once.
#include "common.cuh"

namespace aidet {

constexpr int kSoftThreads = 1024;

struct SoftBest { float s; int p; };

__device__ __forceinline__ SoftBest soft_better(SoftBest a, SoftBest b) {   // higher score, then lower position
  if (b.p < 0) return a;
  if (a.p < 0) return b;
  if (b.s > a.s || (b.s == a.s && b.p < a.p)) return b;
  return a;
}

// rows: (n_total, 6) [x1, y1, x2, y2, score, original index]; group g owns rows [off[g], off[g+1]).
__global__ void __launch_bounds__(kSoftThreads)
soft_nms_kernel(float* __restrict__ rows_all, const int* __restrict__ off, float thr, int method, float sigma,
                float min_score, int* __restrict__ n_out, int cap) {
  __shared__ SoftBest s_best[32];
  __shared__ int s_scan[32];
  __shared__ int s_total, s_maxpos, s_nholes;
  extern __shared__ int s_list[];                   // [2][cap]: holes, fillers (cap = largest group)
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = off[g];
  int n = off[g + 1] - base;
  float* R = rows_all + (size_t)base * 6;
  int* holes = s_list;
  int* fillers = s_list + cap;

  for (int i = 0; i < n; ++i) {
    // ---- 1. first maximum of scores[i .. n).  The reference starts from max = scores[i] and updates on
    //         `max < scores[pos]`: a NaN at i is never beaten, a NaN elsewhere never wins.
    const float si = R[i * 6 + 4];
    if (tid == 0) { s_maxpos = i; s_nholes = 0; }
    if (si == si) {                                  // CTA uniform
      SoftBest b{0.f, -1};
      for (int p = i + tid; p < n; p += kSoftThreads) {
        const float sc = R[p * 6 + 4];
        if (sc == sc && (b.p < 0 || sc > b.s)) b = SoftBest{sc, p};          // strict >: the first one stays
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        SoftBest o{__shfl_xor_sync(0xffffffffu, b.s, d), __shfl_xor_sync(0xffffffffu, b.p, d)};
        b = soft_better(b, o);
      }
      if (lane == 0) s_best[warp] = b;
      __syncthreads();
      if (warp == 0) {
        SoftBest c = s_best[lane];
#pragma unroll
        for (int d = 16; d; d >>= 1) {
          SoftBest o{__shfl_xor_sync(0xffffffffu, c.s, d), __shfl_xor_sync(0xffffffffu, c.p, d)};
          c = soft_better(c, o);
        }
        if (lane == 0 && c.p >= 0) s_maxpos = c.p;
      }
    }
    __syncthreads();
    // ---- 2. swap i <-> max position
    const int mp = s_maxpos;
    if (tid < 6 && mp != i) {
      const float a = R[i * 6 + tid], c = R[mp * 6 + tid];
      R[i * 6 + tid] = c; R[mp * 6 + tid] = a;
    }
    __syncthreads();
    const float ix1 = R[i * 6 + 0], iy1 = R[i * 6 + 1], ix2 = R[i * 6 + 2], iy2 = R[i * 6 + 3];
    const float iarea = __fmul_rn(__fadd_rn(__fsub_rn(ix2, ix1), 1.f), __fadd_rn(__fsub_rn(iy2, iy1), 1.f));
    // ---- 3. decay scores[i+1 .. n); contiguous chunk per thread so that prefix sums follow array order
    const int m = n - (i + 1);
    const int per = (m + kSoftThreads - 1) / kSoftThreads;
    const int lo = i + 1 + min(m, tid * per), hi = i + 1 + min(m, (tid + 1) * per);
    int dead_mine = 0;
    for (int p = lo; p < hi; ++p) {
      float* r = R + p * 6;
      const float x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3];
      const float area = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
      const float xx1 = fmaxf(ix1, x1), yy1 = fmaxf(iy1, y1), xx2 = fminf(ix2, x2), yy2 = fminf(iy2, y2);
      const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
      const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, area), inter));
      float weight = 1.f;
      if (method == 1) { if (ovr > thr) weight = __fsub_rn(1.f, ovr); }
      else if (method == 2) weight = expf(__fdiv_rn(-__fmul_rn(ovr, ovr), sigma));
      else { weight = (ovr > thr) ? 0.f : 1.f; }
      const float sc = __fmul_rn(weight, r[4]);
      r[4] = sc;
      dead_mine += (sc < min_score) ? 1 : 0;
    }
    // ---- 4. drop the dead boxes (two-pointer partition of (i, n))
    int x = dead_mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int v = s_scan[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += y; }
      s_scan[lane] = v;
      if (lane == 31) s_total = v;
    }
    __syncthreads();
    const int total_dead = s_total;
    if (total_dead > 0) {                            // CTA uniform
      const int new_n = n - total_dead;
      int dead_before = x - dead_mine + (warp ? s_scan[warp - 1] : 0);        // dead boxes at positions < lo
      for (int p = lo; p < hi; ++p) {
        const bool dead = R[p * 6 + 4] < min_score;
        if (dead) {
          if (p < new_n) { holes[dead_before] = p; atomicAdd(&s_nholes, 1); }  // k-th dead position, ascending
          ++dead_before;
        } else if (p >= new_n) {
          const int alive_after = (n - 1 - p) - (total_dead - dead_before);   // survivors behind p
          fillers[alive_after] = p;                                           // k-th survivor from the end
        }
      }
      __syncthreads();
      const int n_holes = s_nholes;                 // dead positions below new_n == survivors at or above it
      for (int k = tid; k < n_holes * 6; k += kSoftThreads) {
        const int h = k / 6, c = k - h * 6;
        R[holes[h] * 6 + c] = R[fillers[h] * 6 + c];
      }
      n = new_n;
    }
    __syncthreads();
  }
  if (tid == 0) n_out[g] = n;
}

}  // namespace aidet

using namespace aidet;

extern "C" {

int aidet_soft_nms_f32(float* rows, const int* group_offsets, int n_groups, int max_group, float iou_thr, int method,
                       float sigma, float min_score, int* n_out, int device, void* stream) {
  AIDET_REQUIRE(n_groups >= 0 && max_group >= 0, "aidet_soft_nms_f32: bad sizes n_groups=%d max_group=%d", n_groups, max_group);
  AIDET_REQUIRE(method >= 0 && method <= 2, "aidet_soft_nms_f32: method must be 0 (hard), 1 (linear) or 2 (gaussian)");
  if (n_groups == 0) return AIDET_OK;
  AIDET_REQUIRE(rows && group_offsets && n_out, "aidet_soft_nms_f32: null pointer");
  AIDET_REQUIRE(max_group <= 65535, "aidet_soft_nms_f32: groups of more than 65535 boxes are not supported (got %d)", max_group);
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int cap = max_group > 0 ? max_group : 1;
  const size_t smem = (size_t)cap * 2 * sizeof(int);
  if (smem > 48 * 1024) {
    AIDET_REQUIRE(smem <= 200 * 1024, "aidet_soft_nms_f32: group of %d boxes exceeds the shared-memory lists", max_group);
    AIDET_CUDA(cudaFuncSetAttribute(soft_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  soft_nms_kernel<<<n_groups, kSoftThreads, smem, s>>>(rows, group_offsets, iou_thr, method, sigma, min_score, n_out, cap);
  count_launch(1);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

}  // extern "C"
