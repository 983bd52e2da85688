// rnms.cu -- batched (image x class) NMS for oriented and axis-aligned boxes on sm_100a.
//
// Replaces (reference):
//   mmdet/ops/nms/src/nms_kernel.cu:24-68   64x64 bitmask tiles (lower triangle computed too)
//   mmdet/ops/nms/src/nms_kernel.cu:71-139  device sort, D2H copy of the whole mask, serial host
//                                           scan, H2D of keeps, cudaMalloc/Free per call
//   mmdet/core/post_processing/rbbox_nms.py:29-49,83-106  Python loop over classes
// with ONE device-side pass over all groups, no host round trip.
// n <= 8192 boxes and <= 1024 groups (every per-image config): ONE cooperative launch, nms_fused_kernel:
//   rank (keys in shared memory, rank by counting) -> mask (warp-level units) -> scan (one CTA per group) -> compaction.
// Larger inputs:
//   1. order   group asc, score desc, original index asc on ties: counting sort by group id (histogram, scan, scatter)
//              + rank-by-counting inside each bucket; inputs whose groups average > 24576 boxes: 64-bit keys + CUB
//              radix sort (the only library call) + gather.  Sorted boxes -> prepared records (geom.cuh), group bounds
//   2. mask    upper-triangle suppression bitmask, 32-bit half-words = __ballot_sync of "IoU > thr" over 32 column
//              boxes held in registers: 64 x 256 tiles handed out by a ticket counter with the row records staged by
//              TMA, or warp-level units when the groups average < 2048 boxes
//   3. scan    one CTA per group walks the rows in score order (greedy), entirely on device
//   4. compact kept flags -> ascending original indices (nms_kernel.cu:135-138 semantics), done by whichever scan CTA
//              finishes last
// Scene form (section 7): per-tile NMS + cross-tile merge NMS of one scene and the class-major compaction, three entry
// points around the same kernels (aidet_scene_*).
// Bound: the mask, FP32 issue (same pair arithmetic as riou.cu); everything else is latency.
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "geom.cuh"

#ifndef AIDET_NMS_COMPACT_PAIRS
#define AIDET_NMS_COMPACT_PAIRS 1          // sparse mask units: lanes walk their own circle-test candidates
#endif

namespace aidet {

// ------------------------------------------------------------------ box kinds
struct NmsRect {
  // rows (the box that is transformed) carry their per-box constants (48 B RectA), columns are Rect records that a lane
  // keeps in registers as RectB; intersection and areas are on the common HALF scale of geom.cuh: rect_inter_half
  using Row = RectA; using Col = Rect; using Reg = RectB;
  static constexpr int FMT = 5;
  __device__ static __forceinline__ void prepare(const float* p, float, Row* r, Col* c) { rect_prepare(p, r, c); }
  __device__ static __forceinline__ bool disjoint(const Row& a, const Col& b) {
    const float dx = a.cx - b.cx, dy = a.cy - b.cy, r = a.rad + b.rad;
    return fmaf(dx, dx, dy * dy) > r * r;
  }
  __device__ static __forceinline__ float inter(const Row& a, const Reg& b, float) { return rect_inter_half(a, b); }
  __device__ static __forceinline__ float area(const Row& a, float) { return a.harea; }
  __device__ static __forceinline__ float area_c(const Reg& b, float) { return b.harea; }
};
struct NmsQuad {
  using Row = QuadRow; using Col = QuadCol; using Reg = QuadReg;
  static constexpr int FMT = 8;
  __device__ static __forceinline__ void prepare(const float* p, float, Row* r, Col* c) { quad_prepare(p, r, c); }
  __device__ static __forceinline__ float overlap(const Row& a, const Col& b, float) { return quad_overlap(a, b, MODE_IOU); }
  __device__ static __forceinline__ bool disjoint(const Row& a, const Col& b) {
    const float dx = a.mx - b.mx, dy = a.my - b.my, r = a.rad + b.rad;
    return fmaf(dx, dx, dy * dy) > r * r;
  }
  __device__ static __forceinline__ float inter(const Row& a, const Reg& b, float) { return quad_inter(a, b); }
  __device__ static __forceinline__ float area(const Row& a, float) { return a.area; }
  __device__ static __forceinline__ float area_c(const Col& b, float) { return b.area; }
};
struct NmsHbb {
  using Row = HbbBox; using Col = HbbBox; using Reg = HbbBox;
  static constexpr int FMT = 4;
  __device__ static __forceinline__ void prepare(const float* p, float, Row* r, Col* c) {
    HbbBox b{p[0], p[1], p[2], p[3]};
    *r = b; *c = b;
  }
  __device__ static __forceinline__ float overlap(const Row& a, const Col& b, float one) {
    return hbb_overlap(a, b, one, MODE_IOU);
  }
  __device__ static __forceinline__ bool disjoint(const Row&, const Col&) { return false; }
  __device__ static __forceinline__ float inter(const Row& a, const Col& b, float one) {
    const float w = fmaxf(fminf(a.x2, b.x2) - fmaxf(a.x1, b.x1) + one, 0.f);
    const float h = fmaxf(fminf(a.y2, b.y2) - fmaxf(a.y1, b.y1) + one, 0.f);
    return w * h;
  }
  __device__ static __forceinline__ float area(const Row& a, float one) { return (a.x2 - a.x1 + one) * (a.y2 - a.y1 + one); }
  __device__ static __forceinline__ float area_c(const Col& b, float one) { return (b.x2 - b.x1 + one) * (b.y2 - b.y1 + one); }
};

// "IoU cmp thr" without the division: inter cmp thr * (area_a + area_b - inter).  The union is either 0
// (both boxes degenerate: IoU is defined 0 here, NaN > thr == false in the reference's HBB kernel) or far
// above the denormal range.  `zero_hit` = what the comparison gives for IoU == 0.
template <class O, bool GE>
__device__ __forceinline__ bool nms_hit(const typename O::Row& a, const typename O::Reg& b, float area_b, float one,
                                        float th, bool zero_hit) {
  if (O::disjoint(a, b)) return zero_hit;
  const float area_a = O::area(a, one);
  float inter = O::inter(a, b, one);
  if constexpr (O::FMT != 4) inter = fminf(fmaxf(inter, 0.0f), fminf(area_a, area_b));
  const float den = area_a + area_b - inter;
  const float rhs = th * den;
  const bool h = GE ? (inter >= rhs) : (inter > rhs);
  return den > 0.0f ? h : zero_hit;
}

// nms_hit for "IoU > thr" inside a DENSE block (the caller's probe found every lane's bounding circle meeting the block's
// first and last row): no per-pair circle test and no early-out branch, and the comparison as
//     inter (1 + thr) > thr area_a + thr area_b          (k1 = 1 + thr, thb = thr area_b: per column)
// -- 3 instructions after the clamp of `inter` to the smaller area (which keeps zero-area boxes from suppressing anything:
// inter <= 0 there).  A pair whose circles are disjoint after all yields inter = 0 or rounding noise ~1e-7 of the areas,
// far below thr (area_a + area_b) for the thresholds the caller admits (thr >= 1e-4).
template <class O>
__device__ __forceinline__ bool nms_hit_dense(const typename O::Row& a, const typename O::Reg& b, float area_b, float one,
                                              float th, float k1, float thb) {
  const float area_a = O::area(a, one);
  const float inter = fminf(O::inter(a, b, one), fminf(area_a, area_b));
  return inter * k1 > fmaf(th, area_a, thb);
}

// The 32 greedy decisions of one diagonal block: keep_k = !init_k && no kept row j < k has bit k set in its diagonal word
// D_j (upper triangular: D_j only holds bits > j).  The sequential form costs ~40 cycles per decision in dependent ALU /
// predicate latency (measured: 0.7 us per block whether the words come from shuffles or from shared memory).  Here the
// triangular system is solved by fixed-point iteration over all 32 rows at once: K <- ~(init | OR_{j in K} D_j), one
// redux.or per round, starting from "every undecided row is kept".  After t rounds the decisions of rows < t are final
// (row k depends on rows < k only), so it terminates within 33 rounds with the unique solution -- the greedy result --
// and in practice within a few: one round when nothing overlaps, two when the first row suppresses the rest.
// init: warp-uniform; dmine: the diagonal word of row `lane` (anything for rows that init marks).
__device__ __forceinline__ uint32_t greedy_block(uint32_t init, uint32_t dmine, int lane) {
  uint32_t K = ~init;
  for (;;) {
    const uint32_t R = init | __reduce_or_sync(0xffffffffu, ((K >> lane) & 1u) ? dmine : 0u);
    const uint32_t Kn = ~R;
    if (Kn == K) break;
    K = Kn;
  }
  return K;
}

// ------------------------------------------------------------------ 1. keys
__device__ __forceinline__ uint32_t orderable(float f) {      // ascending uint <=> ascending float
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256) nms_keys_kernel(const float* __restrict__ scores, const int* __restrict__ groups,
                                                       int n, uint64_t* keys, int* idx, uint8_t* flags, int* gbounds,
                                                       int n_groups) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n_groups) gbounds[i] = 0;          // [start | end) of every group, empty by default
  if (i >= n) return;
  uint32_t g = groups ? (uint32_t)groups[i] : 0u;
  g = min(g, (uint32_t)n_groups);                // ids outside [0, n_groups): behind every group, never scanned
  keys[i] = ((uint64_t)g << 32) | (uint64_t)(~orderable(scores[i]));
  idx[i] = i;
  flags[i] = 0;
}

// Small inputs (n <= kRankSortMax, every per-image config): the seven launches of a radix sort cost more than
// the sort itself, so ONE kernel does keys + sort + gather + group bounds by counting.  Keys are made unique
// -- (group, ~score, original index) packed into 64 bits -- so the sorted position of a box is simply
// rank = #{j : key_j < key_i}: the order of the stable radix sort (ties -> ascending original index).
// A warp ranks kRankBoxes boxes at once (every key it builds is compared against all of them) plus, for warp
// w <= n_groups, the first position of group w (= #{key_j < w << 45}): gstart[w], and gend[w-1].  The prepared
// record (geom.cuh) of every box goes straight to its sorted slot.
constexpr int kRankSortMax = 8192;          // 13 index bits
#ifndef AIDET_NMS_BUCKET_SORT
#define AIDET_NMS_BUCKET_SORT 1             // n > kRankSortMax: bucket + rank kernels (1) or key kernel + CUB radix sort + gather (0)
#endif
constexpr bool kBucketSort = AIDET_NMS_BUCKET_SORT != 0;
constexpr int kRankBoxes = 4;
constexpr int kRankGroupShift = 45;         // 13 index + 32 score bits below the group
constexpr int kRankMaxGroups = (1 << 19) - 2;

__device__ __forceinline__ uint64_t rank_key(const float* __restrict__ scores, const int* __restrict__ groups, int j,
                                             uint32_t n_groups) {
  uint32_t g = groups ? (uint32_t)__ldg(groups + j) : 0u;
  g = min(g, n_groups);                      // ids outside [0, n_groups) sort behind every group and are never scanned
  return ((uint64_t)g << kRankGroupShift) | ((uint64_t)(~orderable(__ldg(scores + j))) << 13) | (uint64_t)j;
}

constexpr int kRankWarps = 4;               // CTA = 4 warps = 16 boxes: ~2.5 CTAs per SM at config C2

template <class O>
__global__ void __launch_bounds__(kRankWarps * 32)
nms_rank_gather_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ groups,
                       int n, float one, typename O::Row* rows, typename O::Col* cols, int* __restrict__ order,
                       uint8_t* __restrict__ flags, int* gstart, int* gend, int n_groups, int* __restrict__ counters) {
  extern __shared__ __align__(16) uint64_t skeys[];          // all n keys: loaded once per CTA with every load in flight
  __shared__ int s_rank[kRankWarps * kRankBoxes];            // (walking global memory instead costs one L2 / DRAM latency
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;  //  per step: the warps of an SM advance in lock step)
  const int w = blockIdx.x * kRankWarps + wl;
  if (blockIdx.x == 0 && threadIdx.x < 2) counters[threadIdx.x] = 0;     // tile ticket, scan-done count
  for (int j = threadIdx.x; j < n; j += kRankWarps * 32) skeys[j] = rank_key(scores, groups, j, (uint32_t)n_groups);
  __syncthreads();
  const int i0 = w * kRankBoxes;
  uint64_t ki[kRankBoxes];
#pragma unroll
  for (int b = 0; b < kRankBoxes; ++b) ki[b] = (i0 + b < n) ? skeys[i0 + b] : 0ull;
  const uint64_t kg = (w <= n_groups) ? ((uint64_t)w << kRankGroupShift) : 0ull;
  int c[kRankBoxes] = {0, 0, 0, 0}, cg = 0;
  const int n_loop = (i0 < n || w <= n_groups) ? n : 0;     // idle warps of the last CTA only wait at the barrier
#pragma unroll 4
  for (int j = lane; j < n_loop; j += 32) {
    const uint64_t kj = skeys[j];
#pragma unroll
    for (int b = 0; b < kRankBoxes; ++b) c[b] += (kj < ki[b]) ? 1 : 0;
    cg += (kj < kg) ? 1 : 0;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
#pragma unroll
    for (int b = 0; b < kRankBoxes; ++b) c[b] += __shfl_xor_sync(0xffffffffu, c[b], d);
    cg += __shfl_xor_sync(0xffffffffu, cg, d);
  }
  if (lane == kRankBoxes && w <= n_groups) {
    if (w < n_groups) gstart[w] = cg;
    if (w > 0) gend[w - 1] = cg;
  }
  // the CTA's ranks go through shared memory so that consecutive lanes of ONE warp prepare the records
  // (rect_prepare evaluates sin/cos in double; a few active lanes in every warp would multiply the FP64 work)
  if (lane < kRankBoxes) {
    int rank = c[0];
#pragma unroll
    for (int b = 1; b < kRankBoxes; ++b) if (lane == b) rank = c[b];
    s_rank[wl * kRankBoxes + lane] = rank;
  }
  __syncthreads();
  if (threadIdx.x < kRankWarps * kRankBoxes) {
    const int i = blockIdx.x * kRankWarps * kRankBoxes + threadIdx.x;
    if (i < n) {
      const int rank = s_rank[threadIdx.x];
      order[rank] = i;
      flags[i] = 0;
      float bx[O::FMT];
#pragma unroll
      for (int k = 0; k < O::FMT; k++) bx[k] = boxes[(size_t)i * O::FMT + k];
      typename O::Row r; typename O::Col cc;
      O::prepare(bx, one, &r, &cc);
      rows[rank] = r;
      if constexpr (!std::is_same<typename O::Row, typename O::Col>::value) cols[rank] = cc;
    }
  }
}

// ------------------------------------------------------------------ 1b. bucket + rank (n > kRankSortMax)
// The order NMS needs -- group ascending, score descending, original index ascending -- without a radix sort: a 64-bit
// key sort of 16 k - 50 k elements is 4-5 latency-bound passes of ~10 us on three CTAs (59-69 us: a third of an 8-tile
// batch).  Boxes are instead dealt to their group's bucket (counting sort by group id: histogram, scan, scatter -- warp
// aggregated atomics, order inside a bucket arbitrary) and every box then ranks itself INSIDE its bucket by counting the
// smaller keys (score bits, index): sum over groups of n_g^2 comparisons at ~0.12 warp instructions each -- 5 % of what the
// mask kernel spends on the same dense group, ~60 % on a sparse one -- so the host keeps the radix sort for inputs whose
// groups average more than 24576 boxes.
__global__ void __launch_bounds__(256) nms_bucket_hist_kernel(const int* __restrict__ groups, int n, int n_groups,
                                                              unsigned* __restrict__ gcnt, uint8_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flags[i] = 0;
  const uint32_t g = groups ? min((uint32_t)__ldg(groups + i), (uint32_t)n_groups) : 0u;   // ids outside [0, n_groups): last bucket
  const unsigned m = __match_any_sync(__activemask(), g);
  if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(gcnt + g, (unsigned)__popc(m));
}

// exclusive scan of the n_groups + 1 bucket counts (one CTA): gbase[g] .. gbase[g + 1] = bucket g; cursors for the scatter;
// group bounds of the n_groups real groups for the mask / scan kernels
__global__ void __launch_bounds__(1024) nms_bucket_scan_kernel(unsigned* __restrict__ gcnt /* in: counts, out: gbase */,
                                                               unsigned* __restrict__ gcursor, int n_groups,
                                                               int* __restrict__ gstart, int* __restrict__ gend) {
  __shared__ unsigned warp_sum[32];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base <= n_groups; base += 1024) {
    const int g = base + threadIdx.x;
    const unsigned v = (g <= n_groups) ? gcnt[g] : 0u;
    unsigned x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += y; }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const unsigned incl = x + (warp ? warp_sum[warp - 1] : 0u) + carry;
    if (g <= n_groups) {
      const unsigned excl = incl - v;
      gcnt[g] = excl; gcursor[g] = excl;
      if (g < n_groups) { gstart[g] = (int)excl; gend[g] = (int)incl; }
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) gcnt[n_groups + 1] = carry;           // = n: end of the last bucket
}

__global__ void __launch_bounds__(256) nms_bucket_scatter_kernel(const float* __restrict__ scores, const int* __restrict__ groups,
                                                                 int n, int n_groups, unsigned* __restrict__ gcursor,
                                                                 uint64_t* __restrict__ bkeys, int* __restrict__ bgrp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int lane = threadIdx.x & 31;
  const uint32_t g = groups ? min((uint32_t)__ldg(groups + i), (uint32_t)n_groups) : 0u;
  const unsigned m = __match_any_sync(__activemask(), g);
  const int leader = __ffs(m) - 1;
  unsigned base = 0;
  if (lane == leader) base = atomicAdd(gcursor + g, (unsigned)__popc(m));
  base = __shfl_sync(m, base, leader);
  const unsigned pos = base + (unsigned)__popc(m & ((1u << lane) - 1u));
  bkeys[pos] = ((uint64_t)(~orderable(__ldg(scores + i))) << 32) | (uint64_t)(uint32_t)i;   // better boxes: smaller keys
  bgrp[pos] = (int)g;
}

// A CTA owns 64 consecutive bucket positions (8 per warp).  The keys of the buckets those positions lie in -- one contiguous
// range -- pass through shared memory in tiles; a box counts the keys of ITS bucket that are smaller than its own.  The
// records are prepared by the first two warps (consecutive lanes: rect_prepare evaluates sin / cos in double).
constexpr int kBucketTile = 2048;
constexpr int kBucketPer = 8;                      // boxes per warp
constexpr int kBucketCta = 8 * kBucketPer;         // positions per CTA
template <class O>
__global__ void __launch_bounds__(256) nms_bucket_rank_kernel(const float* __restrict__ boxes, const uint64_t* __restrict__ bkeys,
                                                              const int* __restrict__ bgrp, const unsigned* __restrict__ gbase,
                                                              int n, int n_groups, float one, typename O::Row* rows,
                                                              typename O::Col* cols, int* __restrict__ order) {
  __shared__ uint64_t tkeys[kBucketTile];
  __shared__ int tgrp[kBucketTile];
  __shared__ int s_pos[kBucketCta], s_idx[kBucketCta];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p_first = blockIdx.x * kBucketCta, p_last = min(n, p_first + kBucketCta) - 1;
  // the last bucket (ids outside [0, n_groups): never scanned, never kept) needs no order: its positions are skipped, and
  // the CTA's key range ends at the last real bucket
  const int g_first = __ldg(bgrp + p_first), g_last = min(__ldg(bgrp + p_last), n_groups - 1);
  if (g_first >= n_groups) return;
  const unsigned lo = gbase[g_first], hi = gbase[g_last + 1];                                // union of the CTA's buckets
  const bool cta_one = g_first == g_last;          // every position of the CTA in one bucket (big buckets): no group ids needed
  uint64_t ki[kBucketPer]; int gi[kBucketPer]; int cnt[kBucketPer];
#pragma unroll
  for (int b = 0; b < kBucketPer; ++b) {
    const int p = p_first + warp * kBucketPer + b;
    ki[b] = (p < n) ? __ldg(bkeys + p) : 0ull;
    gi[b] = (p < n) ? __ldg(bgrp + p) : -1;
    if (gi[b] >= n_groups) gi[b] = -1;
    cnt[b] = 0;
  }
  bool warp_one = gi[0] >= 0;                      // the warp's boxes share a bucket: clip the tile to it once
#pragma unroll
  for (int b = 1; b < kBucketPer; ++b) warp_one = warp_one && gi[b] == gi[0];
  for (unsigned t0 = lo; t0 < hi; t0 += kBucketTile) {
    const int tn = (int)min((unsigned)kBucketTile, hi - t0);
    __syncthreads();
    for (int j = tid; j < tn; j += 256) { tkeys[j] = __ldg(bkeys + t0 + j); if (!cta_one) tgrp[j] = __ldg(bgrp + t0 + j); }
    __syncthreads();
    if (warp_one) {
      const int j0 = max(0, (int)((long long)gbase[gi[0]] - (long long)t0)), j1 = min(tn, (int)((long long)gbase[gi[0] + 1] - (long long)t0));
#pragma unroll 2
      for (int j = j0 + lane; j < j1; j += 32) {
        const uint64_t kj = tkeys[j];
#pragma unroll
        for (int b = 0; b < kBucketPer; ++b) cnt[b] += (kj < ki[b]) ? 1 : 0;
      }
    } else {
      for (int j = lane; j < tn; j += 32) {
        const uint64_t kj = tkeys[j]; const int gj = cta_one ? g_first : tgrp[j];
#pragma unroll
        for (int b = 0; b < kBucketPer; ++b) cnt[b] += (gj == gi[b] && kj < ki[b]) ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int b = 0; b < kBucketPer; ++b) cnt[b] = __reduce_add_sync(0xffffffffu, cnt[b]);
  if (lane < kBucketPer) {
    int c = cnt[0], g = gi[0]; uint64_t k = ki[0];
#pragma unroll
    for (int b = 1; b < kBucketPer; ++b) if (lane == b) { c = cnt[b]; g = gi[b]; k = ki[b]; }
    s_pos[warp * kBucketPer + lane] = (g >= 0) ? (int)gbase[g] + c : -1;
    s_idx[warp * kBucketPer + lane] = (int)(uint32_t)k;
  }
  __syncthreads();
  if (tid < kBucketCta && s_pos[tid] >= 0) {
    const int pos = s_pos[tid], i = s_idx[tid];
    order[pos] = i;
    float bx[O::FMT];
#pragma unroll
    for (int k = 0; k < O::FMT; k++) bx[k] = __ldg(boxes + (size_t)i * O::FMT + k);
    typename O::Row r; typename O::Col cc;
    O::prepare(bx, one, &r, &cc);
    rows[pos] = r;
    if constexpr (!std::is_same<typename O::Row, typename O::Col>::value) cols[pos] = cc;
  }
}

// ------------------------------------------------------------------ 2. gather
template <class O>
__global__ void __launch_bounds__(256) nms_gather_kernel(const float* __restrict__ boxes, const uint64_t* __restrict__ keys,
                                                         const int* __restrict__ order, int n, float one,
                                                         typename O::Row* rows, typename O::Col* cols, int* gstart,
                                                         int* gend, int n_groups) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int i = order[p];
  float b[O::FMT];
#pragma unroll
  for (int k = 0; k < O::FMT; k++) b[k] = boxes[(size_t)i * O::FMT + k];
  typename O::Row r; typename O::Col c;
  O::prepare(b, one, &r, &c);
  rows[p] = r;
  if constexpr (!std::is_same<typename O::Row, typename O::Col>::value) cols[p] = c;
  int g = (int)(keys[p] >> 32);
  if (g >= 0 && g < n_groups) {
    if (p == 0 || (int)(keys[p - 1] >> 32) != g) gstart[g] = p;
    if (p == n - 1 || (int)(keys[p + 1] >> 32) != g) gend[g] = p + 1;
  }
}

// units of a group with ng boxes: strips of 32 columns c = 0..ns-1, strip c needs rows 0 .. min(ng, 32c+32)-1 (every row
// that precedes one of its columns, plus the rest of the diagonal block so that all of its words get written) in
// chunks of RC rows: (c+1) * 32/RC units, the last strip ceil(ng / RC)
__host__ __device__ __forceinline__ int fused_units(int ng, int q /* 32 / RC */, int rc) {
  if (ng <= 0) return 0;
  const int ns = (ng + 31) >> 5;
  return q * ((ns - 1) * ns / 2) + (ng + rc - 1) / rc;
}

// tiles of tile_rows x 256 cols over the bounding rectangle of every group (tiles under the
// diagonal are skipped by the mask kernel); prefix[g] = first tile id of group g.
// unit_rc > 0: counts the warp-level mask units (fused_units with unit_rc rows per unit) instead of tiles.
__global__ void __launch_bounds__(1024) nms_tile_prefix_kernel(const int* __restrict__ gstart, const int* __restrict__ gend,
                                                               int n_groups, int tile_rows, int* prefix, int unit_rc) {
  __shared__ int warp_sum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_groups; base += 1024) {
    int g = base + threadIdx.x;
    int v = 0;
    if (g < n_groups) {
      int ng = gend[g] - gstart[g];
      v = unit_rc ? fused_units(ng, 32 / unit_rc, unit_rc) : ((ng + tile_rows - 1) / tile_rows) * ((ng + 255) / 256);
    }
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= d) x += y; }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sum[threadIdx.x];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, d); if (threadIdx.x >= d) w += y; }
      warp_sum[threadIdx.x] = w;
    }
    __syncthreads();
    int incl = x + ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0) + carry;
    if (g < n_groups) prefix[g] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) { prefix[n_groups] = carry; prefix[n_groups + 1] = 0; prefix[n_groups + 2] = 0; }   // + tile ticket, scan-done count
}

// ------------------------------------------------------------------ 3. mask
constexpr int kTileRows = 64;      // upper bound; small problems use 32/16/8-row tiles so every SM gets several CTAs
constexpr int kTileCols = 256;
constexpr int kLocalGroups = 255;   // the mask kernel scans the tile counts of up to this many groups itself

struct NmsTile { int start, ng, r0, cq0, nr, g; };   // nr == 0: no tile left

// Tiles are handed out by a ticket counter (balanced although diagonal tiles are cheaper and groups differ in
// size).  Thread 0 is the scheduler: while the CTA works on tile k it draws the ticket of tile k+1, skips tiles
// left of the diagonal, and starts the TMA copy of its row records into the other staging buffer.
// 4 resident CTAs per SM (<= 64 registers).  A dual-row step like the overlap-matrix kernel's was measured here and
// is no faster (one dense 16384-box group: r1 0.780 vs 0.763 ms, r2 with the saturation arithmetic 0.897 vs 0.893 ms
// per call): the ballot per row already breaks the chains.
#ifndef AIDET_NMS_MINB
#define AIDET_NMS_MINB 4
#endif
template <class O, bool GE>
__global__ void __launch_bounds__(kTileCols, O::FMT == 8 ? 3 : AIDET_NMS_MINB)
nms_mask_kernel(const typename O::Row* __restrict__ rows, const typename O::Col* __restrict__ cols,
                const int* __restrict__ gstart, const int* __restrict__ gend, const int* __restrict__ prefix,
                int n_groups, const float* __restrict__ thr, int n_thr, float one, int tile_rows,
                uint32_t* __restrict__ mask32, long long pitch32, int* __restrict__ ticket, int local_prefix) {
  using Row = typename O::Row; using Col = typename O::Col;
  __shared__ __align__(128) Row stage[2][kTileRows];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ NmsTile desc[2];
  __shared__ int sprefix[kLocalGroups + 1];
  __shared__ int swarp[8];
  if (local_prefix) {            // n_groups <= kLocalGroups: every CTA scans the tile counts itself (saves a launch)
    const int g = threadIdx.x;
    int v = 0;
    if (g < n_groups) { int ng = gend[g] - gstart[g]; v = ((ng + tile_rows - 1) / tile_rows) * ((ng + kTileCols - 1) / kTileCols); }
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= d) x += y; }
    if ((threadIdx.x & 31) == 31) swarp[threadIdx.x >> 5] = x;
    __syncthreads();
    int carry = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) carry += swarp[w];
    if (g <= n_groups) sprefix[g] = x + carry - v;            // exclusive; entry n_groups = total (v = 0 there)
    __syncthreads();
    prefix = sprefix;
  }
  const int total = prefix[n_groups];
  auto schedule = [&](int buf) {                       // thread 0 only
    for (;;) {
      const int t = atomicAdd(ticket, 1);
      if (t >= total) { desc[buf].nr = 0; return; }
      int lo = 0, hi = n_groups;                       // last g with prefix[g] <= t
      while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (prefix[mid] <= t) lo = mid; else hi = mid; }
      const int start = gstart[lo], ng = gend[lo] - start;
      const int ncq = (ng + kTileCols - 1) / kTileCols;
      const int local = t - prefix[lo];
      const int r0 = (local / ncq) * tile_rows, cq0 = (local % ncq) * kTileCols;
      if (cq0 + kTileCols <= r0) continue;             // tile entirely left of the diagonal
      const int nr = min(tile_rows, ng - r0);
      desc[buf] = NmsTile{start, ng, r0, cq0, nr, lo};
      const uint32_t bytes = (uint32_t)(nr * (int)sizeof(Row));
      mbar_expect_tx(&bar[buf], bytes);
      tma_load_1d(&stage[buf][0], rows + start + r0, bytes, &bar[buf]);
      return;
    }
  };
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); schedule(0); }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int it = 0;; ++it) {
    const int buf = it & 1;
    const NmsTile d = desc[buf];
    if (d.nr == 0) break;
    if (threadIdx.x == 0) schedule(buf ^ 1);           // stage[buf^1] / desc[buf^1] were released by the last barrier
    const float th = thr[n_thr == 1 ? 0 : d.g];
    const bool zero_hit = GE ? (0.0f >= th) : (0.0f > th);
    const int c0 = d.cq0 + warp * 32;
    const bool need = (c0 + 31 > d.r0) && (c0 < d.ng); // some column of the strip follows some row of the tile (warp-uniform)
    const int j = c0 + lane;
    const bool live = j < d.ng;
    typename O::Reg me;
    float area_me = 0.f;
    if (need) { me = cols[d.start + (live ? j : d.ng - 1)]; area_me = O::area_c(me, one); }
    mbar_wait(&bar[buf], (it >> 1) & 1);
    if (need) {
      const Row* st = &stage[buf][0];
      const bool above = d.r0 + d.nr - 1 < c0;         // every row of the tile precedes every column of the strip
      for (int rb = 0; rb < d.nr; rb += 32) {          // 32 rows -> one half-word per lane
        const int lim = min(32, d.nr - rb);
        uint32_t word = 0;
        if (above) {                                   // the common case: no per-row diagonal tests
          bool dense = false;
          if constexpr (!GE && O::FMT != 4)
            dense = th >= 1e-4f && __all_sync(0xffffffffu, live && !O::disjoint(st[rb], me) && !O::disjoint(st[rb + lim - 1], me));
          if (dense) {
            if constexpr (!GE && O::FMT != 4) {
              const float k1 = 1.0f + th, thb = th * area_me;
#pragma unroll 2
              for (int k = 0; k < lim; ++k) {
                const uint32_t b = __ballot_sync(0xffffffffu, nms_hit_dense<O>(st[rb + k], me, area_me, one, th, k1, thb));
                if (lane == k) word = b;
              }
            }
          } else {
#pragma unroll 2
            for (int k = 0; k < lim; ++k) {
              const bool hit = nms_hit<O, GE>(st[rb + k], me, area_me, one, th, zero_hit);
              const uint32_t b = __ballot_sync(0xffffffffu, hit && live);
              if (lane == k) word = b;
            }
          }
        } else {
          for (int k = 0; k < lim; ++k) {
            const int i = d.r0 + rb + k;
            uint32_t b = 0;
            if (i < c0 + 31) {                         // otherwise no column of this strip follows row i
              const bool hit = nms_hit<O, GE>(st[rb + k], me, area_me, one, th, zero_hit);
              b = __ballot_sync(0xffffffffu, hit && live && j > i);
            }
            if (lane == k) word = b;
          }
        }
        if (lane < lim) mask32[(long long)(d.start + d.r0 + rb + lane) * pitch32 + (c0 >> 5)] = word;
      }
    }
    __syncthreads();                                   // stage[buf] and desc[buf] are free again
  }
}

// Many small groups (batched tiles, a scene's tile x class groups: a few hundred boxes each): 256-column tiles leave most
// warps of a CTA without a strip above the diagonal (ncu on 120 groups of ~394 boxes: ~55 % of the warp slots idle at the
// tile barrier, mask at 40 % of the roofline).  Here the work is the fused kernel's warp-level unit -- 32 columns (one
// record per lane, in registers) x rc rows (warp-uniform loads), only units at or above the diagonal -- dealt round-robin
// to all warps; no staging, no CTA barrier.  Row-major mask layout of the tile kernel (the scan kernel reads it).
template <class O, bool GE>
__global__ void __launch_bounds__(256, O::FMT == 8 ? 3 : 4)
nms_mask_units_kernel(const typename O::Row* __restrict__ rows, const typename O::Col* __restrict__ cols,
                      const int* __restrict__ gstart, const int* __restrict__ gend, const int* __restrict__ prefix,
                      int n_groups, const float* __restrict__ thr, int n_thr, float one, int rc_rows,
                      uint32_t* __restrict__ mask32, long long pitch32, int stage_prefix) {
  using Row = typename O::Row;
  extern __shared__ int sprefix_u[];                                   // [n_groups + 1] when stage_prefix
  __shared__ __align__(16) Row wrows[8][32];                           // a warp's unit rows: ONE round trip per unit instead of one per row
  __shared__ uint32_t s_ww[8][32];                                     // a warp's row words while its lanes walk their own candidates
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (stage_prefix) {
    for (int g = tid; g <= n_groups; g += 256) sprefix_u[g] = __ldg(prefix + g);
    __syncthreads();
    prefix = sprefix_u;
  }
  const int q = 32 / rc_rows;
  const int total = prefix[n_groups];
  const int u_step = gridDim.x * 8;
  int cur_g = -1, cur_c = -1, start = 0, ng = 0;
  typename O::Reg me; float area_me = 0.f, th = 0.f; bool zero_hit = false;
  for (int u = blockIdx.x * 8 + warp; u < total; u += u_step) {
    int lo = 0, hi = n_groups;                                         // last g with prefix[g] <= u
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= u) lo = mid; else hi = mid; }
    const int g = lo;
    if (g != cur_g) { start = __ldg(gstart + g); ng = __ldg(gend + g) - start; }
    const int ns = (ng + 31) >> 5;
    const int l = u - prefix[g];
    int c = (int)((sqrtf(8.0f * (float)l / (float)q + 1.0f) - 1.0f) * 0.5f);   // strip: the largest c with q c (c+1) / 2 <= l
    c = min(c, ns - 1);
    while (c > 0 && q * (c * (c + 1) / 2) > l) --c;
    while (c < ns - 1 && q * ((c + 1) * (c + 2) / 2) <= l) ++c;
    const int r0 = (l - q * (c * (c + 1) / 2)) * rc_rows;
    const int c0 = c << 5;
    const int r_end = min(min(ng, c0 + 32), r0 + rc_rows);
    const int j = c0 + lane;
    const bool live = j < ng;
    if (g != cur_g || c != cur_c) {
      me = cols[start + (live ? j : ng - 1)];
      area_me = O::area_c(me, one);
      th = __ldg(thr + (n_thr == 1 ? 0 : g));
      zero_hit = GE ? (0.0f >= th) : (0.0f > th);
      cur_g = g; cur_c = c;
    }
    {                                                                  // the unit's rows are contiguous records: copy them as 16-byte pieces
      constexpr int RQ = (int)sizeof(Row) / 16;
      const float4* src = reinterpret_cast<const float4*>(rows + start + r0);
      float4* dst = reinterpret_cast<float4*>(&wrows[warp][0]);
      __syncwarp();                                                    // the previous unit's reads are done
      for (int i = lane; i < (r_end - r0) * RQ; i += 32) dst[i] = __ldg(src + i);
      __syncwarp();
    }
    const Row* rr = &wrows[warp][0] - r0;                              // rr[i] = row i of the group, r0 <= i < r_end
    uint32_t word = 0;
    bool dense = false;
    if constexpr (!GE && O::FMT != 4)                                  // unit above the diagonal, every circle meets its first and last row
      dense = r_end <= c0 && r_end > r0 && th >= 1e-4f &&
              __all_sync(0xffffffffu, live && !O::disjoint(rr[r0], me) && !O::disjoint(rr[r_end - 1], me));
    if (dense) {
      if constexpr (!GE && O::FMT != 4) {
        const float k1 = 1.0f + th, thb = th * area_me;
#pragma unroll 2
        for (int i = r0; i < r_end; ++i) {
          const uint32_t bb = __ballot_sync(0xffffffffu, nms_hit_dense<O>(rr[i], me, area_me, one, th, k1, thb));
          if (lane == (i & 31)) word = bb;
        }
      }
    } else if (O::FMT != 4 && !zero_hit && AIDET_NMS_COMPACT_PAIRS) {
      // sparse unit: circle tests for all rows, then every lane walks its own candidates (see nms_fused_kernel)
      uint32_t cand = 0;
      for (int i = r0; i < r_end; ++i)
        if (live && j > i && !O::disjoint(rr[i], me)) cand |= 1u << (i - r0);
      uint32_t* ww = &s_ww[warp][0];
      ww[lane] = 0u;
      __syncwarp();
      while (cand) {
        const int k = __ffs(cand) - 1;
        cand &= cand - 1;
        const Row& a = rr[r0 + k];
        const float area_a = O::area(a, one);
        float inter = O::inter(a, me, one);
        inter = fminf(fmaxf(inter, 0.0f), fminf(area_a, area_me));
        const float den = area_a + area_me - inter;
        const float rhs = th * den;
        const bool h = GE ? (inter >= rhs) : (inter > rhs);
        if (h && den > 0.0f) atomicOr(&ww[k], 1u << lane);
      }
      __syncwarp();
      const int k_mine = (r0 & ~31) + lane - r0;
      if (k_mine >= 0 && k_mine < r_end - r0) word = ww[k_mine];
      __syncwarp();
    } else {
#pragma unroll 2
      for (int i = r0; i < r_end; ++i) {
        const Row a = rr[i];                                           // warp-uniform address: one transaction
        const bool hit = nms_hit<O, GE>(a, me, area_me, one, th, zero_hit);
        const uint32_t bb = __ballot_sync(0xffffffffu, hit && live && j > i);
        if (lane == (i & 31)) word = bb;
      }
    }
    // rows r0 .. r_end-1 lie in one 32-row block (rc_rows divides 32): lane (i & 31) holds row i's word
    const int i_mine = (r0 & ~31) + lane;
    if (i_mine >= r0 && i_mine < r_end) mask32[(long long)(start + i_mine) * pitch32 + c] = word;
  }
}

// ------------------------------------------------------------------ 4. scan
// One CTA (256 threads) per group walks the group's rows in score order, 32 rows (one diagonal
// half-word) per step, entirely on the device -- the reference copies the whole mask to the host
// and scans it there (nms_kernel.cu:105-131).  `removed` (one bit per sorted position) lives in
// shared memory.  The mask rows of step b+1 are prefetched with cp.async (LDGSTS) into a
// double-buffered panel (32 rows x kScanPW half-words, up to the next 16384 columns) while step b runs:
//   warp 0 : greedy chain over the 32 diagonal bits (all lanes redundantly, no divergence)
//   all    : OR the rows of the kept boxes into `removed` -- from the panel, and straight from
//            global memory for columns beyond the panel (groups of more than 16384 boxes only)
// Steps whose 32 rows are all suppressed already (common once a few strong boxes are kept) skip the
// chain, the OR and the prefetch.
constexpr int kScanPWMax = 512;    // panel width in half-words, chosen per launch: min(512, ceil(n / 32))

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256) nms_scan_kernel(const uint32_t* __restrict__ mask32, long long pitch32,
                                                       const int* __restrict__ gstart, const int* __restrict__ gend,
                                                       const int* __restrict__ order, uint8_t* __restrict__ flags,
                                                       int removed_cap, int kScanPW, int n,
                                                       long long* __restrict__ keep_out, int* __restrict__ n_keep,
                                                       int* __restrict__ done) {
  extern __shared__ uint32_t sm[];
  uint32_t* removed = sm;                                   // [removed_cap]
  uint32_t* panel = sm + removed_cap;                       // [2][32][kScanPW + 4]
  __shared__ uint32_t keep_word;
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int start = gstart[g], ng = gend[g] - start;
  const int nhw = (ng + 31) >> 5;
  if (ng > 0) {
  for (int h = tid; h < nhw; h += 256) removed[h] = 0;

  auto dead = [&](int b) {                                 // all 32 rows of step b suppressed (bits are only ever added)
    uint32_t cur = removed[b];
    if (b == nhw - 1 && (ng & 31)) cur |= ~0u << (ng & 31);
    return cur == ~0u;
  };
  // Panel rows hold the words from b & ~3 on (mask rows are 16-byte aligned: pitch32 is a multiple of 4), copied 16 bytes
  // at a time -- a quarter of the cp.async instructions of a word-by-word copy; word b + w of row r is at
  // [r * kPanelStride + (b & 3) + w].  The up to three words past the group's last one lie inside the row's pitch.
  const int kPanelStride = kScanPW + 4;
  auto fetch = [&](int b, int buf) {
    const int pwn = min(kScanPW, nhw - b);
    const int b4 = b & ~3, nq = (pwn + (b & 3) + 3) >> 2;     // aligned first word, quads per row
    uint32_t* dst = panel + buf * (32 * kPanelStride);
    for (int idx = tid; idx < 32 * nq; idx += 256) {
      const int r = idx / nq, q = idx - r * nq;
      const int row = b * 32 + r;
      uint32_t* d4 = dst + r * kPanelStride + 4 * q;
      if (row < ng) cp_async16(d4, mask32 + (long long)(start + row) * pitch32 + b4 + 4 * q);
      else *reinterpret_cast<uint4*>(d4) = make_uint4(0u, 0u, 0u, 0u);
    }
    cp_async_commit();
  };

  __shared__ int s_next;
  __syncthreads();
  fetch(0, 0);
  // b is always a LIVE block (some row not suppressed yet).  Runs of dead blocks -- the common case once a few strong
  // boxes are kept: one dense 16384-box group keeps 4 boxes and leaves ~500 of its 512 blocks dead -- are jumped over by
  // ONE parallel search of removed[] instead of one barrier round per block (r1: 108 us of a 0.9 ms call).
  int b = 0, buf = 0;
  while (b < nhw) {
    // removed[] is final for block b here (trailing barrier of the previous step)
    const int nb = b + 1;
    bool pre = false;
    if (nb < nhw && !dead(nb)) { fetch(nb, buf ^ 1); pre = true; cp_async_wait<1>(); }   // speculative: b may still kill nb
    else cp_async_wait<0>();
    __syncthreads();                                        // panel[buf] landed
    const uint32_t* pan = panel + buf * (32 * kPanelStride) + (b & 3);     // word b + w of row r: pan[r * kPanelStride + w]
    if (tid < 32) {
      const uint32_t d = pan[lane * kPanelStride];          // diagonal half-word of row 32b + lane (0 past the end)
      uint32_t cur = removed[b];
      if (b == nhw - 1 && (ng & 31)) cur |= ~0u << (ng & 31);   // positions past the group end
      const uint32_t keep = greedy_block(cur, d, lane);
      if (lane == 0) { keep_word = keep; s_next = nhw; }
    }
    __syncthreads();
    const uint32_t keep = keep_word;
    if (tid == 0) removed[b] = keep;                        // slot b is not read again: it now holds the keep bits
    const int pwn = min(kScanPW, nhw - b);
    // OR the kept rows into the later half-words: 8 threads per half-word, 4 rows each, merged with atomicOr
    for (int idx = tid; idx < (pwn - 1) * 8; idx += 256) {
      const int w = 1 + (idx >> 3), part = idx & 7;
      uint32_t acc = 0, todo = (keep >> (part * 4)) & 0xfu;
      while (todo) { const int k = part * 4 + __ffs(todo) - 1; todo &= todo - 1; acc |= pan[k * kPanelStride + w]; }
      if (acc) atomicOr(removed + b + w, acc);
    }
    for (int h = b + kScanPW + tid; h < nhw; h += 256) {    // beyond the panel: only for groups > 16384 boxes
      uint32_t acc = 0, todo = keep;
      while (todo) {
        const int k = __ffs(todo) - 1; todo &= todo - 1;
        acc |= mask32[(long long)(start + b * 32 + k) * pitch32 + h];
      }
      removed[h] |= acc;
    }
    __syncthreads();                                        // removed[] is final for the blocks after b
    if (pre && !dead(nb)) { b = nb; buf ^= 1; continue; }   // CTA-uniform: the prefetched block is the next live one
    // jump: first live block at or after nb (all threads search 256 words per round)
    for (int base = nb; base < nhw; base += 256) {
      const int h = base + tid;
      const unsigned m = __ballot_sync(0xffffffffu, h < nhw && !dead(h));
      if (m && lane == 0) atomicMin(&s_next, base + (tid & ~31) + __ffs(m) - 1);
      __syncthreads();
      if (s_next < nhw) break;
    }
    const int b2 = s_next;                                  // nhw: no live block left
    for (int h = nb + tid; h < b2; h += 256) removed[h] = 0u;   // jumped-over blocks keep nothing
    if (b2 >= nhw) break;
    cp_async_wait<0>();                                     // a wasted prefetch may still be landing in panel[buf ^ 1]
    __syncthreads();                                        // (also orders the reads of s_next before its next reset)
    buf ^= 1;
    fetch(b2, buf);
    b = b2;
  }
  cp_async_wait<0>();
  __syncthreads();
  // removed[b] now holds the keep bits of block b: mark the kept boxes (original indices) in parallel
  for (int i = tid; i < ng; i += 256)
    if ((removed[i >> 5] >> (i & 31)) & 1u) flags[order[start + i]] = 1;
  }
  // 5. compaction, by whichever CTA finishes last (saves a launch): kept flags -> ascending original indices
  //    (nms_kernel.cu:135-138 semantics).  `done` was zeroed by an earlier kernel of the same call.
  if (!done) return;                                        // large inputs: a separate compaction launch follows
  __shared__ int s_last;
  __shared__ int warp_sum[8];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(done, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // each thread owns a run of `per` (multiple of 16) flags, read 16 at a time through L2 (other CTAs wrote them)
  const int per = ((n + 255) / 256 + 15) & ~15;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  const uint4* f4 = reinterpret_cast<const uint4*>(flags);          // workspace slot is 128 B aligned and padded
  int cnt = 0;
  for (int i = lo; i < hi; i += 16) {
    const uint4 v = __ldcg(f4 + (i >> 4));
    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 16; ++q) if (i + q < hi) cnt += (wds[q >> 2] >> (8 * (q & 3))) & 1u;
  }
  int x = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  if (lane == 31) warp_sum[tid >> 5] = x;
  __syncthreads();
  int off = x - cnt;
  for (int w = 0; w < (tid >> 5); ++w) off += warp_sum[w];
  for (int i = lo; i < hi; i += 16) {
    const uint4 v = __ldcg(f4 + (i >> 4));
    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (i + q < hi && ((wds[q >> 2] >> (8 * (q & 3))) & 1u)) keep_out[off++] = i + q;
  }
  if (tid == 255) *n_keep = off;
}

// ------------------------------------------------------------------ 5. compact (n > kFusedCompactMax)
constexpr int kFusedCompactMax = 16384;

__global__ void __launch_bounds__(1024) nms_compact_kernel(const uint8_t* __restrict__ flags, int n,
                                                           long long* __restrict__ keep_out, int* __restrict__ n_keep) {
  __shared__ int warp_sum[32];
  // each thread owns a run of `per` (multiple of 16) flags, read 16 at a time (the workspace slot is 128 B aligned and padded)
  const int per = ((n + 1023) / 1024 + 15) & ~15;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  const uint4* f4 = reinterpret_cast<const uint4*>(flags);
  int cnt = 0;
  for (int i = lo; i < hi; i += 16) {
    const uint4 v = __ldg(f4 + (i >> 4));
    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 16; ++q) if (i + q < hi) cnt += (wds[q >> 2] >> (8 * (q & 3))) & 1u;
  }
  int x = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if ((threadIdx.x & 31) >= d) x += y; }
  if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = warp_sum[threadIdx.x];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, w, d); if (threadIdx.x >= d) w += y; }
    warp_sum[threadIdx.x] = w;
  }
  __syncthreads();
  int off = x - cnt + ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0);
  for (int i = lo; i < hi; i += 16) {
    const uint4 v = __ldg(f4 + (i >> 4));
    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (i + q < hi && ((wds[q >> 2] >> (8 * (q & 3))) & 1u)) keep_out[off++] = i + q;
  }
  if (threadIdx.x == 1023) *n_keep = off;
}

// ------------------------------------------------------------------ fused path (n <= 8192)
// Per-image problems (config C1: 2000 boxes in one group; C2: ~5800 boxes in 15 classes) hold a few microseconds of
// arithmetic, so three dependent launches with their ramp-up, per-CTA prologues and ticket atomics cost far more
// than the work (r1: C2 52 us, C1 138 us under graph replay).  ONE cooperative launch runs all phases, separated
// by two grid barriers:
//   rank   the CTAs that own boxes bucket all keys by group in shared memory (histogram + cursor atomics, warp
//          aggregated), so a box is ranked against its own group only (rank = sorted position: group, score
//          descending, index ascending); lanes 0-7 of a warp prepare the records of its 8 boxes and store them to
//          their sorted slots
//   mask   warp-level units: 32 columns (one record per lane, in registers) x RC rows (records read with
//          warp-uniform loads), only units at or above the diagonal are enumerated; a CTA takes a contiguous run
//          of units (same group, same strip: rows hit L1, the column record is reused) and its warps draw from it
//          through a shared-memory counter -- no staging, no CTA barrier, no global ticket.  The words of a 32-row
//          block are stored [word][row]: a coalesced 128-byte store per unit, one contiguous run per block
//   scan   one CTA per group, warp-specialised: warp 0 runs the greedy chain on the 32x32 diagonal block and takes
//          the contribution of a block to the NEXT three words itself (predicated ORs inside the chain's empty issue
//          slots); warp 1 streams the following blocks into a shared-memory ring, ONE TMA bulk copy (cp.async.bulk
//          + mbarrier) per block; warps 2-7 OR the kept rows into the words four and more blocks ahead (one redux.or
//          per word, each word owned by a fixed warp: no atomics).  The chain never meets a CTA barrier: a step costs
//          its 32 dependent bit decisions (taken two at a time), and blocks whose rows are all suppressed already
//          cost neither a chain nor a wait for their data
//   compact  by the CTA that finishes last.
namespace cg = cooperative_groups;

constexpr int kFusedMaxBoxes = 8192;
constexpr int kFusedMaxGroups = 1024;
#ifndef AIDET_NMS_PRODUCER_SLEEP
#define AIDET_NMS_PRODUCER_SLEEP __nanosleep(64)
#endif
#ifndef AIDET_NMS_FUSED_THREADS
#define AIDET_NMS_FUSED_THREADS 512
#endif
constexpr int kFusedThreads = AIDET_NMS_FUSED_THREADS;
constexpr int kFusedUnitBoxes = 8;         // boxes ranked + prepared per warp unit
constexpr int kScanSlotsMax = 32;          // mbarrier slots (blocks in flight <= ring capacity / block size <= 32)
constexpr int kScanHelpers = 6;              // helper warps 2 .. 7 of the scan (further warps of a larger CTA sit the scan out)
static_assert(kFusedThreads / 32 >= 2 + kScanHelpers, "the fused kernel needs 8 warps");
constexpr int kSoloBlocks = 16;           // groups of <= 16 blocks (512 boxes): scanned by one warp from shared memory
#ifndef AIDET_NMS_UNITS_PER_WARP_LARGE
#define AIDET_NMS_UNITS_PER_WARP_LARGE 4   // nms_mask_units_kernel: the same rule for the batched path
#endif
#ifndef AIDET_NMS_UNITS_PER_WARP
#define AIDET_NMS_UNITS_PER_WARP 1      // fused kernel: mask units are split finer until each warp has this many
#endif
#ifndef AIDET_NMS_FUSED_MINB
#define AIDET_NMS_FUSED_MINB 1          // fused kernel: resident CTAs per SM the register budget allows (512 threads x 128 registers)
#endif
#ifndef AIDET_NMS_FUSED_CTAS
#define AIDET_NMS_FUSED_CTAS 1          // fused kernel: CTAs per SM (the grid barriers cost grows with the CTA count)
#endif
#ifndef AIDET_NMS_P1_WAVES
#define AIDET_NMS_P1_WAVES 2            // CTAs per SM that take part in the fused kernel's ranking phase
#endif
#ifndef AIDET_NMS_SMALL_RING
#define AIDET_NMS_SMALL_RING 18432     // ring words of problems of <= 2048 boxes (72 KB: two CTAs per SM in the mask phase)
#endif
constexpr int kSmallRing = AIDET_NMS_SMALL_RING;
constexpr int kPanelBlocks = 64;           // groups of <= 64 blocks (2048 boxes): the chain's words live in a 32 KB panel

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v;
}

// key of box j inside its group: better boxes have smaller keys (score descending, then index ascending)
__device__ __forceinline__ uint64_t bucket_key(const float* __restrict__ scores, int j) {
  return ((uint64_t)(~orderable(__ldg(scores + j))) << 13) | (uint64_t)j;
}
__device__ __forceinline__ int bucket_group(const int* __restrict__ groups, int j, int n_groups) {
  const uint32_t g = groups ? (uint32_t)__ldg(groups + j) : 0u;
  return (int)min(g, (uint32_t)n_groups);     // ids outside [0, n_groups) go to a bucket behind every group: never scanned
}

// ONE CTA of 512 threads per SM with up to 128 registers.  The two grid barriers cost 1.5 + 2.2 us with 148 CTAs, 1.8 + 2.5
// with 296 and 3.5 + 5.5 with 592 (4 CTAs of 256 threads and 64 registers, where the overlap arithmetic also spilled);
// measured at config C2 43 -> 35 us per call for 592 -> 296 CTAs and another ~1.2 us of phase time for 296 -> 148, at C1
// 74 -> 66 -> 64 us.  Warps 8 .. 15 sit the scan phase out (six helper warps are enough; fourteen were slower).
template <class O, bool GE>
__global__ void __launch_bounds__(kFusedThreads, AIDET_NMS_FUSED_MINB)
nms_fused_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ groups,
                 int n, int n_groups, const float* __restrict__ thr, int n_thr, float one, int rc_rows,
                 typename O::Row* rows, typename O::Col* cols, int* order, uint8_t* flags, int* gstart, int* gend,
                 int* done, int* ticket, uint32_t* mask32, long long pitch32, int ring_words, int p1_ctas,
                 long long* __restrict__ keep_out, int* __restrict__ n_keep, long long* __restrict__ stamps) {
  using Row = typename O::Row; using Col = typename O::Col;
  extern __shared__ __align__(128) unsigned char dyn[];
  // phase-2 tables (unit prefix, group starts / ends) live in the dynamic shared memory, which is free between the
  // key array of phase 1 and the ring of phase 3: (n_groups + 2) ints each
  int* const sprefix = reinterpret_cast<int*>(dyn);
  int* const sstart = sprefix + (n_groups + 2);
  int* const send = sstart + (n_groups + 2);
  // phase 2 also stages each warp's unit rows there (8 warps x 32 records behind the tables, 128 B aligned)
  Row* const wrows = reinterpret_cast<Row*>(dyn + (((size_t)(n_groups + 2) * 12 + 127) & ~(size_t)127));
  __shared__ int swarp[kFusedThreads / 32];
  __shared__ uint32_t s_ww[kFusedThreads];     // phase 2: a warp's 32 row words while its lanes walk their own candidates
  __shared__ __align__(8) uint64_t bar_full[kScanSlotsMax], bar_k[kScanSlotsMax];
  __shared__ uint32_t s_issued;              // blocks whose copy the producer has issued
  __shared__ uint32_t s_hprog[kFusedThreads / 32];            // blocks finished by each helper warp (written by its lane 0 only)
  __shared__ uint32_t s_keep[kFusedMaxBoxes / 32];
  __shared__ unsigned long long s_rem64[kFusedMaxBoxes / 32];   // (block + 1) << 32 | removed bits of the block, ONE 8-byte store
  __shared__ int s_last;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kFusedThreads / 32;
  // diagnostics (aidet_prof_enable(2); stamps == nullptr otherwise): phase time stamps of CTA 0 (SM clock) in stamps[0..7],
  // and in stamps[16 + k] the LATEST arrival of any CTA at boundary k (global nanosecond timer; [16] = negated EARLIEST start),
  // left at the end of the workspace for scripts/r2_nms_phases.py
  auto stamp = [&](int k) {
    if (stamps && tid == 0) {
      if (blockIdx.x == 0) stamps[k] = clock64();
      if (k < 8) {
        unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMax(reinterpret_cast<unsigned long long*>(stamps) + 16 + k, k == 0 ? ~t : t);
      }
    }
  };
  stamp(0);

  // exclusive prefix of v(g), g < count, into sprefix[0..count] (all threads call; count <= kFusedMaxGroups + 1)
  auto cta_prefix = [&](auto value_of, int count) {
    int carry = 0;
    for (int base = 0; base < count; base += kFusedThreads) {
      const int g = base + tid;
      const int v = (g < count) ? value_of(g) : 0;
      int x = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
      if (lane == 31) swarp[warp] = x;
      __syncthreads();
      int wsum = 0, tot = 0;
      for (int w = 0; w < kWarps; ++w) { const int sw = swarp[w]; if (w < warp) wsum += sw; tot += sw; }
      if (g < count) sprefix[g] = carry + wsum + x - v;
      carry += tot;
      __syncthreads();
    }
    if (tid == 0) sprefix[count] = carry;
    __syncthreads();
  };

  // ---------------------------------------------------------------- phase 1: rank + prepare + group bounds
  {
    // every CTA that owns units keeps ALL keys in shared memory.  The loads are issued in batches of 8 per thread
    // before the first key is built (a plain loop exposes one L2 latency per iteration: 23 of them at config C2)
    uint64_t* skeys = reinterpret_cast<uint64_t*>(dyn);
    const int u_box = (n + kRankBoxes - 1) / kRankBoxes, u_all = u_box + n_groups + 1;
    // p1_ctas CTAs take part (all of them with one CTA per SM; the knob dates from grids of several CTAs per SM)
    const int ctas_p1 = min(p1_ctas, (u_all + kWarps - 1) / kWarps);
    if (blockIdx.x == 0 && tid == 0) { *done = 0; *ticket = 0; }
    if ((int)blockIdx.x < ctas_p1) {
      auto put_key = [&](int j, float sc, int gr) {
        const uint32_t g = min((uint32_t)gr, (uint32_t)n_groups);     // ids outside [0, n_groups): behind every group, never scanned
        skeys[j] = ((uint64_t)g << kRankGroupShift) | ((uint64_t)(~orderable(sc)) << 13) | (uint64_t)j;
      };
      const bool vec = ((reinterpret_cast<uintptr_t>(scores) | reinterpret_cast<uintptr_t>(groups)) & 15) == 0;
      if (vec) {
        // four boxes per 128-bit load, four loads of each array in flight per thread: two round trips for n = 8192
        const int nq = n >> 2;
        for (int q0 = 0; q0 < nq; q0 += 4 * kFusedThreads) {
          float4 sc[4]; int4 gr[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int qq = q0 + k * kFusedThreads + tid;
            sc[k] = (qq < nq) ? __ldg(reinterpret_cast<const float4*>(scores) + qq) : make_float4(0.f, 0.f, 0.f, 0.f);
            gr[k] = (qq < nq && groups) ? __ldg(reinterpret_cast<const int4*>(groups) + qq) : make_int4(0, 0, 0, 0);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int qq = q0 + k * kFusedThreads + tid;
            if (qq < nq) {
              put_key(4 * qq, sc[k].x, gr[k].x); put_key(4 * qq + 1, sc[k].y, gr[k].y);
              put_key(4 * qq + 2, sc[k].z, gr[k].z); put_key(4 * qq + 3, sc[k].w, gr[k].w);
            }
          }
        }
        for (int j = 4 * nq + tid; j < n; j += kFusedThreads) put_key(j, __ldg(scores + j), groups ? __ldg(groups + j) : 0);
      } else {
        for (int j0 = 0; j0 < n; j0 += 8 * kFusedThreads) {
          float sc[8]; int gr[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int j = j0 + k * kFusedThreads + tid;
            sc[k] = (j < n) ? __ldg(scores + j) : 0.f;
            gr[k] = (j < n && groups) ? __ldg(groups + j) : 0;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int j = j0 + k * kFusedThreads + tid;
            if (j < n) put_key(j, sc[k], gr[k]);
          }
        }
      }
      __syncthreads();
      // Candidates usually arrive class by class (multiclass_nms builds them that way): when the group ids are
      // non-decreasing a group is a contiguous index range, so a box is ranked against its own group only and the
      // group bounds fall out of the boundary search -- 15x less counting at config C2.  Both paths give the same order.
      int* gfirst = reinterpret_cast<int*>(skeys + n);                 // [n_groups + 1] first index / one past the last index
      int* glast = gfirst + (n_groups + 1);
      for (int g = tid; g <= n_groups; g += kFusedThreads) { gfirst[g] = 0; glast[g] = 0; }
      __syncthreads();
      int unsorted = 0;
      for (int j = tid; j < n; j += kFusedThreads) {
        const int gj = (int)(skeys[j] >> kRankGroupShift), gp = j ? (int)(skeys[j - 1] >> kRankGroupShift) : -1;
        unsorted |= gj < gp;
        if (gj != gp) { gfirst[gj] = j; if (j) glast[gp] = j; }
        if (j == n - 1) glast[gj] = n;
      }
      const bool sorted = !__syncthreads_or(unsorted);
      stamp(13);
      if (sorted && blockIdx.x == 0)
        for (int g = tid; g < n_groups; g += kFusedThreads) { gstart[g] = gfirst[g]; gend[g] = glast[g]; }
      const int n_p1warps = ctas_p1 * kWarps;
      for (int u = blockIdx.x * kWarps + warp; u < (sorted ? u_box : u_all); u += n_p1warps) {
        if (u < u_box) {
          const int i0 = u * kRankBoxes, mine = i0 + lane;
          // lanes 0..3 prepare the unit's boxes first: the FP64 sincos of rect_prepare overlaps the other warps' counting
          Row r; Col cc;
          if (lane < kRankBoxes && mine < n) {
            float bx[O::FMT];
#pragma unroll
            for (int k = 0; k < O::FMT; k++) bx[k] = __ldg(boxes + (size_t)mine * O::FMT + k);
            O::prepare(bx, one, &r, &cc);
          }
          uint64_t ki[kRankBoxes];
#pragma unroll
          for (int b = 0; b < kRankBoxes; ++b) ki[b] = (i0 + b < n) ? skeys[i0 + b] : 0ull;
          // sorted: the unit's boxes lie in groups g_first .. g_last, whose members are the indices [j_lo, j_hi); every key
          // before j_lo is smaller than theirs (lower group), every key from j_hi on larger
          int j_lo = 0, j_hi = n;
          if (sorted) {
            j_lo = gfirst[(int)(ki[0] >> kRankGroupShift)];
            j_hi = glast[(int)(skeys[min(i0 + kRankBoxes, n) - 1] >> kRankGroupShift)];
          }
          int c[kRankBoxes] = {0, 0, 0, 0};
#pragma unroll 4
          for (int j = j_lo + lane; j < j_hi; j += 32) {
            const uint64_t kj = skeys[j];
#pragma unroll
            for (int b = 0; b < kRankBoxes; ++b) c[b] += (kj < ki[b]) ? 1 : 0;
          }
#pragma unroll
          for (int b = 0; b < kRankBoxes; ++b) c[b] = __reduce_add_sync(0xffffffffu, c[b]);
          if (lane < kRankBoxes && mine < n) {
            int rank = c[0];
#pragma unroll
            for (int b = 1; b < kRankBoxes; ++b) if (lane == b) rank = c[b];
            rank += j_lo;
            order[rank] = mine;
            flags[mine] = 0;
            rows[rank] = r;
            if constexpr (!std::is_same<Row, Col>::value) cols[rank] = cc;
          }
        } else {
          const int g = u - u_box;                                     // first sorted position of group g (g == n_groups: end)
          const uint64_t kg = (uint64_t)g << kRankGroupShift;
          int cnt = 0;
          for (int j = lane; j < n; j += 32) cnt += (skeys[j] < kg) ? 1 : 0;
          cnt = __reduce_add_sync(0xffffffffu, cnt);
          if (lane == 0) {
            if (g < n_groups) gstart[g] = cnt;
            if (g > 0) gend[g - 1] = cnt;
          }
        }
      }
    }
  }
  stamp(1);
  grid.sync();
  stamp(2);

  // ---------------------------------------------------------------- phase 2: suppression mask, warp-level units
  const int q = 32 / rc_rows;
  {
    // unit prefix over the groups; group starts / sizes / thresholds stay in shared memory (no L2 round trip per unit)
    for (int g = tid; g < n_groups; g += kFusedThreads) { sstart[g] = __ldcg(gstart + g); send[g] = __ldcg(gend + g); }
    __syncthreads();
    stamp(14);
    cta_prefix([&](int g) { return fused_units(send[g] - sstart[g], q, rc_rows); }, n_groups);
    stamp(15);
    const int total = sprefix[n_groups];
    // static, interleaved: run r = units 8r .. 8r+7 (one per warp: same strip as a rule, so the column record is
    // fetched from L2 once per CTA) goes to CTA r % grid -- neighbouring runs cost alike (same group, same place in the
    // image), so dealing them out round-robin balances the CTAs without a work counter
    int u = blockIdx.x * kWarps + warp;
    const int u_hi = total, u_step = gridDim.x * kWarps;
    int cur_g = -1, cur_c = -1, start = 0, ng = 0;
    typename O::Reg me; float area_me = 0.f, th = 0.f; bool zero_hit = false;
    while (u < u_hi) {
      int lo = 0, hi = n_groups;                                     // last g with sprefix[g] <= u
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sprefix[mid] <= u) lo = mid; else hi = mid; }
      const int g = lo;
      if (g != cur_g) { start = sstart[g]; ng = send[g] - start; }
      const int ns = (ng + 31) >> 5;
      const int l = u - sprefix[g];
      // strip c: the largest c with q c (c+1) / 2 <= l, capped at ns - 1
      int c = (int)((sqrtf(8.0f * (float)l / (float)q + 1.0f) - 1.0f) * 0.5f);
      c = min(c, ns - 1);
      while (c > 0 && q * (c * (c + 1) / 2) > l) --c;
      while (c < ns - 1 && q * ((c + 1) * (c + 2) / 2) <= l) ++c;
      const int r0 = (l - q * (c * (c + 1) / 2)) * rc_rows;
      const int c0 = c << 5;
      const int r_end = min(min(ng, c0 + 32), r0 + rc_rows);
      const int j = c0 + lane;
      const bool live = j < ng;
      if (u < u_step) stamp(26);
      {                                                              // the unit's rows are contiguous records: 16-byte
        constexpr int RQ = (int)sizeof(Row) / 16;                     // pieces into the warp's slice of shared memory.  Their
        const float4* src = reinterpret_cast<const float4*>(rows + start + r0);   // loads are issued BEFORE the column record's,
        float4* dst = reinterpret_cast<float4*>(wrows + warp * 32);   // so a unit costs one L2 round trip, not two
        const int pieces = (r_end - r0) * RQ;
        float4 v[RQ];
#pragma unroll
        for (int k = 0; k < RQ; ++k) if (lane + 32 * k < pieces) v[k] = __ldcg(src + lane + 32 * k);
        if (g != cur_g || c != cur_c) {
          me = cols[start + (live ? j : ng - 1)];
          th = __ldg(thr + (n_thr == 1 ? 0 : g));
          area_me = O::area_c(me, one);
          zero_hit = GE ? (0.0f >= th) : (0.0f > th);
          cur_g = g; cur_c = c;
        }
        __syncwarp();                                                 // the previous unit's reads are done
#pragma unroll
        for (int k = 0; k < RQ; ++k) if (lane + 32 * k < pieces) dst[lane + 32 * k] = v[k];
        __syncwarp();
      }
      if (u < u_step) stamp(27);
      const Row* rr = wrows + warp * 32 - r0;                         // rr[i] = row i of the group, r0 <= i < r_end
      uint32_t word = 0;
      bool dense = false;
      if constexpr (!GE && O::FMT != 4)                               // unit above the diagonal, every circle meets its first and last row
        dense = r_end <= c0 && r_end > r0 && th >= 1e-4f &&
                __all_sync(0xffffffffu, live && !O::disjoint(rr[r0], me) && !O::disjoint(rr[r_end - 1], me));
      if (dense) {
        if constexpr (!GE && O::FMT != 4) {
          const float k1 = 1.0f + th, thb = th * area_me;
#pragma unroll 2
          for (int i = r0; i < r_end; ++i) {
            const uint32_t bb = __ballot_sync(0xffffffffu, nms_hit_dense<O>(rr[i], me, area_me, one, th, k1, thb));
            if (lane == (i & 31)) word = bb;
          }
        }
      } else if (O::FMT != 4 && !zero_hit && AIDET_NMS_COMPACT_PAIRS) {
        // Sparse unit.  A row costs the warp the full overlap arithmetic as soon as ONE of its 32 columns passes the circle
        // test, and with 32 columns that is nearly every row even on DOTA-shaped boxes (the unit's columns are neighbours
        // in score order, not in space).  So: circle tests for all rows first (a candidate bit per row in each lane),
        // then every lane walks ITS OWN candidates -- the column stays in its registers, the row record comes from the
        // warp's staged rows -- and the warp runs max-over-lanes(candidates) passes instead of one per row.
        uint32_t cand = 0;
        for (int i = r0; i < r_end; ++i)
          if (live && j > i && !O::disjoint(rr[i], me)) cand |= 1u << (i - r0);
        uint32_t* ww = s_ww + warp * 32;
        ww[lane] = 0u;
        __syncwarp();
        while (cand) {
          const int k = __ffs(cand) - 1;
          cand &= cand - 1;
          const Row& a = rr[r0 + k];
          const float area_a = O::area(a, one);
          float inter = O::inter(a, me, one);
          inter = fminf(fmaxf(inter, 0.0f), fminf(area_a, area_me));
          const float den = area_a + area_me - inter;
          const float rhs = th * den;
          const bool h = GE ? (inter >= rhs) : (inter > rhs);
          if (h && den > 0.0f) atomicOr(&ww[k], 1u << lane);
        }
        __syncwarp();
        const int k_mine = (r0 & ~31) + lane - r0;
        if (k_mine >= 0 && k_mine < r_end - r0) word = ww[k_mine];
        __syncwarp();
      } else {
#pragma unroll 2
        for (int i = r0; i < r_end; ++i) {
          const bool hit = nms_hit<O, GE>(rr[i], me, area_me, one, th, zero_hit);
          const uint32_t bb = __ballot_sync(0xffffffffu, hit && live && j > i);
          if (lane == (i & 31)) word = bb;
        }
      }
      // rows r0 .. r_end-1 live in one 32-row block (rc_rows divides 32): lane (i & 31) holds row i's word.
      // Blocked layout: word (block b, word c, row r) of the group at ((b * pitch32 + c) * 32 + r).
      if (u < u_step) stamp(28);
      const int i_mine = (r0 & ~31) + lane;
      if (i_mine >= r0 && i_mine < r_end)
        mask32[(long long)(start + 32 * g) * pitch32 + ((long long)(r0 >> 5) * pitch32 + c) * 32 + lane] = word;
      if (u < u_step) stamp(24);
      u += u_step;
    }
  }
  stamp(25);
  __syncthreads();
  stamp(3);
  grid.sync();
  stamp(4);

  // ---------------------------------------------------------------- phase 3: greedy scan, one CTA per group
  uint32_t* ring = reinterpret_cast<uint32_t*>(dyn);
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int start = __ldcg(gstart + g), ng = __ldcg(gend + g) - start;
    if (ng <= 0) continue;
    const int nhw = (ng + 31) >> 5;
    const int slot_words = 32 * nhw;
    // groups of <= 2048 boxes keep words b .. b+3 of every block b (all the chain reads) in a 32 KB panel at the end
    // of the dynamic shared memory, loaded once; the ring then only feeds the helpers
    const bool use_panel = nhw <= kPanelBlocks && (ring_words - nhw * 128) / slot_words >= 2;
    uint32_t* panel = ring + (ring_words - nhw * 128);
    const int D = min(kScanSlotsMax, (use_panel ? ring_words - nhw * 128 : ring_words) / slot_words);   // >= 2 (host)
    const uint32_t* mgrp = mask32 + (long long)(start + 32 * g) * pitch32;
    __syncthreads();                                                   // previous group done with the barriers / s_keep
    if (nhw <= kSoloBlocks) {
      // ---- small group (<= 512 boxes: a class of a per-image problem): its whole upper triangle (<= 17 KB) is copied into
      //      shared memory by all warps, then ONE warp scans it with no cross-warp hand-off at all -- per block the 32
      //      greedy decisions plus one redux.or per later word; lane w keeps the removed bits of block w in a register.
      //      (The warp-specialised form below pays ~0.9 us per block in flag round trips on such groups: C2 11.6 us.)
      uint32_t* tri = ring;                                            // block b at tri + 32 (b nhw - b (b - 1) / 2)
      for (int b = warp; b < nhw; b += kWarps) {
        const uint4* src = reinterpret_cast<const uint4*>(mgrp + ((long long)b * pitch32 + b) * 32);
        uint4* dst = reinterpret_cast<uint4*>(tri + 32 * (b * nhw - b * (b - 1) / 2));
        for (int i = lane; i < (nhw - b) * 8; i += 32) dst[i] = __ldcg(src + i);
      }
      __syncthreads();
      if (warp == 0) {
        uint32_t rem = 0;                                              // lane w: rows of block w suppressed so far
        for (int b = 0; b < nhw; ++b) {
          const int rows_b = min(32, ng - 32 * b);
          const uint32_t* blk = tri + 32 * (b * nhw - b * (b - 1) / 2);
          uint32_t cur = __shfl_sync(0xffffffffu, rem, b);
          if (rows_b < 32) cur |= ~0u << rows_b;
          uint32_t keep = 0;
          if (cur != ~0u) {
            keep = greedy_block(cur, blk[lane], lane);                 // diagonal word of row 32 b + lane
            const bool kept = (keep >> lane) & 1u;
            for (int w = b + 1; w < nhw; w += 4) {                     // kept rows' words w .. w+3 -> four independent redux.or
              uint32_t v[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) v[q] = (kept && w + q < nhw) ? blk[(w + q - b) * 32 + lane] : 0u;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t r = __reduce_or_sync(0xffffffffu, v[q]);
                if (lane == w + q) rem |= r;
              }
            }
          }
          if (lane == 0) s_keep[b] = keep;
        }
      }
      __syncthreads();
      for (int i = tid; i < ng; i += kFusedThreads)
        if ((s_keep[i >> 5] >> (i & 31)) & 1u) flags[order[start + i]] = 1;
      continue;
    }
    if (tid < kScanSlotsMax) { mbar_init(&bar_full[tid], 1); mbar_init(&bar_k[tid], 1); fence_barrier_init(); }
    for (int h = tid; h < nhw; h += kFusedThreads) { s_keep[h] = 0u; s_rem64[h] = 0ull; }
    if (tid < kWarps) s_hprog[tid] = 0u;
    if (tid == kWarps) s_issued = 0u;
    if (use_panel) {
      // block b: its words b .. b+3 are 128 consecutive words of the blocked layout -> one 16-byte load per lane
#pragma unroll 8
      for (int b = warp; b < nhw; b += kWarps) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (b + (lane >> 3) < nhw) v = __ldcg(reinterpret_cast<const uint4*>(mgrp + ((long long)b * pitch32 + b) * 32) + lane);
        reinterpret_cast<uint4*>(panel + b * 128)[lane] = v;
      }
    }
    __syncthreads();
    if (warp == 1) {
      // ---- producer: block b = words b .. nhw-1 of rows 32b .. 32b+31, contiguous in the blocked layout
      //      (panel mode needs no ring: the chain reads the panel, the helpers stream their words from L2)
      if (lane == 0 && !use_panel) {
        const bool pdiag = stamps && blockIdx.x == 0;
        long long tp_h = 0, tp_l = 0, p_sleeps = 0;
        const long long tp_begin = pdiag ? clock64() : 0;
        int slot = 0; uint32_t par_prev = 1;                           // b % D and ((b - D) / D) & 1 without a division per block
        for (int b = 0; b < nhw; ++b, ++slot) {
          if (slot == D) { slot = 0; par_prev ^= 1u; }
          if (b >= D) {                                                // every helper warp is done with block b - D
            const uint32_t need = (uint32_t)(b - D + 1);
            const long long tp0 = pdiag ? clock64() : 0;
            for (int h = 0; h < kScanHelpers; ++h)
              while (ld_acquire_u32(&s_hprog[h]) < need) { AIDET_NMS_PRODUCER_SLEEP; ++p_sleeps; }
            const long long tp1 = pdiag ? clock64() : 0;
            mbar_wait(&bar_full[slot], par_prev);                      // its copy has landed (the chain skips that wait on dead blocks)
            if (pdiag) { tp_h += tp1 - tp0; tp_l += clock64() - tp1; }
          }
          const uint32_t bytes = (uint32_t)((nhw - b) * 128);
          mbar_expect_tx(&bar_full[slot], bytes);
          tma_load_1d(ring + slot * slot_words, mgrp + ((long long)b * pitch32 + b) * 32, bytes, &bar_full[slot]);
          asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(&s_issued)), "r"((uint32_t)(b + 1)) : "memory");
        }
        if (pdiag) { stamps[24] = tp_h; stamps[25] = tp_l; stamps[26] = p_sleeps; stamps[27] = clock64() - tp_begin; stamps[28] = D; }
      }
    } else if (warp == 0) {
      // ---- chain.  Per block: removed bits of its 32 rows (helpers: blocks <= b-4; own carries: blocks b-3..b-1) ->
      //      32 greedy decisions, two per step -> keep word -> the kept rows' words b+1..b+3 OR-ed into three carries
      //      (redux.or), so the helpers have four blocks of slack.  Groups of <= 2048 boxes read their words b..b+3
      //      from the panel that was loaded once (no barrier probe per block: measured 36 cycles for the probe, 43 for
      //      the issue-counter load, and neither overlaps the chain because both block the warp); larger groups take
      //      them from the ring.
      const bool diag = stamps && blockIdx.x == 0;                     // diagnostics: cycles the chain waits for the helpers
      long long t_chain_wait = 0;
      uint32_t c1 = 0, c2 = 0, c3 = 0;
      uint32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;                         // words b .. b+3 of row 32 b + lane
      bool have = false;                                               // d* hold block b's words
      int slot = 0; uint32_t par = 0;                                  // b % D, (b / D) & 1 without a division per block
      auto fetch = [&](int b, const uint32_t* blk, uint32_t& e0, uint32_t& e1, uint32_t& e2, uint32_t& e3) {
        e0 = e1 = e2 = e3 = 0u;                                        // blk: word w of row `lane` at blk[(w - b) * 32]
        if (lane < ng - 32 * b) {
          e0 = blk[0];
          if (b + 1 < nhw) e1 = blk[32];
          if (b + 2 < nhw) e2 = blk[64];
          if (b + 3 < nhw) e3 = blk[96];
        }
      };
      if (use_panel) { fetch(0, panel + lane, d0, d1, d2, d3); have = true; }
      for (int b = 0; b < nhw; ++b) {
        uint32_t cur = c1;
        if (b >= 4) {
          const long long tw = diag ? clock64() : 0;
          // the word's owner publishes it, stamped with its block, in one 8-byte store once the blocks <= b-4 are in:
          // no flag next to the data, so no fence on the helper's side and no wait for the OTHER helpers
          unsigned long long v;
          do {
            asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(&s_rem64[b])) : "memory");
          } while ((uint32_t)(v >> 32) != (uint32_t)(b + 1));
          cur |= (uint32_t)v;
          if (diag) t_chain_wait += clock64() - tw;
        }
        const int rows_b = min(32, ng - 32 * b);
        if (rows_b < 32) cur |= ~0u << rows_b;
        int slot_n = slot + 1; uint32_t par_n = par;
        if (slot_n == D) { slot_n = 0; par_n ^= 1u; }
        uint32_t keep = 0, n1 = 0, n2 = 0, n3 = 0;
        uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
        bool have_n = false;
        if (use_panel) {
          if (b + 1 < nhw) { fetch(b + 1, panel + (b + 1) * 128 + lane, e0, e1, e2, e3); have_n = true; }
        } else if (cur != ~0u) {
          if (!have) {
            // the barrier counts for THIS block only once the producer has armed it (blocks skipped as dead are not
            // waited for, so the slot's previous phase may still be open)
            while (ld_acquire_u32(&s_issued) <= (uint32_t)b) {}
            mbar_wait(&bar_full[slot], par);
            fetch(b, ring + slot * slot_words + lane, d0, d1, d2, d3);
          }
          if (b + 1 < nhw && ld_acquire_u32(&s_issued) > (uint32_t)(b + 1) && mbar_try_wait(&bar_full[slot_n], par_n)) {
            fetch(b + 1, ring + slot_n * slot_words + lane, e0, e1, e2, e3);
            have_n = true;
          }
        }
        if (cur != ~0u) {                                              // a dead block needs neither its data nor a chain
          keep = greedy_block(cur, d0, lane);
          const bool kept = (keep >> lane) & 1u;
          n1 = __reduce_or_sync(0xffffffffu, kept ? d1 : 0u);
          n2 = __reduce_or_sync(0xffffffffu, kept ? d2 : 0u);
          n3 = __reduce_or_sync(0xffffffffu, kept ? d3 : 0u);
        }
        if (lane == 0) { s_keep[b] = keep; mbar_arrive(&bar_k[b % kScanSlotsMax]); }
        c1 = c2 | n1;
        c2 = c3 | n2;
        c3 = n3;
        d0 = e0; d1 = e1; d2 = e2; d3 = e3; have = have_n;
        slot = slot_n; par = par_n;
      }
      if (diag && lane == 0) { stamps[11] = t_chain_wait; stamps[12] = nhw; }
    } else if (warp < 2 + kScanHelpers) {
      // ---- helpers: warp hw owns the words w = hw + kScanHelpers * i; lane (i % 32) keeps the accumulator of its i-th
      //      word in a register (two per lane cover nhw <= 256) and publishes it when the word's last block is done
      const int hw = warp - 2;
      const bool diag = stamps && blockIdx.x == 0 && hw == 0;          // diagnostics: where helper warp 0 spends its cycles
      long long t_k = 0, t_data = 0, t_work = 0, t_iss = 0, t_iss_max = 0; int b_iss_max = 0, n_iss_slow = 0;
      uint32_t acc0 = 0, acc1 = 0, issued = 0;
      if (use_panel) {
        // groups of <= 2048 boxes: no ring, no producer, no issue counter.  The 192 helper lanes pair up, a pair per word
        // (words 4 .. 63): a lane keeps 16 of the 32 rows of ITS word for the next two blocks in registers (64 contiguous
        // bytes of the blocked layout, four 16-byte loads straight from L2), ANDs them with the keep bits as soon as
        // the chain publishes those, folds the pair with one shuffle and ORs the result into the word's accumulator; the
        // pair's first lane publishes word b + 4 after block b, stamped, in one 8-byte store.  (Measured at config C1: the
        // ring form spent ~900 cycles per block on the helper side -- redux.or per word, two integer divisions for the
        // slot, the issue counter -- and the chain waited ~450 of its ~1190 cycles per block for it.)
        const int gl = hw * 32 + lane, w = gl >> 1, half = gl & 1;
        uint4 cur[4], nxt[4], nx2[4];
        auto load = [&](int b, uint4 (&v)[4]) {
          const bool on = w >= b + 4 && w < nhw && b < nhw;
          const uint4* src = reinterpret_cast<const uint4*>(mgrp + ((long long)b * pitch32 + w) * 32 + half * 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = on ? __ldcg(src + k) : make_uint4(0u, 0u, 0u, 0u);
        };
        load(0, nxt);
        load(1, nx2);
        for (int b = 0; b < nhw; ++b) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { cur[k] = nxt[k]; nxt[k] = nx2[k]; }
          load(b + 2, nx2);
          long long t0 = diag ? clock64() : 0;
          mbar_wait(&bar_k[b % kScanSlotsMax], (b / kScanSlotsMax) & 1);
          const uint32_t keep = ld_acquire_u32(&s_keep[b]);
          if (diag) { const long long t1 = clock64(); t_k += t1 - t0; t0 = t1; }
          if (keep && b + 4 < nhw) {
            const uint32_t kk = keep >> (half * 16);
            uint32_t m = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              m |= ((kk >> (4 * k)) & 1u) ? cur[k].x : 0u;
              m |= ((kk >> (4 * k + 1)) & 1u) ? cur[k].y : 0u;
              m |= ((kk >> (4 * k + 2)) & 1u) ? cur[k].z : 0u;
              m |= ((kk >> (4 * k + 3)) & 1u) ? cur[k].w : 0u;
            }
            m |= __shfl_xor_sync(0xffffffffu, m, 1);
            acc0 |= m;
          }
          if (w == b + 4 && w < nhw && half == 0) {
            const unsigned long long v = ((unsigned long long)(uint32_t)(w + 1) << 32) | acc0;
            asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(smem_u32(&s_rem64[w])), "l"(v) : "memory");
          }
          if (diag) t_work += clock64() - t0;
        }
      } else {
      int slot = 0; uint32_t par = 0;                                  // b % D, (b / D) & 1 without a division per block
      for (int b = 0; b < nhw; ++b, ++slot) {
        if (slot == D) { slot = 0; par ^= 1u; }
        long long t0 = diag ? clock64() : 0;
        mbar_wait(&bar_k[b % kScanSlotsMax], (b / kScanSlotsMax) & 1);
        const uint32_t keep = ld_acquire_u32(&s_keep[b]);
        if (diag) { const long long t1 = clock64(); t_k += t1 - t0; t0 = t1; }
        if (keep && b + 4 < nhw) {
          // the block's copy must have landed (in panel mode the chain never waits for the ring); the barrier counts for
          // this block only once the producer has armed it
          while (issued <= (uint32_t)b) issued = ld_acquire_u32(&s_issued);   // cached: the producer runs blocks ahead
          if (diag) { const long long t1 = clock64(); t_iss += t1 - t0; if (t1 - t0 > t_iss_max) { t_iss_max = t1 - t0; b_iss_max = b; } n_iss_slow += (t1 - t0 > 200); }
          mbar_wait(&bar_full[slot], par);
          if (diag) { const long long t1 = clock64(); t_data += t1 - t0; t0 = t1; }
          const uint32_t* blk = ring + slot * slot_words + lane - b * 32;
          const bool kept = (keep >> lane) & 1u;
          int i = (b + 4 - hw + kScanHelpers - 1) / kScanHelpers;      // first owned word >= b + 4
          i = max(i, 0);
          for (int w = hw + kScanHelpers * i; w < nhw; w += 4 * kScanHelpers, i += 4) {
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int wk = w + k * kScanHelpers; v[k] = (wk < nhw && kept) ? blk[wk * 32] : 0u; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t r = __reduce_or_sync(0xffffffffu, v[k]);
              const int ik = i + k;
              if (lane == (ik & 31)) { if (ik < 32) acc0 |= r; else acc1 |= r; }
            }
          }
        }
        // word b + 4 has now received every contribution the helpers owe it: publish it if this warp owns it
        const int wf = b + 4;
        if (wf < nhw && wf % kScanHelpers == hw) {
          const int i = wf / kScanHelpers;
          if (lane == (i & 31)) {
            const unsigned long long v = ((unsigned long long)(uint32_t)(wf + 1) << 32) | ((i < 32) ? acc0 : acc1);
            asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(smem_u32(&s_rem64[wf])), "l"(v) : "memory");
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(&s_hprog[hw])), "r"((uint32_t)(b + 1)) : "memory");
        if (diag) t_work += clock64() - t0;
      }
      }
      if (diag && lane == 0) { stamps[8] = t_k; stamps[9] = t_data; stamps[10] = t_work; stamps[29] = t_iss; stamps[30] = t_iss_max; stamps[31] = b_iss_max * 1000 + n_iss_slow; }
    }
    __syncthreads();
    for (int i = tid; i < ng; i += kFusedThreads)
      if ((s_keep[i >> 5] >> (i & 31)) & 1u) flags[order[start + i]] = 1;
  }

  stamp(5);
  // ---------------------------------------------------------------- compaction by the CTA that finishes last
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(done, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last || !keep_out) return;                                   // keep_out == nullptr: the caller only wants the flags
  __threadfence();
  const int per = ((n + kFusedThreads - 1) / kFusedThreads + 15) & ~15;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  const uint4* f4 = reinterpret_cast<const uint4*>(flags);            // workspace slot is 128 B aligned and padded
  int cnt = 0;
  for (int i = lo; i < hi; i += 16) {
    const uint4 v = __ldcg(f4 + (i >> 4));
    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; ++k) if (i + k < hi) cnt += (wds[k >> 2] >> (8 * (k & 3))) & 1u;
  }
  int x = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
  if (lane == 31) swarp[warp] = x;
  __syncthreads();
  int off = x - cnt;
  for (int w = 0; w < warp; ++w) off += swarp[w];
  for (int i = lo; i < hi; i += 16) {
    const uint4 v = __ldcg(f4 + (i >> 4));
    const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (i + k < hi && ((wds[k >> 2] >> (8 * (k & 3))) & 1u)) keep_out[off++] = i + k;
  }
  if (tid == kFusedThreads - 1) { *n_keep = off; if (stamps) { stamps[6] = clock64(); stamps[7] = blockIdx.x; } }
}

// ------------------------------------------------------------------ host side
struct NmsLayout {
  size_t keys_in, keys_out, idx_in, order, rows, cols, gbounds, prefix, flags, mask, cub, gcnt, gcursor, total;
  long long pitch32;
  size_t cub_bytes;
};

static int group_bits(int n_groups) { int b = 0; while ((1LL << b) < (long long)n_groups) ++b; return b; }

static NmsLayout nms_layout(int n, int n_groups, int fmt, size_t cub_bytes) {
  NmsLayout L;
  const size_t rec = (fmt == 8) ? 64 : (fmt == 5 ? 48 : 16);            // row records
  const size_t rec_c = (fmt == 8) ? 64 : (fmt == 5 ? 32 : 0);           // column records (fmt 4: the row records serve)
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 128); return o; };
  L.keys_in = take((size_t)n * 8); L.keys_out = take((size_t)n * 8);
  L.idx_in = take((size_t)n * 4); L.order = take((size_t)n * 4);
  L.rows = take((size_t)n * rec); L.cols = take((size_t)n * rec_c);
  L.gbounds = take((size_t)n_groups * 8); L.prefix = take((size_t)(n_groups + 3) * 4);
  L.flags = take((size_t)n + 16);
  L.pitch32 = 4LL * ((n + 127) / 128);                       // multiple of 4 words: rows are 16 B aligned (TMA bulk copies of the fused scan)
  // the fused kernel stores every group's mask in 32-row blocks: up to 31 rows of padding per group
  const size_t mask_rows = (size_t)n + ((n <= 8192 && n_groups <= 1024) ? 32 * (size_t)n_groups : 0);
  L.mask = take(mask_rows * (size_t)L.pitch32 * 4);
  L.cub_bytes = cub_bytes; L.cub = take(cub_bytes);
  L.gcnt = take((size_t)(n_groups + 2) * 4); L.gcursor = take((size_t)(n_groups + 2) * 4);      // bucket counts / bases, cursors
  L.total = off + 384;                                       // the last 256 bytes: phase stamps of the fused kernel (diagnostics)
  return L;
}

static int cub_temp_bytes(int n, int end_bit, size_t* bytes) {
  *bytes = 0;
  AIDET_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, *bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                             (const int*)nullptr, (int*)nullptr, n, 0, end_bit, (cudaStream_t)0));
  return AIDET_OK;
}

static bool coop_supported(int device) {
  static int cached[64];                                    // 0 unknown, 1 yes, 2 no (benign race: same value)
  if (device < 0 || device >= 64) return false;
  if (!cached[device]) {
    int v = 0;
    cached[device] = (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, device) == cudaSuccess && v) ? 1 : 2;
  }
  return cached[device] == 1;
}

// resident CTAs per SM of the fused kernel with `smem` bytes of dynamic shared memory (0: does not fit)
static int fused_occupancy(void* fn, size_t smem) {
  if (smem > 200 * 1024) return 0;
  // always opt in: the 48 KB default limit counts the kernel's static shared memory too (a 48 KB dynamic request
  // alone is refused and the occupancy query then answers 0)
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max(smem, (size_t)48 * 1024)) != cudaSuccess) return 0;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kFusedThreads, smem) != cudaSuccess) return 0;
  return occ;
}

template <class O>
static int run_nms(const float* boxes, const float* scores, const int* groups, int n, const float* thr, int n_thr,
                   int n_groups, int cmp, float one, long long* keep_out, int* n_keep, char* ws, const NmsLayout& L,
                   int device, cudaStream_t s) {
  using Row = typename O::Row; using Col = typename O::Col;
  uint64_t* keys_in = (uint64_t*)(ws + L.keys_in); uint64_t* keys_out = (uint64_t*)(ws + L.keys_out);
  int* idx_in = (int*)(ws + L.idx_in); int* order = (int*)(ws + L.order);
  Row* rows = (Row*)(ws + L.rows);
  Col* cols = std::is_same<Row, Col>::value ? (Col*)rows : (Col*)(ws + L.cols);
  int* gstart = (int*)(ws + L.gbounds); int* gend = gstart + n_groups;
  int* prefix = (int*)(ws + L.prefix);
  uint8_t* flags = (uint8_t*)(ws + L.flags);
  uint32_t* mask32 = (uint32_t*)(ws + L.mask);

  int* counters = prefix + n_groups + 1;                    // [0] tile ticket of the mask kernel, [1] finished scan CTAs
  ProfScope prof(PROF_NMS_MASK, s);                          // device time of the WHOLE call (every kernel it enqueues)
  if (n <= kFusedMaxBoxes && n_groups <= kFusedMaxGroups && coop_supported(device)) {
    // one cooperative launch: rank -> mask -> scan -> compaction (see nms_fused_kernel)
    const int slot_max = 32 * ((n + 31) >> 5);               // ring slot if all n boxes fall into one group
    // dynamic shared memory: all n keys (phase 1), later the scan's ring (+ the chain's panel for groups <= 2048 boxes,
    // carved from its end).  Any group must get two ring slots; beyond that, problems of more than 2048 boxes stay at
    // 48 KB (per-image inputs hold many small groups, whose slots are tiny), smaller
    // ones -- possibly ONE group -- take up to 72 KB so that the helpers' ring runs several blocks ahead
    const int ring_words = (n > 32 * kPanelBlocks) ? max(2 * slot_max, 12288) : min(kSmallRing, 12 * slot_max + kPanelBlocks * 128);
    const size_t smem = max(max((size_t)n * 8 + (size_t)(n_groups + 1) * 8, (size_t)(n_groups + 2) * 12 + 128 + (kFusedThreads / 32) * 32 * sizeof(Row)),
                            (size_t)ring_words * 4);
    void* fn = (cmp == AIDET_CMP_GE) ? (void*)nms_fused_kernel<O, true> : (void*)nms_fused_kernel<O, false>;
    const int occ = fused_occupancy(fn, smem);
    if (occ > 0) {
      const int sms = sm_count(device);
      const int grid = sms * min(occ, AIDET_NMS_FUSED_CTAS);
      const int n_gwarps = grid * (kFusedThreads / 32);
      // rows per mask unit: the finest split until there are >= 8 units per warp (estimated for evenly filled groups)
      const int side = max(1, n / n_groups);
      int rc_rows = 4;
      while (rc_rows < 32 && (long long)n_groups * fused_units(side, 32 / rc_rows, rc_rows) >= (long long)AIDET_NMS_UNITS_PER_WARP * n_gwarps) rc_rows <<= 1;
      int nn = n, ng_ = n_groups, nthr = n_thr, rwords = ring_words, p1 = min(grid, sms * AIDET_NMS_P1_WAVES);
      long long pitch = L.pitch32;
      int* done = counters + 1;
      int* ticket = counters;
      long long* stamps = prof_level() >= 2 ? (long long*)(ws + L.total - 256) : nullptr;   // phase stamps, see the kernel
      void* args[] = {(void*)&boxes, (void*)&scores, (void*)&groups, &nn, &ng_, (void*)&thr, &nthr, &one, &rc_rows,
                      &rows, &cols, &order, &flags, &gstart, &gend, &done, &ticket, &mask32, &pitch, &rwords, &p1, &keep_out, &n_keep,
                      &stamps};
      AIDET_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kFusedThreads), args, smem, s));
      count_launch(1);
      return AIDET_OK;
    }
  }
  const bool small = n <= kRankSortMax && n_groups <= kRankMaxGroups;
  if (small) {
    const int warps = max(ceil_div(n, kRankBoxes), n_groups + 1);
    const size_t key_smem = (size_t)n * 8;
    if (key_smem > 48 * 1024)
      AIDET_CUDA(cudaFuncSetAttribute(nms_rank_gather_kernel<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)key_smem));
    nms_rank_gather_kernel<O><<<ceil_div(warps, kRankWarps), kRankWarps * 32, key_smem, s>>>(
        boxes, scores, groups, n, one, rows, cols, order, flags, gstart, gend, n_groups, counters);
  } else {
    // rank-by-counting costs ~0.12 warp instructions per key pair inside a bucket: against the mask kernel's 0.19 (sparse
    // groups, upper triangle) to 2.1 (dense) per pair it pays off while buckets stay below a few 10 k boxes; beyond that
    // average size (one class of a whole scene) the radix sort's fixed ~70 us is cheaper
    if (kBucketSort && (long long)n <= 24576LL * n_groups) {
      unsigned* gcnt = (unsigned*)(ws + L.gcnt); unsigned* gcursor = (unsigned*)(ws + L.gcursor);
      AIDET_CUDA(cudaMemsetAsync(gcnt, 0, (size_t)(n_groups + 2) * 4, s));
      nms_bucket_hist_kernel<<<ceil_div(n, 256), 256, 0, s>>>(groups, n, n_groups, gcnt, flags);
      nms_bucket_scan_kernel<<<1, 1024, 0, s>>>(gcnt, gcursor, n_groups, gstart, gend);
      nms_bucket_scatter_kernel<<<ceil_div(n, 256), 256, 0, s>>>(scores, groups, n, n_groups, gcursor, keys_in, idx_in);
      nms_bucket_rank_kernel<O><<<ceil_div(n, kBucketCta), 256, 0, s>>>(boxes, keys_in, idx_in, gcnt, n, n_groups, one, rows, cols, order);
    } else {
    const int nb = ceil_div(max(n, 2 * n_groups), 256);
    nms_keys_kernel<<<nb, 256, 0, s>>>(scores, groups, n, keys_in, idx_in, flags, gstart, n_groups);
    size_t cub_bytes = L.cub_bytes;
    AIDET_CUDA(cub::DeviceRadixSort::SortPairs(ws + L.cub, cub_bytes, keys_in, keys_out, idx_in, order, n, 0,
                                               32 + (groups ? group_bits(n_groups + 1) : 0), s));   // + 1: ids outside [0, n_groups) are clamped to n_groups and sort last
    nms_gather_kernel<O><<<ceil_div(n, 256), 256, 0, s>>>(boxes, keys_out, order, n, one, rows, cols, gstart, gend, n_groups);
    }
  }
  const int sms = sm_count(device);
  // Row-tile height: the group sizes live on the device, so estimate the tile count as if the boxes were
  // spread evenly over the groups (upper triangle of n_groups squares of side n / n_groups) and shrink the
  // tiles until there are ~8 CTAs per SM -- config C2 (5822 boxes, 15 classes) gets 8-row tiles.
  int tile_rows = kTileRows;
  {
    const double side = (double)n / n_groups;
    auto est_tiles = [&](int tr) { return 0.5 * n_groups * (side / tr + 1.0) * (side / kTileCols + 1.0); };
    while (tile_rows > 8 && est_tiles(tile_rows) < 8.0 * sms) tile_rows >>= 1;
  }
  const int local_prefix = (small && n_groups <= kLocalGroups) ? 1 : 0;
  // groups of (on average) fewer than 2048 boxes: warp-level units instead of 256-column tiles (see nms_mask_units_kernel)
  const bool units = !small && (long long)n < 2048LL * n_groups;
  if (units) {
    const int resident = sms * (O::FMT == 8 ? 3 : 4);
    const int side = max(1, n / n_groups);
    int rc_rows = 4;                                          // the finest split until there are >= 8 units per warp
    while (rc_rows < 32 && (long long)n_groups * fused_units(side, 32 / rc_rows, rc_rows) >= (long long)AIDET_NMS_UNITS_PER_WARP_LARGE * resident * 8) rc_rows <<= 1;
    nms_tile_prefix_kernel<<<1, 1024, 0, s>>>(gstart, gend, n_groups, tile_rows, prefix, rc_rows);
    const int stage = n_groups <= 8192 ? 1 : 0;
    const size_t psmem = stage ? (size_t)(n_groups + 1) * 4 : 0;
    if (cmp == AIDET_CMP_GE)
      nms_mask_units_kernel<O, true><<<resident, 256, psmem, s>>>(rows, cols, gstart, gend, prefix, n_groups, thr, n_thr, one,
                                                                 rc_rows, mask32, L.pitch32, stage);
    else
      nms_mask_units_kernel<O, false><<<resident, 256, psmem, s>>>(rows, cols, gstart, gend, prefix, n_groups, thr, n_thr, one,
                                                                  rc_rows, mask32, L.pitch32, stage);
  } else {
  if (!local_prefix) nms_tile_prefix_kernel<<<1, 1024, 0, s>>>(gstart, gend, n_groups, tile_rows, prefix, 0);
  {
    // upper bound of the tile count: every group padded to full tiles
    long long max_tiles = (long long)(ceil_div(n, tile_rows) + n_groups) * (ceil_div(n, kTileCols) + 1);
    int grid = (int)min((long long)sms * 6, max(max_tiles, 1LL));
    if (cmp == AIDET_CMP_GE)
      nms_mask_kernel<O, true><<<grid, kTileCols, 0, s>>>(rows, cols, gstart, gend, prefix, n_groups, thr, n_thr, one,
                                                          tile_rows, mask32, L.pitch32, counters, local_prefix);
    else
      nms_mask_kernel<O, false><<<grid, kTileCols, 0, s>>>(rows, cols, gstart, gend, prefix, n_groups, thr, n_thr, one,
                                                           tile_rows, mask32, L.pitch32, counters, local_prefix);
  }
  }
  const int removed_cap = ceil_div(ceil_div(n, 32), 4) * 4;          // any group may hold all n boxes
  const int scan_pw = min(kScanPWMax, max(32, ceil_div(ceil_div(n, 32), 32) * 32));
  const int sm_words = removed_cap + 2 * 32 * (scan_pw + 4);
  size_t scan_smem = (size_t)sm_words * 4;
  if (scan_smem > 48 * 1024) {
    if (scan_smem > 227 * 1024) { set_error("nms: %d boxes exceed the single-group scan capacity", n); return AIDET_EINVAL; }
    AIDET_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem));
  }
  const bool fused_compact = n <= kFusedCompactMax && keep_out;       // keep_out == nullptr: the caller only wants the flags
  nms_scan_kernel<<<n_groups, 256, scan_smem, s>>>(mask32, L.pitch32, gstart, gend, order, flags, removed_cap, scan_pw,
                                                   n, keep_out, n_keep, fused_compact ? counters + 1 : nullptr);
  if (!fused_compact && keep_out) nms_compact_kernel<<<1, 1024, 0, s>>>(flags, n, keep_out, n_keep);
  count_launch((small ? (local_prefix ? 3 : 4) : 5) + ((fused_compact || !keep_out) ? 0 : 1));        // + the CUB sort passes, which are library kernels and not counted
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}


// ------------------------------------------------------------------ 7. scene merge (config C5)
// Per-tile NMS + cross-tile merge NMS of one scene (tools/parse_results.py:56-76 / dota.py:296-327 of the reference do
// this per class on the host with DOTA_devkit's py_cpu_nms_poly_fast).  Three entry points, one per stage, so that the
// multi-GPU form can put ONE all-reduce of the keep masks between them; every helper below is one small launch.
__global__ void __launch_bounds__(256) scene_prepare_kernel(const float* __restrict__ boxes, int fmt, const int* __restrict__ labels,
                                                            const int* __restrict__ tile_ids, const float* __restrict__ origins,
                                                            int n, int n_tiles, int n_classes, int world, int rank,
                                                            int* __restrict__ groups, float* __restrict__ scene) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tile_ids[i], l = labels[i];
  const bool valid = t >= 0 && t < n_tiles && l >= 0 && l < n_classes;
  groups[i] = (valid && (world == 1 || t % world == rank)) ? t * n_classes + l : n_tiles * n_classes;
  const float ox = valid ? origins[2 * t] : 0.f, oy = valid ? origins[2 * t + 1] : 0.f;
  const float* b = boxes + (size_t)i * fmt;
  float* o = scene + (size_t)i * fmt;
  if (fmt == 5) {
    o[0] = b[0] + ox; o[1] = b[1] + oy; o[2] = b[2]; o[3] = b[3]; o[4] = b[4];
  } else {
    for (int k = 0; k < fmt; k += 2) { o[k] = b[k] + ox; o[k + 1] = b[k + 1] + oy; }
  }
}

__global__ void __launch_bounds__(256) scene_merge_groups_kernel(const int* __restrict__ labels, const uint8_t* __restrict__ surv,
                                                                 int n, int n_classes, int world, int rank, int* __restrict__ groups) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = labels[i];
  const bool mine = surv[i] && l >= 0 && l < n_classes && (world == 1 || l % world == rank);
  groups[i] = mine ? l : n_classes;
}

constexpr int kSceneParts = 16;      // the index range is cut into 16 parts: counts per (part, class), one CTA per (class, part)

__global__ void __launch_bounds__(256) scene_count_kernel(const int* __restrict__ labels, const uint8_t* __restrict__ kept, int n,
                                                          int n_classes, int part_len, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = (i < n && kept[i]) ? labels[i] : -1;
  if (l < 0 || l >= n_classes) return;
  // neighbouring detections mostly share a class (and a part): one atomic per (warp, part, class)
  const int key = (i / part_len) * n_classes + l;
  const unsigned peers = __match_any_sync(__activemask(), key);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + key, __popc(peers));
}

// class c's survivors in ascending original index, behind those of the classes < c: CTA (c, p) compacts part p of the
// index range for class c; its base comes from the (part, class) counts
__global__ void __launch_bounds__(1024) scene_compact_kernel(const float* __restrict__ scene, int fmt, const float* __restrict__ scores,
                                                             const int* __restrict__ labels, const uint8_t* __restrict__ kept, int n,
                                                             int n_classes, int part_len, const int* __restrict__ counts,
                                                             float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                             int* __restrict__ out_labels, int* __restrict__ out_index,
                                                             int* __restrict__ n_out) {
  __shared__ int s_warp[32];
  __shared__ int s_red[32];
  const int c = blockIdx.x, p = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // base = every kept box of a lower class + those of class c in earlier parts (total = all, for n_out)
  int mine = 0, all = 0;
  for (int k = tid; k < kSceneParts * n_classes; k += 1024) {
    const int v = counts[k], kc = k % n_classes, kp = k / n_classes;
    all += v;
    if (kc < c || (kc == c && kp < p)) mine += v;
  }
  mine = __reduce_add_sync(0xffffffffu, mine); all = __reduce_add_sync(0xffffffffu, all);
  if (lane == 0) { s_warp[warp] = mine; s_red[warp] = all; }
  __syncthreads();
  int base = 0, total = 0;
  for (int w = 0; w < 32; ++w) { base += s_warp[w]; total += s_red[w]; }
  if (c == 0 && p == 0 && tid == 0) *n_out = total;
  __syncthreads();
  const int lo = p * part_len, hi = min(n, lo + part_len);
  for (int i0 = lo; i0 < hi; i0 += 1024) {
    const int i = i0 + tid;
    const bool f = i < hi && kept[i] && labels[i] == c;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, sum = 0;
    for (int w = 0; w < 32; ++w) { const int v = s_warp[w]; if (w < warp) before += v; sum += v; }
    if (f) {
      const int o = base + before + __popc(bal & ((1u << lane) - 1u));
      for (int k = 0; k < fmt; ++k) out_boxes[(size_t)o * fmt + k] = scene[(size_t)i * fmt + k];
      out_scores[o] = scores[i];
      out_labels[o] = c;
      out_index[o] = i;
    }
    base += sum;
    __syncthreads();
  }
}

struct SceneLayout { size_t nms, groups, small, total; };
// n_tile_groups = tiles x classes (stage 1), n_classes (stage 2): the NMS area holds the larger of the two layouts -- not
// always stage 1's: with <= 8192 boxes the fused kernel pads every group's mask to 32 rows, but only up to 1024 groups
static int scene_layout(int n, int n_tile_groups, int n_classes, int fmt, SceneLayout* S, NmsLayout* L, int n_groups) {
  size_t cub_bytes = 0;
  if (int rc = cub_temp_bytes(max(n, 1), 64, &cub_bytes)) return rc;
  const size_t nms_bytes = std::max(nms_layout(max(n, 1), n_tile_groups, fmt, cub_bytes).total,
                                    nms_layout(max(n, 1), n_classes, fmt, cub_bytes).total);
  if (L) *L = nms_layout(max(n, 1), n_groups, fmt, cub_bytes);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 128); return o; };
  S->nms = take(nms_bytes); S->groups = take((size_t)n * 4); S->small = take(4096 + 8);
  S->total = off;
  return AIDET_OK;
}

template <class F>
static int scene_dispatch(int fmt, F&& f) {
  if (fmt == 5) return f(NmsRect{});
  if (fmt == 8) return f(NmsQuad{});
  return f(NmsHbb{});
}
}  // namespace aidet

using namespace aidet;

extern "C" {

size_t aidet_nms_workspace_bytes(int n, int n_groups, int fmt) {
  if (n <= 0) return 256;
  if (n_groups < 1) n_groups = 1;
  size_t cub_bytes = 0;
  if (cub_temp_bytes(n, 64, &cub_bytes) != AIDET_OK) return 0;
  return nms_layout(n, n_groups, fmt, cub_bytes).total;
}

int aidet_nms_batched_f32(const float* boxes, int fmt, const float* scores, const int* group_ids, int n,
                          const float* thr, int n_thr, int n_groups, int cmp, int plus_one, long long* keep_out,
                          int* n_keep, void* workspace, size_t ws_bytes, int device, void* stream) {
  AIDET_REQUIRE(fmt == 4 || fmt == 5 || fmt == 8, "aidet_nms_batched_f32: fmt must be 4, 5 or 8, got %d", fmt);
  AIDET_REQUIRE(n >= 0 && n_groups >= 1, "aidet_nms_batched_f32: bad sizes n=%d n_groups=%d", n, n_groups);
  AIDET_REQUIRE(n_thr == 1 || n_thr == n_groups, "aidet_nms_batched_f32: n_thr must be 1 or n_groups");
  AIDET_REQUIRE(cmp == AIDET_CMP_GT || cmp == AIDET_CMP_GE, "aidet_nms_batched_f32: bad cmp %d", cmp);
  AIDET_REQUIRE(n_keep && thr, "aidet_nms_batched_f32: null pointer");
  AIDET_REQUIRE(group_ids || n_groups == 1, "aidet_nms_batched_f32: group_ids required when n_groups > 1");
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) { AIDET_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int), s)); return AIDET_OK; }
  AIDET_REQUIRE(boxes && scores && keep_out && workspace, "aidet_nms_batched_f32: null pointer");
  AIDET_REQUIRE(((uintptr_t)workspace & 127) == 0, "aidet_nms_batched_f32: workspace must be 128 B aligned");
  size_t cub_bytes = 0;
  if (int rc = cub_temp_bytes(n, 64, &cub_bytes)) return rc;
  NmsLayout L = nms_layout(n, n_groups, fmt, cub_bytes);
  if (ws_bytes < L.total) {
    set_error("aidet_nms_batched_f32: workspace %zu < %zu", ws_bytes, L.total);
    return AIDET_EWORKSPACE;
  }
  char* ws = (char*)workspace;
  const float one = plus_one ? 1.0f : 0.0f;
  if (fmt == 5) return run_nms<NmsRect>(boxes, scores, group_ids, n, thr, n_thr, n_groups, cmp, one, keep_out, n_keep, ws, L, device, s);
  if (fmt == 8) return run_nms<NmsQuad>(boxes, scores, group_ids, n, thr, n_thr, n_groups, cmp, one, keep_out, n_keep, ws, L, device, s);
  return run_nms<NmsHbb>(boxes, scores, group_ids, n, thr, n_thr, n_groups, cmp, one, keep_out, n_keep, ws, L, device, s);
}


size_t aidet_scene_workspace_bytes(int n, int n_tiles, int n_classes, int fmt) {
  if (n < 0 || n_tiles < 1 || n_classes < 1 || n_classes > 1024 || (fmt != 4 && fmt != 5 && fmt != 8)) return 0;
  if ((long long)n_tiles * n_classes > (1LL << 24)) return 0;
  aidet::SceneLayout S;
  if (aidet::scene_layout(n, n_tiles * n_classes, n_classes, fmt, &S, nullptr, 1)) return 0;
  return S.total;
}

int aidet_scene_tile_nms_f32(const float* boxes, int fmt, const float* scores, const int* labels, const int* tile_ids,
                             const float* tile_origins, int n, int n_tiles, int n_classes, const float* thr, int world,
                             int rank, unsigned char* keep_mask, float* scene_boxes, void* workspace, size_t ws_bytes,
                             int device, void* stream) {
  using namespace aidet;
  AIDET_REQUIRE(fmt == 4 || fmt == 5 || fmt == 8, "aidet_scene_tile_nms_f32: fmt must be 4, 5 or 8, got %d", fmt);
  AIDET_REQUIRE(n >= 0 && n_tiles >= 1 && n_classes >= 1 && n_classes <= 1024, "aidet_scene_tile_nms_f32: bad sizes");
  AIDET_REQUIRE(world >= 1 && rank >= 0 && rank < world, "aidet_scene_tile_nms_f32: bad world %d / rank %d", world, rank);
  if (int rc = set_device(device)) return rc;
  if (n == 0) return AIDET_OK;
  AIDET_REQUIRE(boxes && scores && labels && tile_ids && tile_origins && thr && keep_mask && scene_boxes && workspace,
                "aidet_scene_tile_nms_f32: null pointer");
  AIDET_REQUIRE(((uintptr_t)workspace & 127) == 0, "aidet_scene_tile_nms_f32: workspace must be 128 B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int n_groups = n_tiles * n_classes;
  SceneLayout S; NmsLayout L;
  if (int rc = scene_layout(n, n_groups, n_classes, fmt, &S, &L, n_groups)) return rc;
  if (ws_bytes < S.total) { set_error("aidet_scene_tile_nms_f32: workspace %zu < %zu", ws_bytes, S.total); return AIDET_EWORKSPACE; }
  char* ws = (char*)workspace;
  int* groups = (int*)(ws + S.groups);
  int* n_keep = (int*)(ws + S.small);
  scene_prepare_kernel<<<ceil_div(n, 256), 256, 0, s>>>(boxes, fmt, labels, tile_ids, tile_origins, n, n_tiles, n_classes, world,
                                                        rank, groups, scene_boxes);
  count_launch(1);
  int rc = scene_dispatch(fmt, [&](auto o) {
    return run_nms<decltype(o)>(boxes, scores, groups, n, thr, 1, n_groups, AIDET_CMP_GT, 0.0f, nullptr, n_keep, ws + S.nms, L, device, s);
  });
  if (rc) return rc;
  AIDET_CUDA(cudaMemcpyAsync(keep_mask, ws + S.nms + L.flags, (size_t)n, cudaMemcpyDeviceToDevice, s));
  return AIDET_OK;
}

int aidet_scene_merge_nms_f32(const float* scene_boxes, int fmt, const float* scores, const int* labels,
                              const unsigned char* survivors, int n, int n_tiles, int n_classes, const float* merge_thr,
                              int world, int rank, unsigned char* keep_mask, void* workspace, size_t ws_bytes, int device,
                              void* stream) {
  using namespace aidet;
  AIDET_REQUIRE(fmt == 4 || fmt == 5 || fmt == 8, "aidet_scene_merge_nms_f32: fmt must be 4, 5 or 8, got %d", fmt);
  AIDET_REQUIRE(n >= 0 && n_tiles >= 1 && n_classes >= 1 && n_classes <= 1024, "aidet_scene_merge_nms_f32: bad sizes");
  AIDET_REQUIRE(world >= 1 && rank >= 0 && rank < world, "aidet_scene_merge_nms_f32: bad world %d / rank %d", world, rank);
  if (int rc = set_device(device)) return rc;
  if (n == 0) return AIDET_OK;
  AIDET_REQUIRE(scene_boxes && scores && labels && survivors && merge_thr && keep_mask && workspace,
                "aidet_scene_merge_nms_f32: null pointer");
  AIDET_REQUIRE(((uintptr_t)workspace & 127) == 0, "aidet_scene_merge_nms_f32: workspace must be 128 B aligned");
  cudaStream_t s = (cudaStream_t)stream;
  SceneLayout S; NmsLayout L;
  if (int rc = scene_layout(n, n_tiles * n_classes, n_classes, fmt, &S, &L, n_classes)) return rc;
  if (ws_bytes < S.total) { set_error("aidet_scene_merge_nms_f32: workspace %zu < %zu", ws_bytes, S.total); return AIDET_EWORKSPACE; }
  char* ws = (char*)workspace;
  int* groups = (int*)(ws + S.groups);
  int* n_keep = (int*)(ws + S.small);
  scene_merge_groups_kernel<<<ceil_div(n, 256), 256, 0, s>>>(labels, survivors, n, n_classes, world, rank, groups);
  count_launch(1);
  int rc = scene_dispatch(fmt, [&](auto o) {
    return run_nms<decltype(o)>(scene_boxes, scores, groups, n, merge_thr, n_classes, n_classes, AIDET_CMP_GT, 0.0f, nullptr, n_keep,
                                ws + S.nms, L, device, s);
  });
  if (rc) return rc;
  AIDET_CUDA(cudaMemcpyAsync(keep_mask, ws + S.nms + L.flags, (size_t)n, cudaMemcpyDeviceToDevice, s));
  return AIDET_OK;
}

int aidet_scene_compact_f32(const float* scene_boxes, int fmt, const float* scores, const int* labels,
                            const unsigned char* kept, int n, int n_classes, float* out_boxes, float* out_scores,
                            int* out_labels, int* out_index, int* n_out, void* workspace, size_t ws_bytes, int device,
                            void* stream) {
  using namespace aidet;
  AIDET_REQUIRE(fmt == 4 || fmt == 5 || fmt == 8, "aidet_scene_compact_f32: fmt must be 4, 5 or 8, got %d", fmt);
  AIDET_REQUIRE(n >= 0 && n_classes >= 1 && n_classes <= 1024, "aidet_scene_compact_f32: bad sizes");
  AIDET_REQUIRE(n_out && workspace && ws_bytes >= (size_t)kSceneParts * n_classes * sizeof(int),
                "aidet_scene_compact_f32: null pointer / workspace too small");
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) { AIDET_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int), s)); return AIDET_OK; }
  AIDET_REQUIRE(scene_boxes && scores && labels && kept && out_boxes && out_scores && out_labels && out_index,
                "aidet_scene_compact_f32: null pointer");
  int* counts = (int*)workspace;
  const int part_len = ceil_div(n, kSceneParts);
  AIDET_CUDA(cudaMemsetAsync(counts, 0, (size_t)kSceneParts * n_classes * sizeof(int), s));
  scene_count_kernel<<<ceil_div(n, 256), 256, 0, s>>>(labels, kept, n, n_classes, part_len, counts);
  scene_compact_kernel<<<dim3(n_classes, kSceneParts), 1024, 0, s>>>(scene_boxes, fmt, scores, labels, kept, n, n_classes, part_len,
                                                                     counts, out_boxes, out_scores, out_labels, out_index, n_out);
  count_launch(2);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

}  // extern "C"
