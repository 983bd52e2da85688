// riou_assign.cu -- max-IoU assignment fused with the overlap computation, on sm_100a.
//
// Replaces (reference): MaxIoUAssigner.assign / assign_wrt_overlaps
// (mmdet/core/bbox/assigners/max_iou_assigner.py:52-195), which materialises the (k, n) overlap matrix
// (bbox_overlaps, :102), takes max/argmax along both axes (:155-158) and then loops over the k ground truths in
// Python, each iteration an (n,) comparison + masked store (:176-182).  Here the matrix is never written:
//
//   pass 1 : tiles of <= 64 gt rows x 256 boxes, the same schedule as the overlap-matrix kernel (gt records staged
//            into shared memory by one 1-D TMA bulk copy, one box record per lane in registers).  Each lane keeps the
//            running (max, first argmax) of its column; row maxima are reduced with redux.sync per warp, then in
//            shared memory per CTA, then one atomicMax per (CTA, row).  Column results of different row tiles are
//            combined with a 64-bit atomicMax on (overlap bits, ~row) so ties resolve to the first gt (:155).
//   pass 2 : the overlaps are RECOMPUTED by the SAME kernel (the pass is a run-time argument, so both passes run the
//            very same SASS and produce identical bits) and compared with the row maxima:
//            per box the last gt i with overlaps[i, j] == gt_max[i] >= min_pos_iou (:176-180, later gts overwrite
//            earlier ones), per gt the first such box (:182 when gt_max_assign_all is false).
//   finish : steps 1-4 of :136-182 per box, labels gathered (:184-190).
//
// The ignore logic (:104-113: boxes whose IoF with any gt_bboxes_ignore exceeds ignore_iof_thr are set to -1) is
// pass 1 run on the ignore boxes in IoF mode; passes 1/2 then treat those columns as -1.
// Bound: FP32 issue like the matrix kernel, but with 2x the arithmetic and no (k, n) store or re-reads.
#include <limits.h>

#include "common.cuh"
#include "geom.cuh"
#include "pairop.cuh"

namespace aidet {

constexpr int kACols = 256;
constexpr int kARows = 64;

// Overlaps are >= 0 (or -1 = "ignored" in a caller-provided matrix): v > 0 -> bits + 1, +-0 -> 1, anything else
// (negative, NaN) -> 0, so unsigned order == float order and 0 means "no value" (the reference's -1).  Both zeros
// map to the same code: the bits of -0.0f (0x80000000) would otherwise rank above every positive overlap.
__host__ __device__ __forceinline__ unsigned ov_enc(float v) {
#if defined(__CUDA_ARCH__)
  return v > 0.0f ? __float_as_uint(v) + 1u : (v == 0.0f ? 1u : 0u);
#else
  if (v == 0.0f) return 1u;
  if (!(v > 0.0f)) return 0u;
  union { float f; unsigned u; } x; x.f = v; return x.u + 1u;
#endif
}
__device__ __forceinline__ float ov_dec(unsigned e) { return e ? __uint_as_float(e - 1u) : -1.0f; }

template <class K, int MODE, bool FROM_MATRIX>
__global__ void __launch_bounds__(kACols)
assign_pass_kernel(const typename PairOp<K>::S* __restrict__ rows, int m,
                   const typename PairOp<K>::R* __restrict__ cols, int n,
                   const float* __restrict__ ov_mat, long long ld, int tile_rows, int PASS,
                   const unsigned long long* __restrict__ ign_best, unsigned ign_thr_enc,
                   unsigned long long* __restrict__ col_best, unsigned* __restrict__ row_e,
                   float min_pos, int* __restrict__ col_last, int* __restrict__ row_first) {
  using P = PairOp<K>;
  using S = typename P::S; using R = typename P::R;
  __shared__ __align__(128) S stage[kARows];
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned s_row[kARows];   // pass 1: row maxima of this CTA; pass 2: the value a hit must equal (0 = none)
  __shared__ int s_first[kARows];

  const int r0 = blockIdx.y * tile_rows;
  const int nr = min(tile_rows, m - r0);
  const int col = blockIdx.x * kACols + threadIdx.x;
  const bool live = col < n;

  if (!FROM_MATRIX && threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < kARows) {
    if (PASS == 1) {
      s_row[threadIdx.x] = 0u;
    } else {
      unsigned e = (int)threadIdx.x < nr ? row_e[r0 + threadIdx.x] : 0u;
      s_row[threadIdx.x] = (e && ov_dec(e) >= min_pos) ? e : 0u;     // max_iou_assigner.py:177
      s_first[threadIdx.x] = INT_MAX;
    }
  }
  __syncthreads();
  if (!FROM_MATRIX && threadIdx.x == 0) {
    uint32_t bytes = (uint32_t)(nr * (int)sizeof(S));
    mbar_expect_tx(&bar, bytes);
    tma_load_1d(&stage[0], rows + r0, bytes, &bar);
  }
  typename P::X me;
  if (!FROM_MATRIX) me = cols[live ? col : n - 1];
  const bool ignored = live && ign_best && (unsigned)(ign_best[col] >> 32) > ign_thr_enc;
  const bool act = live && !ignored;
  if (!FROM_MATRIX) mbar_wait(&bar, 0);

  unsigned best_e = 0u; int best_i = 0;      // pass 1
  int last = -1;                              // pass 2
  const float* mp = FROM_MATRIX ? ov_mat + (long long)r0 * ld + col : nullptr;
#pragma unroll 2
  for (int r = 0; r < nr; ++r) {
    float v;
    if (FROM_MATRIX) { v = live ? __ldg(mp) : -1.0f; mp += ld; }
    else v = P::overlap(stage[r], me, MODE);
    const unsigned e = act ? ov_enc(v) : 0u;
    if (PASS == 1) {
      if (e > best_e) { best_e = e; best_i = r0 + r; }             // strict: the first gt wins ties
      const unsigned w = __reduce_max_sync(0xffffffffu, e);
      if ((threadIdx.x & 31) == 0 && w) atomicMax(&s_row[r], w);
    } else {
      const unsigned tgt = s_row[r];
      if (tgt != 0u && e == tgt) { last = r0 + r; atomicMin(&s_first[r], col); }
    }
  }
  if (PASS == 1) {
    if (best_e) atomicMax(&col_best[col], ((unsigned long long)best_e << 32) | (unsigned long long)(0xffffffffu - (unsigned)best_i));
    __syncthreads();
    if ((int)threadIdx.x < nr && s_row[threadIdx.x]) atomicMax(&row_e[r0 + threadIdx.x], s_row[threadIdx.x]);
  } else {
    if (last >= 0) atomicMax(&col_last[col], last);
    __syncthreads();
    if ((int)threadIdx.x < nr && s_first[threadIdx.x] != INT_MAX) atomicMin(&row_first[r0 + threadIdx.x], s_first[threadIdx.x]);
  }
}

// gt_max_assign_all == false: assigned_gt_inds[gt_argmax_overlaps[i]] = i + 1 in ascending i (max_iou_assigner.py:182)
__global__ void __launch_bounds__(256) assign_scatter_kernel(const int* __restrict__ row_first, int m, int n,
                                                             int* __restrict__ col_sel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int j = row_first[i];
  if (j >= 0 && j < n) atomicMax(&col_sel[j], i);
}

// steps 1-4 of max_iou_assigner.py:136-182 and the label gather of :184-190, one thread per box
__global__ void __launch_bounds__(256)
assign_finish_kernel(int n, const unsigned long long* __restrict__ col_best,
                     const unsigned long long* __restrict__ ign_best, unsigned ign_thr_enc, float pos_thr,
                     float neg_lo, float neg_hi, const int* __restrict__ col_pick,
                     const long long* __restrict__ gt_labels, long long* __restrict__ gt_inds,
                     float* __restrict__ max_ov, long long* __restrict__ labels) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const bool ignored = ign_best && (unsigned)(ign_best[j] >> 32) > ign_thr_enc;
  const unsigned long long key = ignored ? 0ull : col_best[j];
  const float mv = ov_dec((unsigned)(key >> 32));
  const int arg = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : 0;
  long long g = -1;                                   // 1. don't care
  if (mv >= neg_lo && mv < neg_hi) g = 0;             // 2. negative
  if (mv >= pos_thr) g = arg + 1;                     // 3. positive
  const int pk = col_pick[j];
  if (pk >= 0) g = pk + 1;                            // 4. the boxes nearest to each gt
  gt_inds[j] = g;
  max_ov[j] = mv;
  if (labels) labels[j] = (g > 0 && gt_labels) ? gt_labels[g - 1] : 0;
}

struct AssignWs {
  char* rows; char* ign_rows; char* cols;
  unsigned long long* col_best; unsigned long long* ign_best;
  unsigned* row_e; unsigned* ign_row_e;
  int* col_last; int* row_first; int* col_sel;
  size_t total;
};

static AssignWs assign_layout(void* base, int m, int n, int k_ign, int fmt) {
  AssignWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes ? bytes : 1, 128); return (char*)base + o; };
  const size_t rec = fmt ? record_bytes(fmt) : 0;
  w.rows = take((size_t)m * rec);
  w.ign_rows = take((size_t)k_ign * rec);
  w.cols = take((size_t)n * rec);
  w.col_best = (unsigned long long*)take((size_t)n * 8);
  w.ign_best = (unsigned long long*)take(k_ign ? (size_t)n * 8 : 0);
  w.row_e = (unsigned*)take((size_t)m * 4);
  w.ign_row_e = (unsigned*)take((size_t)k_ign * 4);
  w.col_last = (int*)take((size_t)n * 4);
  w.row_first = (int*)take((size_t)m * 4);
  w.col_sel = (int*)take((size_t)n * 4);
  w.total = off;
  return w;
}

static int pick_tile_rows(int m, int n, int device) {
  const int sms = sm_count(device);
  const int n_col_tiles = ceil_div(n, kACols);
  int tile_rows = kARows;
  while (tile_rows > 8 && (long long)ceil_div(m, tile_rows) * n_col_tiles < 2LL * sms) tile_rows >>= 1;
  return tile_rows;
}

struct AssignParams {
  float pos_thr, neg_lo, neg_hi, min_pos; int assign_all;
  const long long* gt_labels; long long* gt_inds; float* max_ov; long long* labels;
};

template <class K, int MODE>
static void launch_pass(int pass, const void* rows, int m, const void* cols, int n, int device, cudaStream_t s,
                        const unsigned long long* ign_best, unsigned ign_thr_enc, unsigned long long* col_best,
                        unsigned* row_e, float min_pos, int* col_last, int* row_first) {
  using P = PairOp<K>;
  const int tile_rows = pick_tile_rows(m, n, device);
  dim3 grid(ceil_div(n, kACols), ceil_div(m, tile_rows));
  assign_pass_kernel<K, MODE, false><<<grid, kACols, 0, s>>>(
      (const typename P::S*)rows, m, (const typename P::R*)cols, n, nullptr, 0, tile_rows, pass, ign_best, ign_thr_enc,
      col_best, row_e, min_pos, col_last, row_first);
  count_launch(1);
}

static int assign_tail(const AssignWs& w, int m, int n, const unsigned long long* ign_best, unsigned ign_thr_enc,
                       const AssignParams& p, cudaStream_t s) {
  const int* pick = w.col_last;
  if (!p.assign_all) {
    assign_scatter_kernel<<<ceil_div(m, 256), 256, 0, s>>>(w.row_first, m, n, w.col_sel);
    count_launch(1);
    pick = w.col_sel;
  }
  assign_finish_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, w.col_best, ign_best, ign_thr_enc, p.pos_thr, p.neg_lo,
                                                        p.neg_hi, pick, p.gt_labels, p.gt_inds, p.max_ov, p.labels);
  count_launch(1);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

// One launch: prepared records of the truths (rows), the ignore boxes (rows) and the candidates (cols), plus the
// initial values of every accumulator the passes combine into with atomics.
template <class K>
__global__ void __launch_bounds__(256)
assign_prepare_kernel(const float* __restrict__ gts, int m, const float* __restrict__ ign, int k_ign,
                      const float* __restrict__ boxes, int n, typename K::Row* rows, typename K::Row* ign_rows,
                      typename K::Col* cols, AssignWs w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float b[K::FMT];
  if (i < m) {
#pragma unroll
    for (int k = 0; k < K::FMT; k++) b[k] = gts[(size_t)i * K::FMT + k];
    typename K::Row r; K::prepare(b, &r, nullptr);
    rows[i] = r;
    w.row_e[i] = 0u; w.row_first[i] = INT_MAX;
    return;
  }
  i -= m;
  if (i < k_ign) {
#pragma unroll
    for (int k = 0; k < K::FMT; k++) b[k] = ign[(size_t)i * K::FMT + k];
    typename K::Row r; K::prepare(b, &r, nullptr);
    ign_rows[i] = r;
    w.ign_row_e[i] = 0u;
    return;
  }
  i -= k_ign;
  if (i < n) {
#pragma unroll
    for (int k = 0; k < K::FMT; k++) b[k] = boxes[(size_t)i * K::FMT + k];
    typename K::Col c; K::prepare(b, nullptr, &c);
    cols[i] = c;
    w.col_best[i] = 0ull; w.col_last[i] = -1; w.col_sel[i] = -1;
    if (k_ign) w.ign_best[i] = 0ull;
  }
}

__global__ void __launch_bounds__(256) assign_init_kernel(int m, int n, AssignWs w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) { w.row_e[i] = 0u; w.row_first[i] = INT_MAX; }
  if (i < n) { w.col_best[i] = 0ull; w.col_last[i] = -1; w.col_sel[i] = -1; }
}

template <class K>
static int assign_fused(const float* gts, int m, const float* boxes, int n, const float* gt_ignore, int k_ign,
                        float ign_thr, int wrt_candidates, const AssignParams& p, void* ws, int device,
                        cudaStream_t s) {
  using P = PairOp<K>;
  AssignWs w = assign_layout(ws, m, n, k_ign, K::FMT);
  assign_prepare_kernel<K><<<ceil_div(m + k_ign + n, 256), 256, 0, s>>>(
      gts, m, gt_ignore, k_ign, boxes, n, (typename K::Row*)w.rows, (typename K::Row*)w.ign_rows, (typename K::Col*)w.cols, w);
  count_launch(1);
  const unsigned long long* ign_best = nullptr;
  unsigned ign_thr_enc = 0;
  if (k_ign) {
    // iof(bboxes, gt_ignore) divides by the box (column) area, iof(gt_ignore, bboxes) by the ignore (row) area
    if (wrt_candidates)
      launch_pass<K, MODE_IOF_B>(1, w.ign_rows, k_ign, w.cols, n, device, s, nullptr, 0, w.ign_best, w.ign_row_e, 0.f,
                                    nullptr, nullptr);
    else
      launch_pass<K, MODE_IOF>(1, w.ign_rows, k_ign, w.cols, n, device, s, nullptr, 0, w.ign_best, w.ign_row_e, 0.f,
                                  nullptr, nullptr);
    ign_best = w.ign_best;
    ign_thr_enc = ov_enc(ign_thr);
  }
  launch_pass<K, MODE_IOU>(1, w.rows, m, w.cols, n, device, s, ign_best, ign_thr_enc, w.col_best, w.row_e, 0.f, nullptr,
                              nullptr);
  launch_pass<K, MODE_IOU>(2, w.rows, m, w.cols, n, device, s, ign_best, ign_thr_enc, nullptr, w.row_e, p.min_pos,
                              w.col_last, w.row_first);
  (void)sizeof(P);
  return assign_tail(w, m, n, ign_best, ign_thr_enc, p, s);
}

}  // namespace aidet

using namespace aidet;

extern "C" {

size_t aidet_assign_workspace_bytes(int m, int n, int k_ign, int fmt) {
  if (m < 0) m = 0;
  if (n < 0) n = 0;
  if (k_ign < 0) k_ign = 0;
  return assign_layout(nullptr, m, n, k_ign, fmt).total + 128;
}

int aidet_max_iou_assign_f32(const float* gts, int m, const float* bboxes, int n, int fmt, const float* gt_ignore,
                             int k_ign, float ignore_iof_thr, int ignore_wrt_candidates, float pos_iou_thr,
                             float neg_lo, float neg_hi, float min_pos_iou, int gt_max_assign_all,
                             const long long* gt_labels, long long* gt_inds, float* max_overlaps, long long* labels,
                             void* workspace, size_t ws_bytes, int device, void* stream) {
  AIDET_REQUIRE(fmt == 4 || fmt == 5 || fmt == 8, "aidet_max_iou_assign_f32: fmt must be 4, 5 or 8, got %d", fmt);
  AIDET_REQUIRE(m > 0 && n > 0, "aidet_max_iou_assign_f32: needs at least one gt and one box (m %d, n %d)", m, n);
  AIDET_REQUIRE(k_ign >= 0, "aidet_max_iou_assign_f32: negative k_ign");
  AIDET_REQUIRE(gts && bboxes && gt_inds && max_overlaps && workspace, "aidet_max_iou_assign_f32: null pointer");
  AIDET_REQUIRE(k_ign == 0 || (gt_ignore && ignore_iof_thr > 0.0f),
                "aidet_max_iou_assign_f32: ignore boxes need a pointer and ignore_iof_thr > 0");
  AIDET_REQUIRE(min_pos_iou > -1.0f, "aidet_max_iou_assign_f32: min_pos_iou must be > -1");
  AIDET_REQUIRE(!labels || gt_labels, "aidet_max_iou_assign_f32: labels requested without gt_labels");
  AIDET_REQUIRE(((uintptr_t)workspace & 15) == 0, "aidet_max_iou_assign_f32: workspace must be 16 B aligned");
  if (ws_bytes < aidet_assign_workspace_bytes(m, n, k_ign, fmt)) {
    set_error("aidet_max_iou_assign_f32: workspace %zu < %zu", ws_bytes, aidet_assign_workspace_bytes(m, n, k_ign, fmt));
    return AIDET_EWORKSPACE;
  }
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  void* ws = (void*)align_up((size_t)(uintptr_t)workspace, 128);
  AssignParams p{pos_iou_thr, neg_lo, neg_hi, min_pos_iou, gt_max_assign_all, gt_labels, gt_inds, max_overlaps, labels};
  if (fmt == 5) return assign_fused<RectKind>(gts, m, bboxes, n, gt_ignore, k_ign, ignore_iof_thr, ignore_wrt_candidates, p, ws, device, s);
  if (fmt == 4) return assign_fused<HbbKind>(gts, m, bboxes, n, gt_ignore, k_ign, ignore_iof_thr, ignore_wrt_candidates, p, ws, device, s);
  return assign_fused<QuadKind>(gts, m, bboxes, n, gt_ignore, k_ign, ignore_iof_thr, ignore_wrt_candidates, p, ws, device, s);
}

int aidet_assign_wrt_overlaps_f32(const float* overlaps, int m, int n, long long ld, float pos_iou_thr, float neg_lo,
                                  float neg_hi, float min_pos_iou, int gt_max_assign_all, const long long* gt_labels,
                                  long long* gt_inds, float* max_overlaps, long long* labels, void* workspace,
                                  size_t ws_bytes, int device, void* stream) {
  AIDET_REQUIRE(m > 0 && n > 0, "aidet_assign_wrt_overlaps_f32: needs at least one gt and one box (m %d, n %d)", m, n);
  AIDET_REQUIRE(overlaps && gt_inds && max_overlaps && workspace, "aidet_assign_wrt_overlaps_f32: null pointer");
  AIDET_REQUIRE(ld >= n, "aidet_assign_wrt_overlaps_f32: ld %lld < n %d", ld, n);
  AIDET_REQUIRE(min_pos_iou > -1.0f, "aidet_assign_wrt_overlaps_f32: min_pos_iou must be > -1");
  AIDET_REQUIRE(!labels || gt_labels, "aidet_assign_wrt_overlaps_f32: labels requested without gt_labels");
  if (ws_bytes < aidet_assign_workspace_bytes(m, n, 0, 0)) {
    set_error("aidet_assign_wrt_overlaps_f32: workspace %zu < %zu", ws_bytes, aidet_assign_workspace_bytes(m, n, 0, 0));
    return AIDET_EWORKSPACE;
  }
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  void* ws = (void*)align_up((size_t)(uintptr_t)workspace, 128);
  AssignWs w = assign_layout(ws, m, n, 0, 0);
  assign_init_kernel<<<ceil_div(m > n ? m : n, 256), 256, 0, s>>>(m, n, w);
  count_launch(1);
  const int tile_rows = pick_tile_rows(m, n, device);
  dim3 grid(ceil_div(n, kACols), ceil_div(m, tile_rows));
  assign_pass_kernel<HbbKind, MODE_IOU, true><<<grid, kACols, 0, s>>>(nullptr, m, nullptr, n, overlaps, ld, tile_rows, 1,
                                                                          nullptr, 0, w.col_best, w.row_e, 0.f, nullptr, nullptr);
  assign_pass_kernel<HbbKind, MODE_IOU, true><<<grid, kACols, 0, s>>>(nullptr, m, nullptr, n, overlaps, ld, tile_rows, 2,
                                                                          nullptr, 0, nullptr, w.row_e, min_pos_iou,
                                                                          w.col_last, w.row_first);
  count_launch(2);
  AssignParams p{pos_iou_thr, neg_lo, neg_hi, min_pos_iou, gt_max_assign_all, gt_labels, gt_inds, max_overlaps, labels};
  return assign_tail(w, m, n, nullptr, 0, p, s);
}

}  // extern "C"
