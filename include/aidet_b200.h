/*
 * aidet_b200.h -- C ABI of libaidet_b200.so: B200 (sm_100a) oriented-bounding-box ops
 * behind AIDet's mmdet.ops interface.
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless
 *     the name ends in _host; the caller owns every buffer (including workspaces)
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the
 *     call returns without synchronising unless stated
 *   - returns 0 on success, a negative AIDET_E* code otherwise; the message is
 *     available from aidet_last_error() (thread local).  Nothing throws, exits or
 *     prints (the reference's pybind layer raises from AT_CHECK/AT_ASSERTM --
 *     mmdet/ops/nms/src/nms_cuda.cpp:4,9, mmdet/ops/roi_align/src/roi_align_cuda.cpp:37-42
 *     -- and its v1 RoIAlign printf()s / exit(-1)s, roi_align_cuda.cpp:56-59,
 *     roi_align_kernel.cu:269-272; the Python shim turns codes into RuntimeError)
 *   - box formats: 4 = HBB (x1,y1,x2,y2), 5 = theta-OBB (cx,cy,w,h,theta[rad]),
 *     8 = point-OBB (x1,y1,...,x4,y4)   (mmdet/core/rbbox/transforms.py:45-55)
 *
 * The reference-side binding for each function is shown in INTEGRATION.md.
 */
#ifndef AIDET_B200_H_
#define AIDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AIDET_OK 0
#define AIDET_EINVAL (-1)   /* bad argument */
#define AIDET_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define AIDET_EWORKSPACE (-3) /* workspace too small */

#define AIDET_MODE_IOU 0
#define AIDET_MODE_IOF 1
#define AIDET_CMP_GT 0      /* suppress when ovr >  thr : mmdet/ops/nms/src/nms_kernel.cu:61 */
#define AIDET_CMP_GE 1      /* suppress when ovr >= thr : mmdet/ops/nms/src/nms_cpu.cpp:56   */

/* ---- library ------------------------------------------------------------ */
const char* aidet_last_error(void);
int aidet_version(void);                       /* 100 * major + minor */
/* Diagnostics (process-wide switch and counters, off by default; the only state of the library besides the
 * thread-local error string).  enable 1: device-side timing of each op (CUDA events recorded on the caller's stream);
 * enable 2: the fused NMS kernel additionally leaves its phase time stamps in the last 256 bytes of its workspace. */
int aidet_prof_enable(int enable);
/* kind: 0 = riou matrix kernel, 1 = batched NMS, the WHOLE call (every kernel it enqueues: the roofline of
 * SURVEY 8d charges the call), 2 = rroi fwd kernel, 3 = rroi bwd (all kernels of the gather form).
 * Synchronises the recorded events, returns accumulated milliseconds and call count since the last reset. */
int aidet_prof_read(int kind, double* ms_total, long long* launches, int reset);
/* Total number of kernel launches issued by this library since load (all ops). */
long long aidet_launch_count(void);
/* FP32 FFMA micro-benchmark (roofline denominator for the ALU-bound kernels):
 * runs `iters` dependent-free FFMA chains on every SM, returns TFLOP/s. */
int aidet_ffma_peak(int device, int iters, double* tflops_out, void* stream);

/* ---- rotated IoU ---------------------------------------------------------
 * Replaces: the polygon IoU AIDet reaches through wwtool (mmdet/datasets/dota.py:23,336)
 * and the (m,n) overlap API of mmdet/core/bbox/geometry.py:4-88 for rotated boxes.
 * a: (m,fmt) row-major, b: (n,fmt); out: m rows of n floats, row stride ld_out
 * (elements).  workspace >= aidet_riou_workspace_bytes(m,n,fmt), 16 B aligned.
 * fmt 4 = axis-aligned boxes with the legacy +1 convention: bbox_overlaps itself
 * (mmdet/core/bbox/geometry.py:57-86), served by the same tiled kernel. */
size_t aidet_riou_workspace_bytes(int m, int n, int fmt);
int aidet_riou_matrix_f32(const float* a, int m, const float* b, int n, int fmt, int mode,
                          float* out, long long ld_out, void* workspace, size_t ws_bytes,
                          int device, void* stream);
/* Row-sharded multi-GPU form: the same m x n block is stored to n_outs (<= 8) destinations -- the local
 * buffer and the matching row block of every peer GPU's buffer (peer / symmetric-memory mappings of this
 * process; HOST array of DEVICE pointers).  The all-gather of the shard results thereby happens from
 * inside the kernel as stores over NVLink, tile by tile, overlapped with the arithmetic.  The caller
 * synchronises the ranks afterwards (all stores are complete when the kernel is). */
int aidet_riou_matrix_multi_f32(const float* a, int m, const float* b, int n, int fmt, int mode,
                                float* const* outs_host, int n_outs, long long ld_out,
                                void* workspace, size_t ws_bytes, int device, void* stream);
/* Same exchange through NVSwitch multicast (NVLS): out_mc is the MULTICAST address of the row block (e.g. torch
 * symmetric memory's multicast_ptr + offset); every element is stored once with multimem.st and the switch
 * replicates it into all GPUs of the multicast group, this one included -- the rank's NVLink carries its block
 * once instead of once per peer.  Needs a multicast-capable allocation; the caller synchronises the ranks. */
int aidet_riou_matrix_mcast_f32(const float* a, int m, const float* b, int n, int fmt, int mode, float* out_mc,
                                long long ld_out, void* workspace, size_t ws_bytes, int device, void* stream);
/* element-wise pairs (is_aligned=True, geometry.py:57-71): out[i] = ovr(a[i], b[i]) */
int aidet_riou_aligned_f32(const float* a, const float* b, int n, int fmt, int mode, float* out,
                           int device, void* stream);

/* Backward of the aligned overlap: grad_a[i] = grad_ov[i] * d ovr(a[i], b[i]) / d (parameters of a[i]), likewise grad_b.
 * fmt 5: (cx,cy,w,h,theta); fmt 8: the corner coordinates of CONVEX quads (either orientation).  grad_ov NULL = ones;
 * ov (n), grad_a (n,fmt), grad_b (n,fmt) may each be NULL.  This is what makes the rotated IoU loss trainable -- the
 * rotated counterpart of iou_loss (mmdet/models/losses/iou_loss.py:10-27), which gets its gradient from autograd
 * through bbox_overlaps. */
int aidet_riou_aligned_grad_f32(const float* a, const float* b, int n, int fmt, int mode, const float* grad_ov,
                                float* ov, float* grad_a, float* grad_b, int device, void* stream);

/* ---- max-IoU assignment fused with the overlap computation -----------------
 * Replaces: MaxIoUAssigner.assign / assign_wrt_overlaps (mmdet/core/bbox/assigners/max_iou_assigner.py:52-195):
 * overlap matrix + max/argmax along both axes + a Python loop over the ground truths.  The (m, n) matrix is never
 * materialised (two passes that recompute the overlaps; see csrc/riou_assign.cu).
 *   gts (m,fmt), bboxes (n,fmt), fmt 4 (HBB, +1 convention = bbox_overlaps), 5 or 8; m, n >= 1
 *   gt_ignore (k_ign,fmt) or NULL/0: boxes whose IoF with any ignore box exceeds ignore_iof_thr (> 0) count as -1
 *     (:104-113); ignore_wrt_candidates != 0 -> IoF over the box's own area, else over the ignore box's
 *   negatives: neg_lo <= max_overlap < neg_hi (:161-167: a float thr is (0, thr); pass neg_lo = +inf for none)
 *   gt_max_assign_all: :178-182
 *   gt_labels (m) int64 or NULL; outputs gt_inds (n) int64 in {-1, 0, 1..m}, max_overlaps (n), labels (n) int64 or NULL
 * Ties along the gt axis resolve to the FIRST gt (the CPU behaviour of Tensor.max(dim=0)). */
size_t aidet_assign_workspace_bytes(int m, int n, int k_ign, int fmt);
int aidet_max_iou_assign_f32(const float* gts, int m, const float* bboxes, int n, int fmt, const float* gt_ignore,
                             int k_ign, float ignore_iof_thr, int ignore_wrt_candidates, float pos_iou_thr,
                             float neg_lo, float neg_hi, float min_pos_iou, int gt_max_assign_all,
                             const long long* gt_labels, long long* gt_inds, float* max_overlaps, long long* labels,
                             void* workspace, size_t ws_bytes, int device, void* stream);
/* Same assignment from a caller-provided (m, n) overlap matrix (row stride ld elements; entries >= 0, or -1 for
 * "ignored"): assign_wrt_overlaps (:122-195).  workspace >= aidet_assign_workspace_bytes(m, n, 0, 0). */
int aidet_assign_wrt_overlaps_f32(const float* overlaps, int m, int n, long long ld, float pos_iou_thr, float neg_lo,
                                  float neg_hi, float min_pos_iou, int gt_max_assign_all, const long long* gt_labels,
                                  long long* gt_inds, float* max_overlaps, long long* labels, void* workspace,
                                  size_t ws_bytes, int device, void* stream);

/* ---- batched NMS (rotated and axis-aligned) -------------------------------
 * Replaces: nms_cuda.nms (mmdet/ops/nms/src/nms_kernel.cu:71-139: sort, 64x64 bitmask
 * tiles, D2H mask copy, host scan) and the per-class Python loops of
 * mmdet/core/post_processing/rbbox_nms.py:29-49,83-106 with one device-side pass:
 * sort by (group, -score, index) -> upper-triangle suppression bitmask (64-bit words)
 * -> on-device greedy scan per group -> compaction.
 *   boxes (n,fmt) fmt in {4,5,8}; scores (n); group_ids (n) in [0,n_groups) or NULL;
 *   thr: n_thr == 1 (shared) or n_thr == n_groups (per group, dota.py:324);
 *   plus_one: legacy +1 pixel convention, fmt 4 only (nms_kernel.cu:17-21);
 *   keep_out (n) int64, ascending ORIGINAL index (nms_kernel.cu:135-138);
 *   n_keep: int32 scalar the device can write: device memory, or pinned host memory that the host polls
 *           (the kernels only ever store to it, once, when the list is complete).                       */
size_t aidet_nms_workspace_bytes(int n, int n_groups, int fmt);
int aidet_nms_batched_f32(const float* boxes, int fmt, const float* scores, const int* group_ids,
                          int n, const float* thr, int n_thr, int n_groups, int cmp, int plus_one,
                          long long* keep_out, int* n_keep, void* workspace, size_t ws_bytes,
                          int device, void* stream);

/* ---- scene merge: per-tile NMS + cross-tile merge NMS of one large scene ------
 * Replaces: the host loop that merges tile results per class (tools/parse_results.py:56-76, which calls
 * mmdet/datasets/dota.py:296-327 -> DOTA_devkit py_cpu_nms_poly_fast per class with the thresholds of dota.py:324)
 * after the per-tile multiclass NMS (mmdet/core/post_processing/bbox_nms.py:32-52 per class).
 * Three stages, one entry point each, all on `stream` with no host synchronisation; with several GPUs the caller
 * all-reduces (sums) the keep masks between the stages (`world` ranks: stage 1 takes the tiles t % world == rank,
 * stage 2 the classes c % world == rank; world = 1, rank = 0 on one GPU).  Suppression is IoU > thr, no +1.
 *   boxes (n, fmt) tile-frame; labels, tile_ids (n) int32; tile_origins (n_tiles, 2) float32 (x, y) of each tile;
 *   thr: DEVICE float (per-tile NMS); merge_thr: DEVICE (n_classes) floats; keep_mask / survivors / kept: (n) uint8.
 * stage 1 writes scene_boxes (n, fmt) = boxes translated by their tile's origin and keep_mask (1 = survives its tile);
 * stage 2 takes the survivors of ALL ranks and writes keep_mask (1 = survives the merge); the compaction writes the kept
 *   detections class by class, ascending original index inside a class (one Task1_<class>.txt each, dota.py:296-308),
 *   into out_* (capacity n) and their number into the device int *n_out.  Labels / tiles out of range are never kept.
 * workspace: >= aidet_scene_workspace_bytes(n, n_tiles, n_classes, fmt) bytes, 128 B aligned (0 = bad arguments).   */
size_t aidet_scene_workspace_bytes(int n, int n_tiles, int n_classes, int fmt);
int aidet_scene_tile_nms_f32(const float* boxes, int fmt, const float* scores, const int* labels, const int* tile_ids,
                             const float* tile_origins, int n, int n_tiles, int n_classes, const float* thr, int world,
                             int rank, unsigned char* keep_mask, float* scene_boxes, void* workspace, size_t ws_bytes,
                             int device, void* stream);
int aidet_scene_merge_nms_f32(const float* scene_boxes, int fmt, const float* scores, const int* labels,
                              const unsigned char* survivors, int n, int n_tiles, int n_classes, const float* merge_thr,
                              int world, int rank, unsigned char* keep_mask, void* workspace, size_t ws_bytes, int device,
                              void* stream);
int aidet_scene_compact_f32(const float* scene_boxes, int fmt, const float* scores, const int* labels,
                            const unsigned char* kept, int n, int n_classes, float* out_boxes, float* out_scores,
                            int* out_labels, int* out_index, int* n_out, void* workspace, size_t ws_bytes, int device,
                            void* stream);

/* ---- Soft-NMS (axis-aligned, +1 convention), batched over groups ------------
 * Replaces: soft_nms_cpu_kernel (mmdet/ops/nms/src/nms_cpu.cpp:70-201), the reference's only Soft-NMS -- CUDA
 * tensors are copied to the host and back around it (mmdet/ops/nms/nms_wrapper.py:92-94,110-114).
 *   rows (n_total, 6) float32 IN/OUT: on entry [x1, y1, x2, y2, score, original index]; the rows of group g are
 *     rows[group_offsets[g] .. group_offsets[g+1]); on return the first n_out[g] rows of every group are its
 *     detections in selection order with decayed scores -- the (k, 6) result of nms_cpu.cpp:190-199.
 *   group_offsets (n_groups + 1) int32 device array; max_group = largest group size (host value)
 *   method 1 = linear, 2 = gaussian (nms_wrapper.py:104), 0 = hard; iou_thr, sigma, min_score as the reference
 * One CTA per group; same operation order in IEEE single precision as the C++ float instantiation. */
int aidet_soft_nms_f32(float* rows, const int* group_offsets, int n_groups, int max_group, float iou_thr, int method,
                       float sigma, float min_score, int* n_out, int device, void* stream);

/* ---- rotated / axis-aligned RoIAlign, multi-level, NHWC -------------------
 * Replaces: roi_align_cuda.forward_v1/v2, backward_v1/v2
 * (mmdet/ops/roi_align/src/roi_align_kernel.cu:64-141,187-283, roi_align_kernel_v2.cu:62-348)
 * and the per-level loop of mmdet/models/roi_extractors/single_level.py:89-107.
 *   feat[l]: (N, H[l], W[l], C) float32 channels-last, l < n_levels (HOST array of
 *            DEVICE pointers, likewise H, W, spatial_scale are host arrays)
 *   rois: (K, roi_fmt) roi_fmt 5 = [b,x1,y1,x2,y2], 6 = [b,cx,cy,w,h,theta]
 *   roi_level: (K) int32 level of each RoI, or NULL when n_levels == 1
 *   variant: 0 = v1 legacy (+1, roi_align_kernel.cu:79-86), 1 = v2 aligned=false,
 *            2 = v2 aligned=true (roi_align_kernel_v2.cu:79-90)
 *   out / grad_out: (K, ph, pw, C) float32
 *   backward accumulates into grad_feat[l] (caller zero-fills, as roi_align.py:63-64) */
int aidet_rroi_align_fwd_f32(const float* const* feat_host, const int* H_host, const int* W_host,
                             const float* scale_host, int n_levels, int N, int C,
                             const float* rois, int roi_fmt, const int* roi_level, int K,
                             int ph, int pw, int sample_num, int variant,
                             float* out, int device, void* stream);
int aidet_rroi_align_bwd_f32(const float* grad_out, float* const* grad_feat_host, const int* H_host,
                             const int* W_host, const float* scale_host, int n_levels, int N, int C,
                             const float* rois, int roi_fmt, const int* roi_level, int K,
                             int ph, int pw, int sample_num, int variant,
                             int device, void* stream);

/* Deterministic (atomic-free) backward: every feature pixel is written exactly once, so grad_feat
 * need NOT be zero-filled by the caller (it is overwritten, zeros included) and the result is
 * bit-reproducible when deterministic != 0 (stable radix sort of the taps; deterministic == 0 buckets
 * them with atomics, faster, summation order then varies like the reference's atomicAdd backward).
 * Same arguments as aidet_rroi_align_bwd_f32 plus a caller-owned workspace of
 * aidet_rroi_align_bwd_workspace_bytes(...) bytes, 256 B aligned.  Requires sample_num > 0,
 * C % 4 == 0 and 16 B aligned pointers (AIDET_EINVAL otherwise: use the scatter entry point).
 * Replaces: roi_align_cuda.backward_v1/v2 + the zero-fill of roi_align.py:63-64 and
 * roi_align_kernel_v2.cu:325-326. */
size_t aidet_rroi_align_bwd_workspace_bytes(const int* H_host, const int* W_host, int n_levels, int N, int K,
                                            int ph, int pw, int sample_num);
int aidet_rroi_align_bwd_gather_f32(const float* grad_out, float* const* grad_feat_host, const int* H_host,
                                    const int* W_host, const float* scale_host, int n_levels, int N, int C,
                                    const float* rois, int roi_fmt, const int* roi_level, int K,
                                    int ph, int pw, int sample_num, int variant, int deterministic,
                                    void* workspace, size_t ws_bytes, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIDET_B200_H_ */
