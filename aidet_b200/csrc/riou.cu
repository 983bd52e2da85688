// riou.cu -- pairwise rotated-IoU matrix / aligned pairs on sm_100a.
//
// Replaces (reference): the (m,n) overlap API of mmdet/core/bbox/geometry.py:4-88 for
// oriented boxes, whose arithmetic AIDet only reaches through wwtool
// (mmdet/datasets/dota.py:23,336).
//
// Layout / schedule:
//   prologue  : one thread per box -> records in the caller's workspace: "row" records (the box that
//               gets transformed: 48 B RectA / 64 B QuadRow) and "col" records (the box whose frame is
//               used: 32 B Rect / 64 B QuadCol), see geom.cuh.
//   main      : persistent CTAs of 256 threads.  A tile is TR rows x 256 columns; each
//               lane keeps ONE column record in registers and walks the row records,
//               which a single thread stages into shared memory with a 1-D TMA bulk
//               copy (cp.async.bulk + mbarrier, double buffered) so the copy of tile
//               t+1 overlaps the arithmetic of tile t.  Row records are read from
//               shared memory with warp-uniform 128-bit loads (broadcast); the result
//               row segment of a warp is one 128 B coalesced streaming store.
//               Per tile a probe picks the loop: solid tiles (every lane's bounding circle meets the
//               first two and the last row) run two clippings per step with no per-pair test, dense
//               tiles the dual-row loop behind a warp vote, sparse tiles the per-lane early-out loop.
//               Several destinations (row-sharded multi-GPU form): riou_matrix_tma_kernel, tiles leave
//               through one TMA tensor store per destination GPU.
//   bound     : FP32 issue (130 instructions per pair at 85 % of the issue slots: 240 Gpairs/s = 83 % of
//               the 256-flop/pair roofline; no tensor-core shape); output traffic 4 B/pair is ~1/7 of HBM
//               peak at that rate -- the DOTA-shaped (sparse) set IS bound by it (5 TB/s of stores).
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)

#include <type_traits>

#include "common.cuh"
#include "geom.cuh"
#include "pairop.cuh"

namespace aidet {

constexpr int kColsPerTile = 256;
constexpr int kMaxTileRows = 64;

// Destination set of one launch.  n == 1: the local result matrix.  n > 1 (row-sharded multi-GPU form):
// the SAME row block of the (m, n) buffer on every GPU of the box -- p[q] points into rank q's buffer,
// mapped into this process through CUDA peer / symmetric memory -- so the all-gather of the shard
// results happens tile by tile from inside the kernel as plain stores over NVLink, overlapped with the
// clipping arithmetic, instead of as a separate collective after it.
constexpr int kMaxPeers = 8;
struct OutSet { float* p[kMaxPeers]; int n; };

// NVSwitch multicast store: ONE store instruction, the switch replicates it into every GPU of the multicast group
// (NVLS), so a row-sharded rank sends its block over its NVLink once instead of once per peer.
__device__ __forceinline__ void st_multicast(float* mc_ptr, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_ptr), "f"(v) : "memory");
}

enum { STORE_LOCAL = 0, STORE_PEERS = 1, STORE_MCAST = 2 };

// kernel-tuning knobs (compile time; see scripts/build_variants.sh): resident CTAs per SM the register allocation
// targets, and the unroll factor of the pair loop
#ifndef AIDET_RIOU_MINB
#define AIDET_RIOU_MINB 4      // 64 registers: room for the two interleaved clippings of the dual-row step
#endif
#ifndef AIDET_RIOU_DUAL
#define AIDET_RIOU_DUAL 1      // dense 32768^2: 199.8 vs 192.3 Gpairs/s, DOTA-shaped 335.9 vs 334.4 (r1e measurements)
#endif
#ifndef AIDET_RIOU_UNROLL
#define AIDET_RIOU_UNROLL 2
#endif
constexpr int kRiouUnroll = AIDET_RIOU_UNROLL;

// (the 8-point kind carries both the parallelogram and the general-quad arithmetic: 3 CTAs per SM, 80 registers)
template <class K, int MODE, int STORE>
__global__ void __launch_bounds__(kColsPerTile, K::FMT == 8 ? 3 : AIDET_RIOU_MINB)
riou_matrix_kernel(const typename PairOp<K>::S* __restrict__ rows, int m,
                   const typename PairOp<K>::R* __restrict__ cols, int n,
                   OutSet outs, long long ld, int tile_rows, int n_row_tiles, int n_tiles,
                   int tiles_per_cta, float* __restrict__ scratch) {
  using P = PairOp<K>;
  using S = typename P::S; using R = typename P::R;
  __shared__ __align__(128) S stage[2][kMaxTileRows];
  __shared__ __align__(8) uint64_t bar[2];

  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(t_begin + tiles_per_cta, n_tiles);
  if (t_begin >= t_end) return;

  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  __syncthreads();

  auto issue = [&](int t, int buf) {      // thread 0 only
    int rt = t % n_row_tiles;
    int r0 = rt * tile_rows;
    uint32_t bytes = (uint32_t)(min(tile_rows, m - r0) * (int)sizeof(S));
    mbar_expect_tx(&bar[buf], bytes);
    tma_load_1d(&stage[buf][0], rows + r0, bytes, &bar[buf]);
  };
  if (threadIdx.x == 0) issue(t_begin, 0);

  int cur_ct = -1;
  typename P::X me;
  bool live = false;
  int col = 0;
  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const int buf = it & 1;
    if (threadIdx.x == 0 && t + 1 < t_end) issue(t + 1, buf ^ 1);
    const int ct = t / n_row_tiles, rt = t % n_row_tiles;
    if (ct != cur_ct) {
      cur_ct = ct;
      col = ct * kColsPerTile + threadIdx.x;
      live = col < n;
      me = cols[live ? col : n - 1];
    }
    mbar_wait(&bar[buf], (it >> 1) & 1);
    const int r0 = rt * tile_rows;
    const int nr = min(tile_rows, m - r0);
    long long off = (long long)r0 * ld + col;        // advanced by one row per iteration
    const S* st = &stage[buf][0];
    // single destination: a running pointer; lanes of columns past n write a scratch line (step 0) instead of
    // branching around every store
    float* pa = live ? outs.p[0] + off : scratch + threadIdx.x;
    const long long step = live ? ld : 0;
    auto put = [&](float v) {
      if (STORE == STORE_LOCAL) {
        __stcs(pa, v);
        pa += step;
      } else if (live) {
        if (STORE == STORE_PEERS) {
#pragma unroll
          for (int q = 0; q < kMaxPeers; ++q) if (q < outs.n) __stcs(outs.p[q] + off, v);
        } else if (STORE == STORE_MCAST) {
          st_multicast(outs.p[0] + off, v);
        } else {
          __stcs(outs.p[0] + off, v);
        }
      }
      off += ld;
    };
    int r = 0;
#if AIDET_RIOU_DUAL
    if constexpr (!std::is_same<K, HbbKind>::value) {
      // two rows per step: when both rows have a lane whose bounding circles meet (the dense case), the two clippings
      // run as ONE straight-line block, so the scheduler can interleave two independent dependency chains.
      // Probe: do the first two rows of the tile both have a lane whose bounding circles meet this warp's columns?
      // Dense tiles then take the dual-row loop, sparse tiles the plain per-lane early-out loop below (both give
      // identical values; the probe only picks the faster schedule).
      bool dual = false, solid = false;
      if (nr >= 2) {
        const bool n0 = P::near(st[0], me), n1 = P::near(st[1], me);
        dual = __any_sync(0xffffffffu, n0) && __any_sync(0xffffffffu, n1);
        // solid: every lane's bounding circle meets the tile's first two and last rows -- the dense set.  Those tiles run
        // without the per-pair circle test, its votes and the final select (7 % of the instructions): a pair whose circles
        // are disjoint after all still gets its edge integrals, which give 0 or rounding noise ~1e-7 of the smaller area
        // (the same as a pair whose circles meet and whose boxes do not).
        solid = __all_sync(0xffffffffu, n0 && n1 && P::near(st[nr - 1], me));
      }
      if (solid) {
#pragma unroll 1
        for (; r + 2 <= nr; r += 2) {
          const S sa = st[r], sb = st[r + 1];
          const float ia = K::inter(sa, me), ib = K::inter(sb, me);
          put(finish_overlap(ia, K::area_row(sa), K::area_col(me), MODE));
          put(finish_overlap(ib, K::area_row(sb), K::area_col(me), MODE));
        }
      } else if (dual) {
#pragma unroll 1
        for (; r + 2 <= nr; r += 2) {
          const S sa = st[r], sb = st[r + 1];
          const bool ha = P::near(sa, me), hb = P::near(sb, me);
          const bool wa = __any_sync(0xffffffffu, ha), wb = __any_sync(0xffffffffu, hb);
          float va = 0.0f, vb = 0.0f;
          if (wa && wb) {
            const float ia = K::inter(sa, me), ib = K::inter(sb, me);
            va = ha ? finish_overlap(ia, K::area_row(sa), K::area_col(me), MODE) : 0.0f;
            vb = hb ? finish_overlap(ib, K::area_row(sb), K::area_col(me), MODE) : 0.0f;
          } else if (wa) {
            va = ha ? finish_overlap(K::inter(sa, me), K::area_row(sa), K::area_col(me), MODE) : 0.0f;
          } else if (wb) {
            vb = hb ? finish_overlap(K::inter(sb, me), K::area_row(sb), K::area_col(me), MODE) : 0.0f;
          }
          put(va);
          put(vb);
        }
      }
    }
#endif
#pragma unroll kRiouUnroll
    for (; r < nr; ++r) put(P::overlap(st[r], me, MODE));
    __syncthreads();      // everyone is done with stage[buf] before it is refilled
  }
}


// ------------------------------------------------------------------ multi-GPU form with TMA tensor stores
// The peer-store kernel above sends every value with its own 4-byte STG: 128 bytes per warp, row and peer -- small
// NVLink packets, one store instruction per destination (r1: 594-678 GB/s of egress, the overlap lost 22 % at N = 2).
// Here a CTA collects the (16 rows x 256 columns) tile in shared memory (conflict-free STS, one per pair as before) and ONE
// thread hands it to the TMA unit: one cp.async.bulk.tensor.2d store per destination GPU (16 KB each, 1 KB rows, clipped
// at the matrix edge by the tensor map), local copy included.  The LSU issues no global store at all, the copy engine
// streams full-size NVLink packets, and the stores of tile t overlap the arithmetic of tile t+1 (two tile buffers,
// bulk-group commit / wait_group.read before a buffer is rewritten).
#ifndef AIDET_TMA_ROWS
#define AIDET_TMA_ROWS 16
#endif
constexpr int kTmaTileRows = AIDET_TMA_ROWS;      // 16 rows: 16 KB tile buffers
#ifndef AIDET_TMA_OBUFS
#define AIDET_TMA_OBUFS 4             // tile buffers per CTA (stores in flight + 1); N = 2 on one box: 309 / 310 / 329 Gpairs/s with 2 / 3 / 4
#endif
constexpr int kTmaOutBufs = AIDET_TMA_OBUFS;

struct TmaOuts { CUtensorMap map[kMaxPeers]; int n; };

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int r0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(r0),
               "r"(smem_u32(smem_src))
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

template <class K, int MODE>
__global__ void __launch_bounds__(kColsPerTile, K::FMT == 8 ? 3 : 4)
riou_matrix_tma_kernel(const typename PairOp<K>::S* __restrict__ rows, int m,
                       const typename PairOp<K>::R* __restrict__ cols, int n,
                       const __grid_constant__ TmaOuts outs, int n_row_tiles, int n_tiles, int tiles_per_cta) {
  using P = PairOp<K>;
  using S = typename P::S; using R = typename P::R;
  extern __shared__ __align__(128) unsigned char dyn[];
  float* otile = reinterpret_cast<float*>(dyn);                           // [kTmaOutBufs][kTmaTileRows][256]
  S* stage = reinterpret_cast<S*>(dyn + kTmaOutBufs * kTmaTileRows * kColsPerTile * 4);   // [2][kTmaTileRows]
  __shared__ __align__(8) uint64_t bar[2];

  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(t_begin + tiles_per_cta, n_tiles);
  if (t_begin >= t_end) return;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  __syncthreads();
  auto issue = [&](int t, int buf) {      // thread 0 only
    const int r0 = (t % n_row_tiles) * kTmaTileRows;
    const uint32_t bytes = (uint32_t)(min(kTmaTileRows, m - r0) * (int)sizeof(S));
    mbar_expect_tx(&bar[buf], bytes);
    tma_load_1d(stage + buf * kTmaTileRows, rows + r0, bytes, &bar[buf]);
  };
  if (threadIdx.x == 0) issue(t_begin, 0);

  int cur_ct = -1;
  typename P::X me;
  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const int buf = it & 1;
    if (threadIdx.x == 0 && t + 1 < t_end) issue(t + 1, buf ^ 1);
    const int ct = t / n_row_tiles, rt = t % n_row_tiles;
    if (ct != cur_ct) {
      cur_ct = ct;
      const int col = ct * kColsPerTile + threadIdx.x;
      me = cols[col < n ? col : n - 1];                    // columns past n are computed and clipped by the tensor map
    }
    mbar_wait(&bar[buf], (it >> 1) & 1);
    const int r0 = rt * kTmaTileRows;
    const int nr = min(kTmaTileRows, m - r0);
    const S* st = stage + buf * kTmaTileRows;
    const int ob = it % kTmaOutBufs;
    float* ot = otile + ob * (kTmaTileRows * kColsPerTile) + threadIdx.x;
    int r = 0;
    if constexpr (!std::is_same<K, HbbKind>::value) {
      bool dual = false, solid = false;
      if (nr >= 2) {
        const bool n0 = P::near(st[0], me), n1 = P::near(st[1], me);
        dual = __any_sync(0xffffffffu, n0) && __any_sync(0xffffffffu, n1);
        solid = __all_sync(0xffffffffu, n0 && n1 && P::near(st[nr - 1], me));       // see riou_matrix_kernel
      }
      if (solid) {
#pragma unroll 1
        for (; r + 2 <= nr; r += 2) {
          const S sa = st[r], sb = st[r + 1];
          const float ia = K::inter(sa, me), ib = K::inter(sb, me);
          ot[r * kColsPerTile] = finish_overlap(ia, K::area_row(sa), K::area_col(me), MODE);
          ot[(r + 1) * kColsPerTile] = finish_overlap(ib, K::area_row(sb), K::area_col(me), MODE);
        }
      } else if (dual) {
#pragma unroll 1
        for (; r + 2 <= nr; r += 2) {
          const S sa = st[r], sb = st[r + 1];
          const bool ha = P::near(sa, me), hb = P::near(sb, me);
          const bool wa = __any_sync(0xffffffffu, ha), wb = __any_sync(0xffffffffu, hb);
          float va = 0.0f, vb = 0.0f;
          if (wa && wb) {
            const float ia = K::inter(sa, me), ib = K::inter(sb, me);
            va = ha ? finish_overlap(ia, K::area_row(sa), K::area_col(me), MODE) : 0.0f;
            vb = hb ? finish_overlap(ib, K::area_row(sb), K::area_col(me), MODE) : 0.0f;
          } else if (wa) {
            va = ha ? finish_overlap(K::inter(sa, me), K::area_row(sa), K::area_col(me), MODE) : 0.0f;
          } else if (wb) {
            vb = hb ? finish_overlap(K::inter(sb, me), K::area_row(sb), K::area_col(me), MODE) : 0.0f;
          }
          ot[r * kColsPerTile] = va;
          ot[(r + 1) * kColsPerTile] = vb;
        }
      }
    }
#pragma unroll kRiouUnroll
    for (; r < nr; ++r) ot[r * kColsPerTile] = P::overlap(st[r], me, MODE);
    // hand the tile to the copy engine: the writes above (generic proxy) become visible to the async proxy, then one
    // thread issues a tensor store per destination; rows past m / columns past n are clipped by the tensor map
    fence_proxy_async();
    __syncthreads();                                        // tile complete; stage[buf] free for the load of tile t+2
    if (threadIdx.x == 0) {
      const float* src = otile + ob * (kTmaTileRows * kColsPerTile);
#pragma unroll
      for (int q = 0; q < kMaxPeers; ++q)
        if (q < outs.n) tma_store_2d(&outs.map[q], src, ct * kColsPerTile, r0);
      tma_commit();
      tma_wait_read<kTmaOutBufs - 1>();                     // the stores of the tile that used the next buffer have read it: it may be rewritten
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) tma_wait_all<0>();                  // every store has completed before the CTA retires
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// (m, n) float32 row block at `base`, row stride ld elements -> tensor map with a (256 x 32) box.  false: not expressible.
static bool make_out_map(CUtensorMap* map, float* base, int m, int n, long long ld) {
  EncodeTiledFn enc = encode_tiled_fn();
  // strides and addresses in 16-byte units; the store clips at the tensor edge in 16-byte units as well (measured: with
  // n % 4 == 2 the two floats after the row end were written), so rows must END on a 16-byte boundary too
  if (!enc || (ld & 3) || (n & 3) || ((uintptr_t)base & 15)) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kColsPerTile, (cuuint32_t)kTmaTileRows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class K>
__global__ void __launch_bounds__(256) riou_aligned_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           int n, int mode, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pa[K::FMT], pb[K::FMT];
#pragma unroll
  for (int k = 0; k < K::FMT; k++) { pa[k] = a[(size_t)i * K::FMT + k]; pb[k] = b[(size_t)i * K::FMT + k]; }
  typename K::Row r; typename K::Col c;
  K::prepare(pa, &r, nullptr);
  K::prepare(pb, nullptr, &c);
  out[i] = PairOp<K>::overlap(r, typename K::Reg(c), mode);
}

// d overlap(a[i], b[i]) / d (box parameters) of both boxes, scaled by the upstream gradient: the backward of the aligned
// overlap, i.e. of the rotated IoU loss (rotated counterpart of mmdet/models/losses/iou_loss.py:10-27).  FMT 5:
// (cx,cy,w,h,theta), geom.cuh: rect_overlap_grad; FMT 8: the corner coordinates of convex quads, quad_overlap_grad.
// One thread per pair.
template <int FMT>
__global__ void __launch_bounds__(256) riou_aligned_grad_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                int n, int mode, const float* __restrict__ grad_ov,
                                                                float* __restrict__ ov, float* __restrict__ grad_a,
                                                                float* __restrict__ grad_b) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pa[FMT], pb[FMT], ga[FMT], gb[FMT];
#pragma unroll
  for (int k = 0; k < FMT; k++) { pa[k] = a[(size_t)i * FMT + k]; pb[k] = b[(size_t)i * FMT + k]; }
  const float v = (FMT == 5) ? rect_overlap_grad(pa, pb, mode, ga, gb) : quad_overlap_grad(pa, pb, mode, ga, gb);
  const float go = grad_ov ? grad_ov[i] : 1.0f;
  if (ov) ov[i] = v;
#pragma unroll
  for (int k = 0; k < FMT; k++) {
    if (grad_a) grad_a[(size_t)i * FMT + k] = go * ga[k];
    if (grad_b) grad_b[(size_t)i * FMT + k] = go * gb[k];
  }
}

template <class K>
static int launch_matrix(const float* a, int m, const float* b, int n, int mode, const OutSet& outs, long long ld,
                         void* ws, int device, cudaStream_t s, bool mcast = false, bool force_tma = false) {
  using P = PairOp<K>;
  using S = typename P::S; using R = typename P::R;
  S* rows = reinterpret_cast<S*>(ws);
  R* cols = reinterpret_cast<R*>(reinterpret_cast<char*>(ws) + align_up((size_t)m * sizeof(S), 128));
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(cols) + align_up((size_t)n * sizeof(R), 128));   // 1 KB
  riou_prepare_both_kernel<K><<<ceil_div(m + n, 256), 256, 0, s>>>(a, m, b, n, rows, cols);
  const int sms = sm_count(device);
  const int n_col_tiles = ceil_div(n, kColsPerTile);
  if ((outs.n > 1 || force_tma) && !mcast) {
    // several destinations (row-sharded multi-GPU form): tiles leave through TMA tensor stores when every destination
    // is expressible as a tensor map (16-byte aligned base, row stride a multiple of 4 elements)
    TmaOuts t{};
    bool ok = true;
    for (int q = 0; q < outs.n && ok; ++q) ok = make_out_map(&t.map[q], outs.p[q], m, n, ld);
    if (ok) {
      t.n = outs.n;
      const int n_row_tiles = ceil_div(m, kTmaTileRows);
      const long long n_tiles_ll = (long long)n_row_tiles * n_col_tiles;
      if (n_tiles_ll > 0x7fffffffLL) { set_error("riou: problem too large (%lld tiles)", n_tiles_ll); return AIDET_EINVAL; }
      const int n_tiles = (int)n_tiles_ll;
      int grid = min(n_tiles, sms * 12);
      const int tiles_per_cta = ceil_div(n_tiles, grid);
      grid = ceil_div(n_tiles, tiles_per_cta);
      const size_t smem = (size_t)kTmaOutBufs * kTmaTileRows * kColsPerTile * 4 + 2 * (size_t)kTmaTileRows * sizeof(S);
      ProfScope prof(PROF_RIOU, s);
      if (mode == MODE_IOF) {
        AIDET_CUDA(cudaFuncSetAttribute(riou_matrix_tma_kernel<K, MODE_IOF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        riou_matrix_tma_kernel<K, MODE_IOF><<<grid, kColsPerTile, smem, s>>>(rows, m, cols, n, t, n_row_tiles, n_tiles, tiles_per_cta);
      } else {
        AIDET_CUDA(cudaFuncSetAttribute(riou_matrix_tma_kernel<K, MODE_IOU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        riou_matrix_tma_kernel<K, MODE_IOU><<<grid, kColsPerTile, smem, s>>>(rows, m, cols, n, t, n_row_tiles, n_tiles, tiles_per_cta);
      }
      count_launch(2);
      AIDET_CUDA(cudaGetLastError());
      return AIDET_OK;
    }
  }
  int tile_rows = kMaxTileRows;
  while (tile_rows > 8 && (long long)ceil_div(m, tile_rows) * n_col_tiles < 8LL * sms) tile_rows >>= 1;
  const int n_row_tiles = ceil_div(m, tile_rows);
  const long long n_tiles_ll = (long long)n_row_tiles * n_col_tiles;
  if (n_tiles_ll > 0x7fffffffLL) { set_error("riou: problem too large (%lld tiles)", n_tiles_ll); return AIDET_EINVAL; }
  const int n_tiles = (int)n_tiles_ll;
  // enough resident CTAs to fill the machine, contiguous tile ranges per CTA so the column
  // record stays in registers across consecutive row tiles
  int grid = min(n_tiles, sms * 8);
  int tiles_per_cta = ceil_div(n_tiles, grid);
  grid = ceil_div(n_tiles, tiles_per_cta);
  {
    ProfScope prof(PROF_RIOU, s);
#define AIDET_LAUNCH_RIOU(MODE_, STORE_)                                                                        \
  riou_matrix_kernel<K, MODE_, STORE_><<<grid, kColsPerTile, 0, s>>>(rows, m, cols, n, outs, ld, tile_rows,     \
                                                                      n_row_tiles, n_tiles, tiles_per_cta, scratch)
    if (mcast)           { if (mode == MODE_IOF) AIDET_LAUNCH_RIOU(MODE_IOF, STORE_MCAST); else AIDET_LAUNCH_RIOU(MODE_IOU, STORE_MCAST); }
    else if (outs.n > 1) { if (mode == MODE_IOF) AIDET_LAUNCH_RIOU(MODE_IOF, STORE_PEERS); else AIDET_LAUNCH_RIOU(MODE_IOU, STORE_PEERS); }
    else                 { if (mode == MODE_IOF) AIDET_LAUNCH_RIOU(MODE_IOF, STORE_LOCAL); else AIDET_LAUNCH_RIOU(MODE_IOU, STORE_LOCAL); }
#undef AIDET_LAUNCH_RIOU
  }
  count_launch(2);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

}  // namespace aidet

using namespace aidet;

extern "C" {

size_t aidet_riou_workspace_bytes(int m, int n, int fmt) {
  size_t rec = record_bytes(fmt);
  return align_up((size_t)(m > 0 ? m : 0) * rec, 128) + align_up((size_t)(n > 0 ? n : 0) * rec, 128) + 1024 + 128;   // + scratch line
}

int aidet_riou_matrix_f32(const float* a, int m, const float* b, int n, int fmt, int mode, float* out,
                          long long ld_out, void* workspace, size_t ws_bytes, int device, void* stream) {
  AIDET_REQUIRE(fmt == 4 || fmt == 5 || fmt == 8, "aidet_riou_matrix_f32: fmt must be 4, 5 or 8, got %d", fmt);
  AIDET_REQUIRE(mode == AIDET_MODE_IOU || mode == AIDET_MODE_IOF, "aidet_riou_matrix_f32: bad mode %d", mode);
  AIDET_REQUIRE(m >= 0 && n >= 0, "aidet_riou_matrix_f32: negative size");
  if (m == 0 || n == 0) return AIDET_OK;
  AIDET_REQUIRE(a && b && out && workspace, "aidet_riou_matrix_f32: null pointer");
  AIDET_REQUIRE(ld_out >= n, "aidet_riou_matrix_f32: ld_out %lld < n %d", ld_out, n);
  AIDET_REQUIRE(((uintptr_t)workspace & 15) == 0, "aidet_riou_matrix_f32: workspace must be 16 B aligned");
  if (ws_bytes < aidet_riou_workspace_bytes(m, n, fmt)) {
    set_error("aidet_riou_matrix_f32: workspace %zu < %zu", ws_bytes, aidet_riou_workspace_bytes(m, n, fmt));
    return AIDET_EWORKSPACE;
  }
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  OutSet outs{}; outs.p[0] = out; outs.n = 1;
  if (fmt == 5) return launch_matrix<RectKind>(a, m, b, n, mode, outs, ld_out, workspace, device, s);
  if (fmt == 4) return launch_matrix<HbbKind>(a, m, b, n, mode, outs, ld_out, workspace, device, s);
  return launch_matrix<QuadKind>(a, m, b, n, mode, outs, ld_out, workspace, device, s);
}

int aidet_riou_matrix_multi_f32(const float* a, int m, const float* b, int n, int fmt, int mode,
                                float* const* outs_host, int n_outs, long long ld_out, void* workspace,
                                size_t ws_bytes, int device, void* stream) {
  AIDET_REQUIRE(fmt == 5 || fmt == 8, "aidet_riou_matrix_multi_f32: fmt must be 5 or 8, got %d", fmt);
  AIDET_REQUIRE(mode == AIDET_MODE_IOU || mode == AIDET_MODE_IOF, "aidet_riou_matrix_multi_f32: bad mode %d", mode);
  AIDET_REQUIRE(m >= 0 && n >= 0, "aidet_riou_matrix_multi_f32: negative size");
  AIDET_REQUIRE(outs_host && n_outs >= 1 && n_outs <= kMaxPeers, "aidet_riou_matrix_multi_f32: n_outs must be in [1,%d]", kMaxPeers);
  if (m == 0 || n == 0) return AIDET_OK;
  AIDET_REQUIRE(a && b && workspace, "aidet_riou_matrix_multi_f32: null pointer");
  AIDET_REQUIRE(ld_out >= n, "aidet_riou_matrix_multi_f32: ld_out %lld < n %d", ld_out, n);
  AIDET_REQUIRE(((uintptr_t)workspace & 15) == 0, "aidet_riou_matrix_multi_f32: workspace must be 16 B aligned");
  if (ws_bytes < aidet_riou_workspace_bytes(m, n, fmt)) {
    set_error("aidet_riou_matrix_multi_f32: workspace %zu < %zu", ws_bytes, aidet_riou_workspace_bytes(m, n, fmt));
    return AIDET_EWORKSPACE;
  }
  OutSet outs{};
  for (int q = 0; q < n_outs; q++) {
    AIDET_REQUIRE(outs_host[q], "aidet_riou_matrix_multi_f32: null destination %d", q);
    outs.p[q] = outs_host[q];
  }
  outs.n = n_outs;
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (fmt == 5) return launch_matrix<RectKind>(a, m, b, n, mode, outs, ld_out, workspace, device, s, false, true);
  return launch_matrix<QuadKind>(a, m, b, n, mode, outs, ld_out, workspace, device, s, false, true);
}

int aidet_riou_matrix_mcast_f32(const float* a, int m, const float* b, int n, int fmt, int mode, float* out_mc,
                                long long ld_out, void* workspace, size_t ws_bytes, int device, void* stream) {
  AIDET_REQUIRE(fmt == 5 || fmt == 8, "aidet_riou_matrix_mcast_f32: fmt must be 5 or 8, got %d", fmt);
  AIDET_REQUIRE(mode == AIDET_MODE_IOU || mode == AIDET_MODE_IOF, "aidet_riou_matrix_mcast_f32: bad mode %d", mode);
  AIDET_REQUIRE(m >= 0 && n >= 0, "aidet_riou_matrix_mcast_f32: negative size");
  if (m == 0 || n == 0) return AIDET_OK;
  AIDET_REQUIRE(a && b && out_mc && workspace, "aidet_riou_matrix_mcast_f32: null pointer");
  AIDET_REQUIRE(ld_out >= n, "aidet_riou_matrix_mcast_f32: ld_out %lld < n %d", ld_out, n);
  AIDET_REQUIRE(((uintptr_t)workspace & 15) == 0, "aidet_riou_matrix_mcast_f32: workspace must be 16 B aligned");
  if (ws_bytes < aidet_riou_workspace_bytes(m, n, fmt)) {
    set_error("aidet_riou_matrix_mcast_f32: workspace %zu < %zu", ws_bytes, aidet_riou_workspace_bytes(m, n, fmt));
    return AIDET_EWORKSPACE;
  }
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  OutSet outs{}; outs.p[0] = out_mc; outs.n = 1;
  if (fmt == 5) return launch_matrix<RectKind>(a, m, b, n, mode, outs, ld_out, workspace, device, s, true);
  return launch_matrix<QuadKind>(a, m, b, n, mode, outs, ld_out, workspace, device, s, true);
}

int aidet_riou_aligned_f32(const float* a, const float* b, int n, int fmt, int mode, float* out, int device,
                           void* stream) {
  AIDET_REQUIRE(fmt == 4 || fmt == 5 || fmt == 8, "aidet_riou_aligned_f32: fmt must be 4, 5 or 8, got %d", fmt);
  AIDET_REQUIRE(mode == AIDET_MODE_IOU || mode == AIDET_MODE_IOF, "aidet_riou_aligned_f32: bad mode %d", mode);
  AIDET_REQUIRE(n >= 0, "aidet_riou_aligned_f32: negative size");
  if (n == 0) return AIDET_OK;
  AIDET_REQUIRE(a && b && out, "aidet_riou_aligned_f32: null pointer");
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (fmt == 5) riou_aligned_kernel<RectKind><<<ceil_div(n, 256), 256, 0, s>>>(a, b, n, mode, out);
  else if (fmt == 4) riou_aligned_kernel<HbbKind><<<ceil_div(n, 256), 256, 0, s>>>(a, b, n, mode, out);
  else riou_aligned_kernel<QuadKind><<<ceil_div(n, 256), 256, 0, s>>>(a, b, n, mode, out);
  count_launch(1);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

int aidet_riou_aligned_grad_f32(const float* a, const float* b, int n, int fmt, int mode, const float* grad_ov,
                                float* ov, float* grad_a, float* grad_b, int device, void* stream) {
  AIDET_REQUIRE(fmt == 5 || fmt == 8, "aidet_riou_aligned_grad_f32: fmt must be 5 (theta-OBB) or 8 (convex point-OBB), got %d", fmt);
  AIDET_REQUIRE(mode == AIDET_MODE_IOU || mode == AIDET_MODE_IOF, "aidet_riou_aligned_grad_f32: bad mode %d", mode);
  AIDET_REQUIRE(n >= 0, "aidet_riou_aligned_grad_f32: negative size");
  if (n == 0) return AIDET_OK;
  AIDET_REQUIRE(a && b, "aidet_riou_aligned_grad_f32: null pointer");
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (fmt == 5) riou_aligned_grad_kernel<5><<<ceil_div(n, 256), 256, 0, s>>>(a, b, n, mode, grad_ov, ov, grad_a, grad_b);
  else riou_aligned_grad_kernel<8><<<ceil_div(n, 256), 256, 0, s>>>(a, b, n, mode, grad_ov, ov, grad_a, grad_b);
  count_launch(1);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

}  // extern "C"
