// common.cuh -- shared host/device plumbing of libaidet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/aidet_b200.h"

namespace aidet {

// ---- errors (thread-local message, C-ABI return codes) ---------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define AIDET_CUDA(call)                                              \
  do {                                                                \
    cudaError_t e_ = (call);                                          \
    if (e_ != cudaSuccess) return ::aidet::cuda_fail(e_, #call);      \
  } while (0)

#define AIDET_REQUIRE(cond, ...)                                      \
  do {                                                                \
    if (!(cond)) { ::aidet::set_error(__VA_ARGS__); return AIDET_EINVAL; } \
  } while (0)

// ---- launch accounting + per-op kernel timing ------------------------------
enum { PROF_RIOU = 0, PROF_NMS_MASK = 1, PROF_ROI_FWD = 2, PROF_ROI_BWD = 3, PROF_KINDS = 4 };
void count_launch(int n = 1);
int prof_level();                      // aidet_prof_enable's argument (0 off, 1 kernel timing, 2 + fused-NMS phase stamps)
// RAII: records a start event at construction and a stop event at destruction on
// `s` when profiling is enabled; the pair is resolved lazily in aidet_prof_read.
struct ProfScope {
  int kind; cudaStream_t s; bool on; cudaEvent_t e0, e1;
  ProfScope(int kind, cudaStream_t s);
  ~ProfScope();
};

int set_device(int device);
int sm_count(int device);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

#if defined(__CUDACC__)
// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// Wait for the phase with the given parity to complete.  try_wait suspends the thread in hardware for a bounded time
// per attempt; failed attempts back off with nanosleep, and a wait that outlasts ~4 s (a copy that can never
// arrive: bad pointer, lost producer) traps instead of hanging the device, so the host sees a launch failure.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 64u) __nanosleep(spins > 4096u ? 256u : 32u);
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16 B aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
#endif

}  // namespace aidet
