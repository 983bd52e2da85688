// Latency of the intra-CTA synchronisation primitives the fused NMS scan uses (one warp, dependent back-to-back use).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../aidet_b200/csrc/common.cuh"
using namespace aidet;

__device__ __forceinline__ uint32_t ld_acq(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }
__device__ __forceinline__ void st_rel(uint32_t* p, uint32_t v) { asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ __forceinline__ void st_vol(uint32_t* p, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }

__global__ void k(long long* out, int iters) {
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t flag[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); flag[0] = 0; flag[1] = 0; }
  __syncthreads();
  long long t0, t1; uint32_t acc = 0;
  if (warp == 0) {
    // 1. arrive (count 1: completes a phase each time) + try_wait on the completed phase
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { if (lane == 0) mbar_arrive(&bar[0]); __syncwarp(); mbar_wait(&bar[0], i & 1); }
    t1 = clock64(); if (lane == 0) out[0] = (t1 - t0) / iters;
    // 2. arrive alone
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { if (lane == 0) mbar_arrive(&bar[1]); }
    t1 = clock64(); if (lane == 0) out[1] = (t1 - t0) / iters;
    // 3. try_wait on an already completed phase (dependent chain through acc)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) acc += mbar_try_wait(&bar[0], (iters & 1) ^ 1) ? 1 : 0;
    t1 = clock64(); if (lane == 0) out[2] = (t1 - t0) / iters;
    // 4. ld.acquire.shared dependent chain
    t0 = clock64();
    uint32_t idx = 0;
    for (int i = 0; i < iters; ++i) idx = ld_acq(&flag[idx & 1]);
    t1 = clock64(); if (lane == 0) out[3] = (t1 - t0) / iters; acc += idx;
    // 5. ld.volatile.shared dependent chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = ld_vol(&flag[idx & 1]);
    t1 = clock64(); if (lane == 0) out[4] = (t1 - t0) / iters; acc += idx;
    // 6. st.release.shared back to back
    t0 = clock64();
    for (int i = 0; i < iters; ++i) if (lane == 0) st_rel(&flag[2], i);
    t1 = clock64(); if (lane == 0) out[5] = (t1 - t0) / iters;
    // 7. st.volatile
    t0 = clock64();
    for (int i = 0; i < iters; ++i) if (lane == 0) st_vol(&flag[3], i);
    t1 = clock64(); if (lane == 0) out[6] = (t1 - t0) / iters;
    // 8. redux.or dependent chain
    t0 = clock64();
    uint32_t r = lane;
    for (int i = 0; i < iters; ++i) r = __reduce_or_sync(0xffffffffu, r + i);
    t1 = clock64(); if (lane == 0) out[7] = (t1 - t0) / iters; acc += r;
    // 9. vote.all dependent
    t0 = clock64();
    bool p = true;
    for (int i = 0; i < iters; ++i) p = __all_sync(0xffffffffu, p || (lane + i) > 100000);
    t1 = clock64(); if (lane == 0) out[8] = (t1 - t0) / iters; acc += p;
    // 10. shfl dependent chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) r = __shfl_sync(0xffffffffu, r, (i + 1) & 31);
    t1 = clock64(); if (lane == 0) out[9] = (t1 - t0) / iters; acc += r;
    // 11. 32 independent shfl + 32-step dependent LOP chain (the greedy chain)
    t0 = clock64();
    uint32_t cur = r, keep = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) { const uint32_t dk = __shfl_sync(0xffffffffu, r + i, kk); if (!((cur >> kk) & 1u)) { keep |= 1u << kk; cur |= dk & 0x55555555u; } }
      cur = keep >> 3;
    }
    t1 = clock64(); if (lane == 0) out[10] = (t1 - t0) / iters; acc += cur + keep;
    // 12. ping-pong through shared memory flags between warp 0 and warp 1 (round trip)
    t0 = clock64();
    for (int i = 1; i <= iters; ++i) { if (lane == 0) { st_vol(&flag[4], i); while (ld_vol(&flag[5]) < (uint32_t)i) {} } __syncwarp(); }
    t1 = clock64(); if (lane == 0) out[11] = (t1 - t0) / iters;
    // 13. ping-pong through mbarriers (warp 0 arrives on bar A, waits bar B)
    if (lane == 0) st_vol(&flag[6], 1);
  } else if (warp == 1) {
    for (int i = 1; i <= iters; ++i) { if (lane == 0) { while (ld_vol(&flag[4]) < (uint32_t)i) {} st_vol(&flag[5], i); } __syncwarp(); }
  }
  if (acc == 0x12345678u) out[15] = acc;
}

__global__ void k2(long long* out, int iters) {   // mbarrier ping-pong between two warps
  __shared__ __align__(8) uint64_t a, b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&a, 1); mbar_init(&b, 1); fence_barrier_init(); }
  __syncthreads();
  if (warp == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { if (lane == 0) mbar_arrive(&a); mbar_wait(&b, i & 1); }
    long long t1 = clock64(); if (lane == 0) out[12] = (t1 - t0) / iters;
  } else {
    for (int i = 0; i < iters; ++i) { mbar_wait(&a, i & 1); if (lane == 0) mbar_arrive(&b); }
  }
}

int main() {
  long long* d; cudaMalloc(&d, 16 * 8); cudaMemset(d, 0, 128);
  k<<<1, 64>>>(d, 2000); k2<<<1, 64>>>(d, 2000);
  long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  const char* names[] = {"arrive+wait(complete)", "arrive", "try_wait(complete)", "ld.acquire.shared chain", "ld.volatile.shared chain",
                         "st.release.shared", "st.volatile.shared", "redux.or chain", "vote.all chain", "shfl chain",
                         "32 shfl + 32-step chain", "smem flag ping-pong round trip", "mbarrier ping-pong round trip"};
  for (int i = 0; i < 13; ++i) printf("%-34s %lld cycles\n", names[i], h[i]);
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
