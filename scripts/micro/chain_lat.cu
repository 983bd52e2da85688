// Stand-alone replica of the fused NMS scan's chain loop (warp 0) with synthetic data: cycles per 32-row block.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../aidet_b200/csrc/common.cuh"
using namespace aidet;
__device__ __forceinline__ uint32_t ld_acq(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory"); return v; }

template <int VARIANT>
__global__ void k(long long* out, int nhw, uint32_t density) {
  extern __shared__ uint32_t panel[];                  // nhw * 128
  __shared__ uint32_t s_removed[256], s_keep[256], s_hprog[8];
  __shared__ __align__(8) uint64_t bar_k[32];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) { mbar_init(&bar_k[threadIdx.x], 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < nhw * 128; i += blockDim.x) { uint32_t h = i * 2654435761u; h ^= h >> 13; h *= 0x5bd1e995u; panel[i] = ((h >> 8) & density) ? 0u : (1u << ((h >> 3) & 31)) | (1u << ((h >> 17) & 31)); }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_removed[i] = 0; s_keep[i] = 0; }
  if (threadIdx.x < 8) s_hprog[threadIdx.x] = 100000;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int ng = nhw * 32;
  uint32_t c1 = 0, c2 = 0, c3 = 0, d0, d1, d2, d3;
  auto fetch = [&](int b, const uint32_t* blk, uint32_t& e0, uint32_t& e1, uint32_t& e2, uint32_t& e3) {
    e0 = e1 = e2 = e3 = 0u;
    if (lane < ng - 32 * b) { e0 = blk[0]; if (b + 1 < nhw) e1 = blk[32]; if (b + 2 < nhw) e2 = blk[64]; if (b + 3 < nhw) e3 = blk[96]; }
  };
  fetch(0, panel + lane, d0, d1, d2, d3);
  const long long t0 = clock64();
  uint32_t total = 0;
  for (int b = 0; b < nhw; ++b) {
    if (VARIANT != 1 && b >= 4) { const uint32_t need = (uint32_t)(b - 3); while (!__all_sync(0xffffffffu, lane >= 6 || ld_acq(&s_hprog[lane & 7]) >= need)) {} }
    const int rows_b = min(32, ng - 32 * b);
    uint32_t cur = ld_acq(&s_removed[b]) | c1;
    if (rows_b < 32) cur |= ~0u << rows_b;
    uint32_t keep = 0, n1 = 0, n2 = 0, n3 = 0, e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    if (b + 1 < nhw) fetch(b + 1, panel + (b + 1) * 128 + lane, e0, e1, e2, e3);
    if (cur != ~0u) {
      if (VARIANT == 2) {
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) { const uint32_t dk = __shfl_sync(0xffffffffu, d0, kk); if (!((cur >> kk) & 1u)) { keep |= 1u << kk; cur |= dk; } }
      } else {
#pragma unroll
        for (int kk = 0; kk < 32; kk += 2) {
          const uint32_t da = __shfl_sync(0xffffffffu, d0, kk), db = __shfl_sync(0xffffffffu, d0, kk + 1);
          const bool ka = !(cur & (1u << kk));
          const bool kb = ka ? !((cur | da) & (2u << kk)) : !(cur & (2u << kk));
          cur |= (ka ? da : 0u) | (kb ? db : 0u);
          keep |= (ka ? (1u << kk) : 0u) | (kb ? (2u << kk) : 0u);
        }
      }
      const bool kept = (keep >> lane) & 1u;
      if (VARIANT != 3) { n1 = __reduce_or_sync(0xffffffffu, kept ? d1 : 0u); n2 = __reduce_or_sync(0xffffffffu, kept ? d2 : 0u); n3 = __reduce_or_sync(0xffffffffu, kept ? d3 : 0u); }
    }
    if (VARIANT != 4) { if (lane == 0) { s_keep[b] = keep; mbar_arrive(&bar_k[b % 32]); } }
    total += __popc(keep);
    c1 = c2 | n1; c2 = c3 | n2; c3 = n3;
    d0 = e0; d1 = e1; d2 = e2; d3 = e3;
  }
  const long long t1 = clock64();
  if (lane == 0) { out[VARIANT * 2] = (t1 - t0) / nhw; out[VARIANT * 2 + 1] = total; }
}

int main() {
  long long* d; cudaMalloc(&d, 32 * 8); cudaMemset(d, 0, 256);
  const int nhw = 63;
  k<0><<<1, 256, nhw * 512>>>(d, nhw, 3);
  k<1><<<1, 256, nhw * 512>>>(d, nhw, 3);
  k<2><<<1, 256, nhw * 512>>>(d, nhw, 3);
  k<3><<<1, 256, nhw * 512>>>(d, nhw, 3);
  k<4><<<1, 256, nhw * 512>>>(d, nhw, 3);
  long long h[32]; cudaMemcpy(h, d, 256, cudaMemcpyDeviceToHost);
  const char* names[] = {"full loop (2-bit chain)", "no helper poll", "1-bit chain", "no redux", "no publish"};
  for (int i = 0; i < 5; ++i) printf("%-28s %lld cycles/block, kept %lld\n", names[i], h[2 * i], h[2 * i + 1]);
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
