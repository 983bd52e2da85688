// What the memory system gives a gather of 1 KB runs: every warp reads 1 KB pixels (two LDG.128 per lane) of a 713 MB map at
// random or RoI-like clustered positions, `batch` pixels in flight per warp, sums them and streams one 1 KB result per 10
// pixels -- the access pattern of the RoIAlign tap-list forward without any of its arithmetic.  Prints DRAM-side GB/s
// (bytes of distinct pixels read + bytes written) for 4 / 5 / 6 resident CTAs of 8 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather1k_bw gather1k_bw.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

template <int BATCH, int OCC>
__global__ void __launch_bounds__(256, OCC) k(const float4* __restrict__ map, const unsigned* __restrict__ idx, int per_warp,
                                              float4* __restrict__ out) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const unsigned* my = idx + (size_t)warp * per_warp;
  float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
  for (int i = 0; i < per_warp; i += BATCH) {
    float4 v[BATCH], u[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const unsigned p = __ldg(my + i + j);
      v[j] = __ldg(map + (size_t)p * 64 + lane);
      u[j] = __ldg(map + (size_t)p * 64 + 32 + lane);
    }
#pragma unroll
    for (int j = 0; j < BATCH; ++j) { a0.x += v[j].x; a0.y += v[j].y; a0.z += v[j].z; a0.w += v[j].w; a1.x += u[j].x; a1.y += u[j].y; a1.z += u[j].z; a1.w += u[j].w; }
    if ((i / BATCH) % (10 / BATCH + 1) == 0) { __stcs(out + ((size_t)warp * per_warp + i) / 10 * 64 + lane, a0); __stcs(out + ((size_t)warp * per_warp + i) / 10 * 64 + 32 + lane, a1); }
  }
  if (a0.x == 123.456f) out[0] = a0;
}

template <int BATCH, int OCC>
static void run(const char* name, const float4* map, const unsigned* idx, int warps, int per_warp, float4* out, double distinct_bytes) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = warps / 8;
  k<BATCH, OCC><<<grid, 256>>>(map, idx, per_warp, out);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaMemsetAsync((void*)out, 0, 256 << 20);          // evict L2 with 256 MB of writes into the output area
    cudaEventRecord(e0);
    k<BATCH, OCC><<<grid, 256>>>(map, idx, per_warp, out);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  const double wr = (double)warps * per_warp / 10 * 1024;
  printf("%-34s %7.3f ms   %7.1f GB/s (distinct pixels read + written)   %7.1f GB/s of loads issued\n", name, best,
         (distinct_bytes + wr) / best / 1e6, ((double)warps * per_warp * 1024 + wr) / best / 1e6);
}

int main() {
  const size_t n_pix = 696320;                          // 8 x (256^2 + 128^2 + 64^2 + 32^2): the C3 maps
  const int warps = 4096 * 8, per_warp = 60;           // 4096 RoIs x 8 warps x ~60 merged taps
  float4 *map, *out; unsigned* idx;
  cudaMalloc(&map, n_pix * 1024); cudaMalloc(&out, 512 << 20); cudaMalloc(&idx, (size_t)warps * per_warp * 4);
  cudaMemset(map, 0, n_pix * 1024);
  std::vector<unsigned> h((size_t)warps * per_warp);
  srand(1);
  // (a) uniformly random pixels
  for (auto& x : h) x = (unsigned)(((size_t)rand() * 32768 + rand()) % n_pix);
  cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  { std::vector<unsigned> s(h); std::sort(s.begin(), s.end()); const double d = (double)(std::unique(s.begin(), s.end()) - s.begin()) * 1024;
    run<4, 4>("random, 4 in flight, 4 CTAs/SM", map, idx, warps, per_warp, out, d);
    run<8, 4>("random, 8 in flight, 4 CTAs/SM", map, idx, warps, per_warp, out, d);
    run<4, 6>("random, 4 in flight, 6 CTAs/SM", map, idx, warps, per_warp, out, d); }
  // (b) RoI-like: a CTA's 8 warps walk a 14 x 9 pixel patch of one 256-wide image (neighbouring warps share pixels)
  for (int c = 0; c < warps / 8; ++c) {
    const unsigned img = rand() % 8, x0 = rand() % 240, y0 = rand() % 240;
    for (int w = 0; w < 8; ++w)
      for (int i = 0; i < per_warp; ++i) {
        const unsigned px = x0 + (i % 10) + (w & 1) * 4, py = y0 + (w >> 1) * 2 + (i / 10) % 3 + (i / 30);
        h[((size_t)c * 8 + w) * per_warp + i] = img * 65536 + py * 256 + px;
      }
  }
  cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  { std::vector<unsigned> s(h); std::sort(s.begin(), s.end()); const double d = (double)(std::unique(s.begin(), s.end()) - s.begin()) * 1024;
    run<4, 4>("RoI-like, 4 in flight, 4 CTAs/SM", map, idx, warps, per_warp, out, d);
    run<8, 4>("RoI-like, 8 in flight, 4 CTAs/SM", map, idx, warps, per_warp, out, d); }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
