// Issue-rate probe for the packed FP32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2: two FP32 lanes per
// instruction) against their scalar forms and against a mix with FMNMX (ALU pipe), the instruction mix of the
// rotated-overlap arithmetic.  Prints lane-operations per clock per SM for each variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_tput f32x2_tput.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;

// VARIANT 0: scalar FFMA, 1: FFMA2, 2: FADD2, 3: FMUL2, 4: scalar FFMA x2 + FMNMX (2:1), 5: FFMA2 + FMNMX x1 (1:1 instr,
// 2:1 lane-ops) 6: FFMA2 + 2 FMNMX (the rect mix: 98.5 fma-pipe lane-ops : 31 fmnmx per pair ~ 3:1), 7: FMNMX alone
template <int V>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
  float2 a[kChains];
  float m[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) { a[i] = make_float2(seed + i, seed - i); m[i] = seed * i; }
  const float2 x = make_float2(1.0001f, 0.9999f), y = make_float2(1e-3f, -1e-3f);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (V == 0) { a[i].x = fmaf(a[i].x, x.x, y.x); a[i].y = fmaf(a[i].y, x.y, y.y); }
      if (V == 1) a[i] = __ffma2_rn(a[i], x, y);
      if (V == 2) a[i] = __fadd2_rn(a[i], y);
      if (V == 3) a[i] = __fmul2_rn(a[i], x);
      if (V == 4) { a[i].x = fmaf(a[i].x, x.x, y.x); a[i].y = fmaf(a[i].y, x.y, y.y); m[i] = fminf(m[i], a[i].x); }
      if (V == 5) { a[i] = __ffma2_rn(a[i], x, y); m[i] = fminf(m[i], a[i].x); }
      if (V == 6) { a[i] = __ffma2_rn(a[i], x, y); m[i] = fminf(m[i], a[i].x); m[i] = fmaxf(m[i], a[i].y); }
      if (V == 7) { m[i] = fminf(m[i], a[i].x); a[i].x = fmaxf(m[i], a[i].y); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += a[i].x + a[i].y + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
static void run(const char* name, double fma_lane_ops, double alu_lane_ops, double instrs) {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  const int grid = sms * 8;
  cudaMalloc(&out, (size_t)grid * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<grid, 256>>>(out, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<V><<<grid, 256>>>(out, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double per = (double)grid * 256 * kIters * kChains * 5;     // loop bodies executed
  const double clocks = ms * 1e-3 * khz * 1e3 * sms;                // SM clocks at the nominal max clock
  printf("%-34s %7.3f ms  fma-pipe lane-ops/clk/SM %6.1f  alu lane-ops/clk/SM %6.1f  warp-instr/clk/SM %5.2f\n", name, ms / 5,
         per * fma_lane_ops / clocks, per * alu_lane_ops / clocks, per * instrs / 32 / clocks);
  cudaFree(out);
}

int main() {
  run<0>("scalar FFMA x2", 2, 0, 2);
  run<1>("FFMA2", 2, 0, 1);
  run<2>("FADD2", 2, 0, 1);
  run<3>("FMUL2", 2, 0, 1);
  run<4>("scalar FFMA x2 + FMNMX", 2, 1, 3);
  run<5>("FFMA2 + FMNMX", 2, 1, 2);
  run<6>("FFMA2 + 2 FMNMX", 2, 2, 3);
  run<7>("2 FMNMX", 0, 2, 2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
