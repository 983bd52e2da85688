// capi.cu -- library-level entry points: errors, launch accounting, kernel timing,
// and the FP32 FFMA micro-benchmark used as the ALU roofline denominator.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace aidet {

static thread_local char g_err[512] = "";
static long long g_launches = 0;
static int g_prof_level = 0;          // 0 off, 1 per-op kernel timing, 2 + phase stamps of the fused NMS kernel
static std::mutex g_mu;
struct EvPair { cudaEvent_t e0, e1; };
static std::vector<EvPair> g_pending[PROF_KINDS];
static double g_ms[PROF_KINDS] = {0, 0, 0, 0};
static long long g_cnt[PROF_KINDS] = {0, 0, 0, 0};

void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return AIDET_ECUDA;
}
void count_launch(int n) { std::lock_guard<std::mutex> l(g_mu); g_launches += n; }

int prof_level() { return g_prof_level; }

ProfScope::ProfScope(int kind_, cudaStream_t s_) : kind(kind_), s(s_), on(g_prof_level > 0), e0(nullptr), e1(nullptr) {
  if (!on) return;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { on = false; return; }
  cudaEventRecord(e0, s);
}
ProfScope::~ProfScope() {
  if (!on) return;
  cudaEventRecord(e1, s);
  std::lock_guard<std::mutex> l(g_mu);
  g_pending[kind].push_back({e0, e1});
}

__global__ void ffma_peak_kernel(float* sink, int iters, float a, float b);

int set_device(int device) {
  int cur = -1;
  AIDET_CUDA(cudaGetDevice(&cur));
  if (cur != device) AIDET_CUDA(cudaSetDevice(device));
  // First use on a device: have the runtime load this library's module through an attribute query.  Left to the first
  // cudaLaunchKernel, the runtime probes the not-yet-loaded kernel with cuKernelGetFunction, gets
  // CUDA_ERROR_INVALID_HANDLE, loads the module and retries -- harmless and invisible to the caller, but
  // compute-sanitizer reports the internal error (round-1 log: one error under assign_fused<HbbKind>, the first launch of
  // that test process; processes whose first call went through cudaFuncSetAttribute showed none).
  static bool primed[64];
  if (device >= 0 && device < 64 && !primed[device]) {
    cudaFuncAttributes attr;
    (void)cudaFuncGetAttributes(&attr, ffma_peak_kernel);
    (void)cudaGetLastError();
    primed[device] = true;
  }
  return AIDET_OK;
}
int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int n = 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) n = 148;
  if (device >= 0 && device < 64) cached[device] = n;
  return n;
}

// 8 independent FFMA chains per thread; 2 flop per FFMA.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* sink, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) sink[0] = s;
}

}  // namespace aidet

using namespace aidet;

extern "C" {

const char* aidet_last_error(void) { return g_err; }
int aidet_version(void) { return 100; }
long long aidet_launch_count(void) { std::lock_guard<std::mutex> l(g_mu); return g_launches; }

int aidet_prof_enable(int enable) { g_prof_level = enable < 0 ? 0 : enable; return AIDET_OK; }

int aidet_prof_read(int kind, double* ms_total, long long* launches, int reset) {
  AIDET_REQUIRE(kind >= 0 && kind < PROF_KINDS, "aidet_prof_read: kind %d out of range", kind);
  std::vector<EvPair> todo;
  { std::lock_guard<std::mutex> l(g_mu); todo.swap(g_pending[kind]); }
  for (auto& p : todo) {
    AIDET_CUDA(cudaEventSynchronize(p.e1));
    float ms = 0.f;
    AIDET_CUDA(cudaEventElapsedTime(&ms, p.e0, p.e1));
    g_ms[kind] += ms; g_cnt[kind] += 1;
    cudaEventDestroy(p.e0); cudaEventDestroy(p.e1);
  }
  if (ms_total) *ms_total = g_ms[kind];
  if (launches) *launches = g_cnt[kind];
  if (reset) { g_ms[kind] = 0; g_cnt[kind] = 0; }
  return AIDET_OK;
}

int aidet_ffma_peak(int device, int iters, double* tflops_out, void* stream) {
  AIDET_REQUIRE(tflops_out && iters > 0, "aidet_ffma_peak: bad arguments");
  if (int rc = set_device(device)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  float* sink = nullptr;
  AIDET_CUDA(cudaMalloc(&sink, sizeof(float)));
  int blocks = sm_count(device) * 8;
  cudaEvent_t e0, e1;
  AIDET_CUDA(cudaEventCreate(&e0)); AIDET_CUDA(cudaEventCreate(&e1));
  ffma_peak_kernel<<<blocks, 256, 0, s>>>(sink, 16, 0.999f, 1e-3f);   // warm-up
  AIDET_CUDA(cudaEventRecord(e0, s));
  ffma_peak_kernel<<<blocks, 256, 0, s>>>(sink, iters, 0.999f, 1e-3f);
  AIDET_CUDA(cudaEventRecord(e1, s));
  count_launch(2);
  AIDET_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  AIDET_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  double flops = 2.0 * 128.0 * (double)iters * 256.0 * (double)blocks;
  *tflops_out = flops / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
  AIDET_CUDA(cudaGetLastError());
  return AIDET_OK;
}

}  // extern "C"
