// FP64 pipe microbenchmark for the roofline denominators (MEASURED_PEAKS.json has no FP64 entry):
//   DFMA  : 8 independent FMA chains per thread
//   DMMA  : mma.sync.aligned.m8n8k4.row.col.f64, 8 independent accumulator tiles per warp
// Prints TFLOP/s for both and the per-SM-sub-partition issue interval of one DMMA.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 1 << 16;
  {
    const int blocks = sms * 2, threads = 512;
    float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double flops = 2.0 * 8 * iters * (double)blocks * threads;
    printf("DFMA: %.2f TFLOP/s (%d SMs, %.3f ms)\n", flops / ms / 1e9, sms, ms);
  }
  for (int warps = 1; warps <= 8; warps *= 2) {
    const int blocks = sms, threads = 128 * warps;  // `warps` warps per SM sub-partition
    float ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters / 8, 1.0000001, 1e-9); });
    double n_mma = 8.0 * (iters / 8) * (double)blocks * (threads / 32);
    double flops = n_mma * 2 * 8 * 8 * 4;
    double cyc = ms * 1e-3 * clk_khz * 1e3;  // at the reported clock
    printf("DMMA m8n8k4: %d warps/SMSP, 8 acc: %.2f TFLOP/s, %.1f cycles per DMMA per SMSP (clock %d MHz, %.3f ms)\n", warps,
           flops / ms / 1e9, cyc / (8.0 * (iters / 8) * warps), clk_khz / 1000, ms);
  }
  {
    const int blocks = sms, threads = 128;
    float ms = time_ms([&] { dmma_kernel<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("DMMA dependent chain (1 warp/SMSP, 1 acc): %.1f cycles latency\n", cyc / iters);
  }
  {
    float ms = time_ms([&] { dfma_kernel<<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
    double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("DFMA 1 warp, 8 chains: %.2f cycles per DFMA issue\n", cyc / (8.0 * iters));
  }
  return 0;
}
