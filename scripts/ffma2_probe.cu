// Micro-benchmark: what does the packed FP32 FMA of sm_100 (PTX fma.rn.f32x2, SASS FFMA2) buy?
//   (a) FFMA alone, (b) FFMA2 alone: FMA lanes per clock per SM                      -> same FP32 peak or not
//   (c) FFMA  + independent ALU work (FMNMX) in a 1:1 mix, (d) FFMA2 + the same     -> issue slots freed by the packed form
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ffma2_probe scripts/ffma2_probe.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, float m) {
  float x[8], y[8];
  uint64_t z[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] = threadIdx.x + j; y[j] = threadIdx.x * 0.5f + j; z[j] = pk2(x[j], y[j]); }
  const uint64_t a2 = pk2(a, a), b2 = pk2(b, b);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (MODE == 0 || MODE == 2) x[j] = fmaf(x[j], a, b);
        if (MODE == 1 || MODE == 3) z[j] = fma2(z[j], a2, b2);
        if (MODE >= 2) y[j] = (r & 1) ? fminf(y[(j + 1) & 7], y[(j + 3) & 7]) : fmaxf(y[(j + 1) & 7], y[(j + 3) & 7]);   // one FMNMX per FMA instruction
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { float lo, hi; up2(z[j], lo, hi); s += x[j] + y[j] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, float* out, int sms, double fma_per_inst) {
  const int blocks = sms * 8, iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) k<MODE><<<blocks, 256>>>(out, iters, 0.999999f, 1e-7f, 3.0f);
  cudaEventRecord(e0);
  for (int w = 0; w < 4; ++w) k<MODE><<<blocks, 256>>>(out, iters, 0.999999f, 1e-7f, 3.0f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double inst = 4.0 * blocks * 256.0 * iters * 64.0;   // FMA-type thread instructions
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %8.3f ms  %7.2f TFLOP/s  %6.1f FMA lanes/clk/SM (at %d MHz nominal)\n", name, ms / 4, inst * fma_per_inst * 2 / (ms * 1e-3) / 1e12,
         inst * fma_per_inst / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000);
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  run<0>("FFMA", out, sms, 1);
  run<1>("FFMA2", out, sms, 2);
  run<2>("FFMA + FMNMX 1:1", out, sms, 1);
  run<3>("FFMA2 + FMNMX 1:1", out, sms, 2);
  run<4>("FMNMX alone (per inst)", out, sms, 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
