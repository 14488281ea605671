// Issue-rate microbenchmark: FFMA, FFMA2 (fma.rn.f32x2) and FHFMA.BF16 (fma.rn.f32.bf16, bf16 x bf16 + fp32) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CH = 16;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const uint32_t* in) {
  float acc[CH];
  uint64_t acc2[CH / 2];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = i;
#pragma unroll
  for (int i = 0; i < CH / 2; ++i) acc2[i] = i;
  uint32_t x = in[threadIdx.x & 31], w = in[32 + (threadIdx.x & 31)];
  float xf = __uint_as_float(x), wf = __uint_as_float(w);
  uint64_t x2, w2;
  asm("mov.b64 %0, {%1,%2};" : "=l"(x2) : "r"(x), "r"(w));
  asm("mov.b64 %0, {%1,%2};" : "=l"(w2) : "r"(w), "r"(x));
  unsigned short xl, xh, wl, wh;
  asm("mov.b32 {%0,%1}, %2;" : "=h"(xl), "=h"(xh) : "r"(x));
  asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(w));
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < CH; ++i) acc[i] = fmaf(xf, wf, acc[i]);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i]) : "l"(x2), "l"(w2));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < CH; i += 2) {
        asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(xl), "h"(wl));
        asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc[i + 1]) : "h"(xh), "h"(wh));
      }
    } else {  // FHFMA interleaved with FFMA2 (do they share a pipe?)
#pragma unroll
      for (int i = 0; i < CH / 2; ++i) {
        asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(xl), "h"(wl));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i]) : "l"(x2), "l"(w2));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < CH / 2; ++i) s += static_cast<float>(acc2[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_instr, int instr_per_iter) {
  float* out; uint32_t* in;
  const int blocks = 148 * 8;
  cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&in, 64 * 4);
  cudaMemset(in, 0x3f, 64 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, in);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, in);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double instr = double(blocks) * 256 / 32 * ITERS * instr_per_iter;  // warp instructions
  const double per_clk_sm = instr / (ms * 1e-3) / 148 / 1.965e9;
  printf("%-28s %8.3f ms  warp-instr/clk/SM %.2f  TFMA/s %.1f\n", name, ms, per_clk_sm, instr * 32 * fma_per_instr / (ms * 1e-3) / 1e12);
  cudaFree(out); cudaFree(in);
}

int main() {
  run<0>("FFMA", 1, CH);
  run<1>("FFMA2 (f32x2)", 2, CH / 2);
  run<2>("FHFMA.BF16 (f32.bf16)", 1, CH);
  run<3>("FHFMA + FFMA2 interleaved", 1.5, CH);
  return 0;
}
