// ubench_fma.cu — FFMA vs FFMA2 issue-rate probe on sm_100a (decides the recall scan's inner-loop instruction).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 av = make_float2(a, a * 1.0001f), bv = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {
        acc[i].x = fmaf(acc[i].x, av.x, bv.x);
        acc[i].y = fmaf(acc[i].y, av.y, bv.y);
      } else {
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(reinterpret_cast<unsigned long long&>(acc[i]))
            : "l"(reinterpret_cast<unsigned long long&>(av)), "l"(reinterpret_cast<unsigned long long&>(bv)));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int mode = 0; mode < 2; ++mode) {
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 0.999f, 0.001f);
        else k<1><<<148, warps * 32>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      double fma = 148.0 * warps * 32 * (double)iters * 32;
      printf("warps/SM=%2d mode=%s  %.3f ms  %.2f TFLOP/s\n", warps, mode ? "FFMA2" : "FFMA ", best, 2 * fma / best * 1e-9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
