// fp64 pipe microbenchmark for sm_100a: dependent-issue latency and per-SM throughput of DADD / DMUL / DFMA / DSETP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_dp tools/ubench_dp.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, int OP>
__global__ void k(double* out, long long* clk, double a, double b, int iters) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = a + c + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        if (OP == 0) x[c] = __dadd_rn(x[c], b);
        else if (OP == 1) x[c] = __dmul_rn(x[c], b);
        else if (OP == 2) x[c] = __fma_rn(x[c], b, a);
        else x[c] = (x[c] < b) ? a : __longlong_as_double(__double_as_longlong(x[c]) + 1);  // DSETP + integer
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int CHAINS, int OP>
void run(const char* name, int threads) {
  double* out; long long* clk;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 8);
  const int iters = 256;
  k<CHAINS, OP><<<1, threads>>>(out, clk, 1.0, 1.0000001, iters);
  k<CHAINS, OP><<<1, threads>>>(out, clk, 1.0, 1.0000001, iters);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double ops = (double)iters * 8 * CHAINS;
  printf("%-6s chains=%d threads=%4d: %7.2f clk per dependent step, %6.2f clk per warp-instr per SM\n", name, CHAINS, threads,
         (double)h / (iters * 8), (double)h / (ops * (threads / 32)));
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<1, 0>("DADD", 32); run<4, 0>("DADD", 32); run<8, 0>("DADD", 32);
  run<1, 0>("DADD", 128); run<4, 0>("DADD", 128); run<4, 0>("DADD", 512); run<8, 0>("DADD", 512);
  run<1, 1>("DMUL", 32); run<4, 1>("DMUL", 512);
  run<1, 2>("DFMA", 32); run<4, 2>("DFMA", 512); run<8, 2>("DFMA", 1024);
  run<1, 3>("DSETP", 32); run<4, 3>("DSETP", 512);
  return 0;
}
