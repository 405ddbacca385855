// ubench_tma_stream.cu — what the TMA load path of the recall filter (recall_tc.cu) can deliver on sm_100a, without any MMA:
// a persistent CTA per SM streams its tiles of a bf16 [rows][dim] matrix through a ring of shared-memory stages
// (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of BOX_ROWS x 64 bf16, dim / 64 boxes per stage); a consumer thread waits
// for each stage, optionally idles `delay` cycles (the MMA time of the stage) and hands it back.
// Question behind it: the 256-queries-per-pass filter at dim 128 (C5 shard shape: 2 stages of 64 KiB next to 64 KiB of
// query operands) ran with the tensor pipe 51 % busy and DRAM at 49 % — is that the load path's limit at that ring depth?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/ubench_tma_stream ubench_tma_stream.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(0x12F0000000000000ull)
      : "memory");
}

struct Cfg {
  int box_rows;      // rows per TMA box
  int stage_rows;    // rows per stage (a multiple of box_rows)
  int subs;          // dim / 64: boxes side by side
  int stages;
  int delay;         // consumer idle cycles per stage
  uint32_t n_stage_tiles;   // stages' worth of rows in the matrix
};

__global__ void __launch_bounds__(64, 1) stream(const __grid_constant__ CUtensorMap map, const Cfg c, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16], empty[16];
  const uint32_t stage_bytes = (uint32_t)c.stage_rows * 128u * (uint32_t)c.subs;
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t mine = (c.n_stage_tiles > blockIdx.x) ? (c.n_stage_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < mine; ++i) {
      const uint32_t s = i % c.stages, ph = (i / c.stages) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      mbar_arrive_expect_tx(&full[s], stage_bytes);
      const int row0 = (int)((blockIdx.x + i * gridDim.x) * (uint32_t)c.stage_rows);
      uint8_t* dst = smem + (size_t)s * stage_bytes;
      for (int b = 0; b < c.stage_rows / c.box_rows; ++b)
        for (int sub = 0; sub < c.subs; ++sub)
          tma_load_2d(dst + (size_t)(sub * c.stage_rows + b * c.box_rows) * 128, &map, sub * 64, row0 + b * c.box_rows, &full[s]);
    }
  } else if (threadIdx.x == 32) {
    unsigned long long acc = 0;
    for (uint32_t i = 0; i < mine; ++i) {
      const uint32_t s = i % c.stages, ph = (i / c.stages) & 1u;
      mbar_wait(&full[s], ph);
      acc += *reinterpret_cast<const unsigned long long*>(smem + (size_t)s * stage_bytes);
      if (c.delay > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < c.delay) {}
      }
      mbar_arrive(&empty[s]);
    }
    sink[blockIdx.x] = acc;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  const size_t bytes = 3200000000ull;   // 12.5 M x 128 bf16 = 25 M x 64 bf16
  uint8_t* buf;
  unsigned long long* sink;
  cudaMalloc(&buf, bytes);
  cudaMemset(buf, 1, bytes);
  cudaMalloc(&sink, 148 * 8);
  cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Run { const char* name; int dim, box_rows, stage_rows, stages, delay; CUtensorMapL2promotion promo; };
  const CUtensorMapL2promotion P256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B, P128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  const Run runs[] = {
      {"dim64  box256 stage32K x5 (c4, 64 q/pass)", 64, 256, 256, 5, 0, P256},
      {"dim64  box256 stage32K x4 (old 256 q/pass)", 64, 256, 256, 4, 0, P256},
      {"dim64  box128 stage16K x8 (new 256 q/pass)", 64, 128, 128, 8, 0, P256},
      {"dim128 box256 stage64K x2 (old c5)", 128, 256, 256, 2, 0, P256},
      {"dim128 box256 stage64K x2 delay 2400", 128, 256, 256, 2, 2400, P256},
      {"dim128 box128 stage32K x4 (new c5)", 128, 128, 128, 4, 0, P256},
      {"dim128 box128 stage32K x4 delay 1200", 128, 128, 128, 4, 1200, P256},
      {"dim128 box256 stage64K x3", 128, 256, 256, 3, 0, P256},
      {"dim128 box256 stage64K x3 delay 2400", 128, 256, 256, 3, 2400, P256},
      {"dim128 box128 stage32K x6", 128, 128, 128, 6, 0, P256},
      {"dim128 box128 stage32K x6 delay 1200", 128, 128, 128, 6, 1200, P256},
      {"dim128 box64  stage16K x8", 128, 64, 64, 8, 0, P256},
      {"dim128 box64  stage16K x8 delay 600", 128, 64, 64, 8, 600, P256},
      {"dim128 box128 stage32K x4 promo128", 128, 128, 128, 4, 0, P128},
      {"dim128 box256 stage64K x2 promo128", 128, 256, 256, 2, 0, P128},
      {"dim64  box256 stage32K x5 promo128", 64, 256, 256, 5, 0, P128},
  };
  for (const Run& r : runs) {
    const uint64_t rows = bytes / ((size_t)r.dim * 2);
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)r.dim, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)r.dim * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)r.box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, r.promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("%s: encode failed %d\n", r.name, (int)cr); continue; }
    Cfg c{r.box_rows, r.stage_rows, r.dim / 64, r.stages, r.delay, (uint32_t)(rows / r.stage_rows)};
    const size_t smem = (size_t)r.stages * r.stage_rows * 128 * (r.dim / 64);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      stream<<<148, 64, smem>>>(map, c, sink);
      cudaEventRecord(e1);
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { printf("%s: %s\n", r.name, cudaGetErrorString(e)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-46s in flight <= %3zu KiB/SM  %.3f ms  %7.1f GB/s\n", r.name, smem / 1024, best, bytes / best * 1e-6);
  }
  return 0;
}
