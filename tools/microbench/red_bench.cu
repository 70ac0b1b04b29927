// Microbenchmark: what bounds the per-cell accumulation and the cell gather of the ray march on B200?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu
// Reports giga-ops/s for scattered vs line-cooperative FP64 REDs and random 32-B gathers,
// for an L2-resident table (64^3 cells) and an HBM-resident one (256^3 cells).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ uint64_t rnd_cell(uint32_t tid, uint32_t it, uint64_t ncell) {
  uint32_t a = hash32(tid * 0x9E3779B9u + it), b = hash32(a ^ 0x85ebca6bu);
  return ((((uint64_t)a << 32) | b) % ncell);
}

template <int NRED>
__global__ void scatter_red(double *acc, uint64_t ncell, int iters, int stride) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    double *a = acc + rnd_cell(tid, it, ncell) * stride;
#pragma unroll
    for (int k = 0; k < NRED; ++k) atomicAdd(a + k, 1.0 + k);
  }
}

// 32 lanes hold 32 cells; 16 rounds, each RED instruction covers two full 128-B lines
__global__ void coop_red16(double *acc, uint64_t ncell, int iters) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, half = lane >> 4, j = lane & 15;
  for (int it = 0; it < iters; ++it) {
    uint64_t cell = rnd_cell(tid, it, ncell);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      uint64_t c = __shfl_sync(0xffffffffu, cell, 2 * k + half);
      atomicAdd(acc + c * 16 + j, 1.0 + j);
    }
  }
}

// same with 8 doubles per lane-group (4 lines per instruction), 8 rounds x 2
__global__ void coop_red8(double *acc, uint64_t ncell, int iters) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, q = lane >> 3, j = lane & 7;
  for (int it = 0; it < iters; ++it) {
    uint64_t cell = rnd_cell(tid, it, ncell);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint64_t c = __shfl_sync(0xffffffffu, cell, 4 * k + q);
      atomicAdd(acc + c * 16 + j, 1.0 + j);
      atomicAdd(acc + c * 16 + 8 + j, 1.0 + j);
    }
  }
}

// f32 x4 vector reductions: 4 instructions cover 16 floats (64 B) of one cell
__global__ void scatter_red_v4f32(float *acc, uint64_t ncell, int iters) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    float *a = acc + rnd_cell(tid, it, ncell) * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a + 4 * k), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
  }
}

__global__ void gather32(const double4 *cells, uint64_t ncell, int iters, double *out) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.;
  for (int it = 0; it < iters; ++it) {
    const double2 *p = reinterpret_cast<const double2 *>(cells + rnd_cell(tid, it, ncell));
    double2 a = __ldg(p), b = __ldg(p + 1);
    s += a.x + a.y + b.x + b.y;
  }
  if (s == 12345.678) out[0] = s;
}

// dependent gathers (each address depends on the previous value) with MLP ways per thread: latency-bound like a ray march
template <int MLP>
__global__ void gather32_dep(const double4 *cells, uint64_t ncell, int iters, double *out) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t c[MLP];
  for (int m = 0; m < MLP; ++m) c[m] = rnd_cell(tid, 1000 + m, ncell);
  double s = 0.;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int m = 0; m < MLP; ++m) {
      const double2 *p = reinterpret_cast<const double2 *>(cells + c[m]);
      double2 a = __ldg(p);
      s += a.x;
      c[m] = (c[m] + (uint64_t)(a.y) + rnd_cell(tid, it * MLP + m, ncell)) % ncell;
    }
  }
  if (s == 12345.678) out[0] = s;
}

template <class F> float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("device %s SMs %d L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
  const int bs = 256, grid = prop.multiProcessorCount * 8;
  const double nthreads = (double)bs * grid;
  for (uint64_t side : {64ull, 256ull}) {
    const uint64_t ncell = side * side * side;
    double *acc; cudaMalloc(&acc, ncell * 16 * sizeof(double)); cudaMemset(acc, 0, ncell * 16 * sizeof(double));
    double4 *cells; cudaMalloc(&cells, ncell * sizeof(double4)); cudaMemset(cells, 0, ncell * sizeof(double4));
    double *out; cudaMalloc(&out, 8);
    const int iters = 64;
    float ms;
    printf("== %llu^3 cells: acc[16] %.1f MB, cells %.1f MB\n", (unsigned long long)side, ncell * 128 / 1e6, ncell * 32 / 1e6);
    ms = time_ms([&] { scatter_red<16><<<grid, bs>>>(acc, ncell, iters, 16); });
    printf("scatter_red<16> (1 lane -> 16 REDs of one line): %.3f ms  %.2f Gcell/s  %.2f GRED/s\n", ms, nthreads * iters / ms / 1e6, 16 * nthreads * iters / ms / 1e6);
    ms = time_ms([&] { coop_red16<<<grid, bs>>>(acc, ncell, iters); });
    printf("coop_red16 (16 lanes per line)                 : %.3f ms  %.2f Gcell/s  %.2f GRED/s\n", ms, nthreads * iters / ms / 1e6, 16 * nthreads * iters / ms / 1e6);
    ms = time_ms([&] { coop_red8<<<grid, bs>>>(acc, ncell, iters); });
    printf("coop_red8 (8 lanes per half line)              : %.3f ms  %.2f Gcell/s  %.2f GRED/s\n", ms, nthreads * iters / ms / 1e6, 16 * nthreads * iters / ms / 1e6);
    ms = time_ms([&] { scatter_red<4><<<grid, bs>>>(acc, ncell, iters, 16); });
    printf("scatter_red<4>                                 : %.3f ms  %.2f Gcell/s  %.2f GRED/s\n", ms, nthreads * iters / ms / 1e6, 4 * nthreads * iters / ms / 1e6);
    ms = time_ms([&] { scatter_red<2><<<grid, bs>>>(acc, ncell, iters, 2); });
    printf("scatter_red<2> stride 2 (H-only layout)        : %.3f ms  %.2f Gcell/s  %.2f GRED/s\n", ms, nthreads * iters / ms / 1e6, 2 * nthreads * iters / ms / 1e6);
    ms = time_ms([&] { scatter_red<1><<<grid, bs>>>(acc, ncell, iters, 2); });
    printf("scatter_red<1> stride 2                        : %.3f ms  %.2f Gcell/s  %.2f GRED/s\n", ms, nthreads * iters / ms / 1e6, 1 * nthreads * iters / ms / 1e6);
    ms = time_ms([&] { scatter_red_v4f32<<<grid, bs>>>((float *)acc, ncell, iters); });
    printf("scatter red.v4.f32 x4 (16 floats)              : %.3f ms  %.2f Gcell/s\n", ms, nthreads * iters / ms / 1e6);
    ms = time_ms([&] { gather32<<<grid, bs>>>(cells, ncell, iters, out); });
    printf("gather32 independent                           : %.3f ms  %.2f Ggather/s  %.1f GB/s\n", ms, nthreads * iters / ms / 1e6, 32 * nthreads * iters / ms / 1e6);
    ms = time_ms([&] { gather32_dep<1><<<grid, bs>>>(cells, ncell, iters, out); });
    printf("gather32 dependent chain MLP=1 (8 CTA/SM x 256): %.3f ms  %.2f Ggather/s\n", ms, nthreads * iters / ms / 1e6);
    ms = time_ms([&] { gather32_dep<4><<<grid, bs>>>(cells, ncell, iters, out); });
    printf("gather32 dependent chain MLP=4                 : %.3f ms  %.2f Ggather/s\n", ms, 4 * nthreads * iters / ms / 1e6);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    cudaFree(acc); cudaFree(cells); cudaFree(out);
  }
  return 0;
}
