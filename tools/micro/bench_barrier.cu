// Grid-barrier microbenchmark on B200: us per barrier for three implementations, cooperative launch, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_barrier bench_barrier.cu && ./bench_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void bar_flip(unsigned int *bar) {   // arrive and poll on the same word (CG style)
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int nb = 1;
    if (blockIdx.x == 0) nb = 0x80000000u - (gridDim.x - 1);
    unsigned int old, cur;
    asm volatile("atom.add.release.gpu.u32 %0,[%1],%2;" : "=r"(old) : "l"(bar), "r"(nb) : "memory");
    do {
      asm volatile("ld.acquire.gpu.u32 %0,[%1];" : "=r"(cur) : "l"(bar) : "memory");
    } while (((old ^ cur) & 0x80000000u) == 0);
  }
  __syncthreads();
}

// arrive on bar[0]; the last arriver resets it and publishes a new generation in bar[32] (another 128 B line)
__device__ __forceinline__ void bar_split(unsigned int *bar, unsigned int &gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++gen;
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.u32 %0,[%1],%2;" : "=r"(old) : "l"(bar), "r"(1u) : "memory");
    if (old == gridDim.x - 1) {
      bar[0] = 0;
      asm volatile("st.release.gpu.u32 [%0],%1;" ::"l"(bar + 32), "r"(gen) : "memory");
    } else {
      unsigned int cur;
      do {
        asm volatile("ld.acquire.gpu.u32 %0,[%1];" : "=r"(cur) : "l"(bar + 32) : "memory");
      } while (cur != gen);
    }
  }
  __syncthreads();
}

template <int MODE>
__global__ void k(unsigned int *bar, int iters, double *sink) {
  cg::grid_group g = cg::this_grid();
  unsigned int gen = 0;
  double acc = 0;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) bar_flip(bar);
    else if (MODE == 1) bar_split(bar, gen);
    else g.sync();
    acc += i;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) sink[0] = acc;
}

template <int MODE>
void run(const char *name, int grid, int threads) {
  unsigned int *bar;
  double *sink;
  cudaMalloc(&bar, 1024);
  cudaMemset(bar, 0, 1024);
  cudaMalloc(&sink, 8);
  int iters = 2000;
  void *args[] = {&bar, &iters, &sink};
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(bar, 0, 1024);
    cudaEventRecord(a);
    cudaError_t e = cudaLaunchCooperativeKernel((void *)k<MODE>, dim3(grid), dim3(threads), args, 0, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep == 2) printf("%-12s grid %4d x %4d : %.3f us per barrier (%s)\n", name, grid, threads, 1e3 * ms / iters, cudaGetErrorString(e));
  }
  cudaFree(bar);
  cudaFree(sink);
}

int main() {
  int nsm;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  for (int threads : {1024, 512, 256}) {
    run<0>("flip", nsm, threads);
    run<1>("split", nsm, threads);
    run<2>("cg", nsm, threads);
  }
  run<0>("flip", 2 * nsm, 512);
  run<1>("split", 2 * nsm, 512);
  run<0>("flip", nsm / 2, 1024);
  run<1>("split", nsm / 2, 1024);
  return 0;
}
