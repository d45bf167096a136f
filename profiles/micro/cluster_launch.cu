// Launch + execution cost of small kernels on B200: plain grid vs thread-block cluster, with and without barriers.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cluster_launch cluster_launch.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void k_plain(int* out) { if (threadIdx.x == 0 && blockIdx.x == 0 && out) out[0] = 1; }
__global__ void __cluster_dims__(8, 1, 1) k_cluster(int* out) { if (threadIdx.x == 0 && blockIdx.x == 0 && out) out[0] = 1; }
__global__ void __cluster_dims__(8, 1, 1) k_cluster_sync(int* out, int nsync) {
  cg::cluster_group c = cg::this_cluster();
  for (int i = 0; i < nsync; ++i) c.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0 && out) out[0] = 1;
}
__global__ void k_coop(int* out, int nsync) {
  cg::grid_group g = cg::this_grid();
  for (int i = 0; i < nsync; ++i) g.sync();
  if (threadIdx.x == 0 && blockIdx.x == 0 && out) out[0] = 1;
}
__global__ void __cluster_dims__(8, 1, 1) k_cluster_smem(int* out) {
  __shared__ int big[10000];
  big[threadIdx.x] = threadIdx.x;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0 && out) out[0] = big[5];
}

template <class F> float timeit(F f, int n) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 20; ++i) f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < n; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms * 1000.f / n;
}
int main() {
  int* d; cudaMalloc(&d, 4);
  const int n = 2000;
  printf("plain   grid=8   x512: %.2f us\n", timeit([&] { k_plain<<<8, 512>>>(d); }, n));
  printf("plain   grid=128 x512: %.2f us\n", timeit([&] { k_plain<<<128, 512>>>(d); }, n));
  printf("cluster grid=8   x512: %.2f us\n", timeit([&] { k_cluster<<<8, 512>>>(d); }, n));
  printf("cluster grid=512 x512: %.2f us\n", timeit([&] { k_cluster<<<512, 512>>>(d); }, n));
  printf("cluster+40KB smem grid=8: %.2f us\n", timeit([&] { k_cluster_smem<<<8, 512>>>(d); }, n));
  for (int s : {1, 6, 20}) printf("cluster grid=8, %2d cluster.sync: %.2f us\n", s, timeit([&] { k_cluster_sync<<<8, 512>>>(d, s); }, n));
  for (int s : {0, 1, 3}) {
    int ns = s; void* args[] = {&d, &ns};
    printf("coop grid=128 x512, %d grid.sync: %.2f us\n", s, timeit([&] { cudaLaunchCooperativeKernel((void*)k_coop, dim3(128), dim3(512), args, 0, 0); }, n));
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
