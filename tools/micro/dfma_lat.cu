// Dev micro-benchmark: FP64 FMA latency / throughput per SM sub-partition on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dfma_lat tools/micro/dfma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, double a, double b, int iters) {
  double x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int warps_per_block) {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  int iters = 2000;
  k<ILP><<<1, 32 * warps_per_block>>>(out, cyc, 0.999, 1e-3, iters);
  k<ILP><<<1, 32 * warps_per_block>>>(out, cyc, 0.999, 1e-3, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = double(h) / (double(iters) * 16 * ILP);
  printf("warps/SM %2d (per SMSP %d) ILP %d: %.2f cycles per DFMA per warp -> %.2f DFMA warp-instr/cycle/SMSP\n",
         warps_per_block, (warps_per_block + 3) / 4, ILP, per, (warps_per_block / 4.0) / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(1); run<2>(1); run<4>(1); run<8>(1);
  run<1>(4); run<1>(8); run<1>(16); run<2>(16); run<4>(16); run<1>(32); run<2>(32);
  return 0;
}
