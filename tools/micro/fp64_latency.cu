// Dependent-chain latencies of the FP64 operations the dense path's critical paths are made of
// (one warp, one chain): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int iters) {
  double x = a + threadIdx.x * 1e-9;
  __shared__ double sh[64];
  sh[threadIdx.x] = b;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (OP == 0) x = __fma_rn(x, b, a);
    if (OP == 1) x = __dadd_rn(x, b);
    if (OP == 2) x = __dmul_rn(x, b);
    if (OP == 3) x = x / b;
    if (OP == 4) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    if (OP == 5) { sh[threadIdx.x] = x; __syncwarp(); x = sh[(threadIdx.x + 1) & 31]; __syncwarp(); }
    if (OP == 6) x = __fma_rn(x, sh[(threadIdx.x + i) & 31], a);
    if (OP == 7) { double inv = 1.0 / b; double q = x * inv; double r = __fma_rn(-b, q, x); x = __fma_rn(r, inv, q); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
template <int OP>
__global__ void barrier_cost(long long* cyc, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA", "DADD", "DMUL", "DDIV (x / b)", "SHFL(double)", "smem store+load", "DFMA with LDS operand", "div via reciprocal (rcp hoisted by compiler?)"};
  const int iters = 4096;
  long long h;
#define RUN(OP) chain<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, iters); chain<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-48s %.1f cycles\n", names[OP], (double)h / iters);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
  for (int nt : {32, 128, 256}) {
    barrier_cost<0><<<1, nt>>>(cyc, iters); barrier_cost<0><<<1, nt>>>(cyc, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads, %3d threads                       %.1f cycles\n", nt, (double)h / iters);
  }
  return 0;
}
