// Microbenchmark: achieved HBM read bandwidth for random contiguous chunks of S bytes (one warp
// per chunk, coalesced 16-byte loads), the access pattern of the PGS record stream.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dram_chunk_bench dram_chunk_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k(const uint4* buf, size_t nchunks_total, int S16, int iters, unsigned long long* sink, int align16) {
  const int lane = threadIdx.x & 31;
  unsigned long long wid = (blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5));
  unsigned long long x = wid * 0x9E3779B97F4A7C15ull + 12345;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int it = 0; it < iters; it++) {
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    unsigned long long r = (x * 0x2545F4914F6CDD1Dull) % nchunks_total;
    const uint4* p = buf + r * (size_t)align16;
    for (int i = lane; i < S16; i += 32) {
      uint4 v = __ldcs(p + i);
      acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) sink[0] = 1;
}

int main(int argc, char** argv) {
  size_t bytes = (size_t)16 << 30;
  uint4* buf; unsigned long long* sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&sink, 8);
  cudaMemset(buf, 1, bytes);
  int sizes[] = {240, 480, 960, 1920, 3840, 7680, 15360, 61440, 245760};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int wps : {16, 32, 64}) {
    for (int S : sizes) {
      int S16 = S / 16;
      size_t nchunks = bytes / S;
      int iters = (int)((size_t)(48ull << 30) / ((size_t)148 * wps * S));
      if (iters < 4) iters = 4;
      dim3 grid(148 * wps / 8), block(256);
      k<<<grid, block>>>(buf, nchunks, S16, 2, sink, S16);
      cudaEventRecord(e0);
      k<<<grid, block>>>(buf, nchunks, S16, iters, sink, S16);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double gb = (double)148 * wps * iters * S / 1e9;
      printf("warps/SM %d chunk %6d B: %.1f GB/s\n", wps, S, gb / (ms * 1e-3));
    }
  }
  return 0;
}
