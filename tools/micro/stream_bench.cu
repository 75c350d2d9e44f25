// Microbenchmark of the PGS record-stream pattern: one-warp CTAs (11 per SM), each streaming its own
// cyclic region in 4 KB TMA bulk copies (one in flight), with a delay loop standing in for the
// stage compute and an optional L2 prefetch ahead.  Prints achieved GB/s and ns per step.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) k(const char* buf, size_t region, int chunk, int steps, int delay, int pf_spans, int pf_kind, int wr,
                                        unsigned long long* sink, int smem_pad) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x;
  const unsigned bar = s32(sm), dst = s32(sm + 128);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const char* base = buf + (size_t)blockIdx.x * region;
  unsigned parity = 0;
  size_t pos = 0, pf = 0;
  unsigned long long acc = 0;
  auto issue = [&](size_t p) {
    if (lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(__cvta_generic_to_global(base + p)), "r"(chunk), "r"(bar) : "memory");
    }
  };
  issue(0);
  for (int t = 0; t < steps; t++) {
    asm volatile("{\n\t.reg .pred p;\nLW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra LD;\n\tbra LW;\nLD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
    parity ^= 1u;
    acc += *reinterpret_cast<const unsigned long long*>(sm + 128 + lane * 8);
    __syncwarp();
    size_t nx = pos + chunk;
    if (nx + chunk > region) nx = 0;
    issue(nx);
    if (pf_spans > 0) {
      long long lead = (long long)pf - (long long)nx;
      if (lead < 0) lead += region;
      for (int q = 0; q < 2; q++) {
        if (lead < (long long)pf_spans * 4096) {
          if (pf_kind == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + pf + lane * 128));
          else if (lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(base + pf)), "r"(4096) : "memory");
          pf += 4096; lead += 4096;
          if (pf + 4096 > region) pf = 0;
        }
      }
    }
    if (wr == 2) {   // compact multiplier array: 32 B per record, contiguous per step, in a separate region
      const int nrec = chunk / 224;
      char* lbase = (char*)buf + (size_t)gridDim.x * region + (size_t)blockIdx.x * (region / 7) + (pos / 224) * 32;
      if (lane < nrec) asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(lbase + lane * 32), "d"((double)t) : "memory");
      if (lane == 0) {   // and read next step's multipliers with a second bulk copy (same barrier would be used in the solver)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(lbase + nrec * 32));
      }
    } else if (wr) {   // write back 32 B per 224 B of the chunk just consumed (the multiplier sectors)
      for (int i = lane; i * 224 + 224 <= chunk; i += 32) {
        double* p = (double*)(base + pos + i * 224 + 192);
        asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"((double)t) : "memory");
      }
    }
    const long long t0 = clock64();
    while (clock64() - t0 < delay) {}
    __syncwarp();
    pos = nx;
  }
  if (acc == 0x1234567) sink[0] = acc;
}

int main() {
  const int grid = 148 * 11;
  const size_t region = 458752;   // 448 KB per warp (a C3 group stream)
  char* buf; unsigned long long* sink;
  cudaMalloc(&buf, region * grid + (region / 7 + 4096) * grid); cudaMalloc(&sink, 8);
  cudaMemset(buf, 0, region * grid);
  const int smem = 19584;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int steps = 3000;
  for (int wr : {0, 1, 2})
    for (int delay : {0, 1700})
      for (int pfk : {-1}) {
        const int spans = pfk < 0 ? 0 : 3;
        k<<<grid, 32, smem>>>(buf, region, 4096, 50, delay, spans, pfk < 0 ? 0 : pfk, wr, sink, 0);
        cudaEventRecord(e0);
        k<<<grid, 32, smem>>>(buf, region, 4096, steps, delay, spans, pfk < 0 ? 0 : pfk, wr, sink, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("write %d delay %4d cyc prefetch %s: %.0f ns/step, %.2f TB/s read (%s)\n", wr, delay, pfk < 0 ? "none" : (pfk ? "bulk" : "line"),
               ms * 1e6 / steps, (double)grid * steps * 4096 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
