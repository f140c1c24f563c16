// Micro-benchmark: cost of small cp.async.bulk shared->global copies (UBLKCP) per SM, as used by the fused
// kernel for the COO slabs.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_s2g tma_s2g.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
template <int MODE>
__global__ void __launch_bounds__(128, 3) k(char* out, size_t out_bytes, int S, int iters, int ncopy) {
  extern __shared__ __align__(128) char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* st = sm + size_t(warp) * 8 * S;
  for (int i = lane * 16; i < 8 * S; i += 512) *reinterpret_cast<int4*>(st + i) = make_int4(i, lane, warp, 7);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t gw = size_t(blockIdx.x) * 4 + warp, nw = size_t(gridDim.x) * 4;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 3) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    size_t base = ((size_t(it) * nw + gw) * 8) * size_t(S);
    base %= (out_bytes - size_t(8) * S);
    base &= ~size_t(127);
    if (MODE == 0 || MODE == 1) {
      if ((lane & 3) == 0 && (lane >> 2) < ncopy) {
        // make the address look divergent to the compiler
        const size_t off = base + size_t(lane >> 2) * S;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + off),
                     "r"(smem_u32(st + (lane >> 2) * S)), "r"(S) : "memory");
      }
    } else {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < ncopy)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + base + size_t(i) * S),
                         "r"(smem_u32(st + i * S)), "r"(S) : "memory");
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    __syncwarp();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  size_t out_bytes = size_t(8) << 30;
  char* out;
  cudaMalloc(&out, out_bytes);
  cudaMemset(out, 0, out_bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int sizes[] = {288, 576, 960, 1152, 2304, 4608};
  for (int mode = 0; mode < 4; ++mode)
    for (int S : sizes)
      for (int ncopy : {8, 2}) {
        const int grid = 148 * 3, iters = 400;
        const size_t smem = size_t(4) * 8 * S;
        if (smem * 3 > 220 * 1024) continue;
        auto launch = [&]() {
          if (mode == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); k<0><<<grid, 128, smem>>>(out, out_bytes, S, iters, ncopy); }
          if (mode == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); k<1><<<grid, 128, smem>>>(out, out_bytes, S, iters, ncopy); }
          if (mode == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); k<2><<<grid, 128, smem>>>(out, out_bytes, S, iters, ncopy); }
          if (mode == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)); k<3><<<grid, 128, smem>>>(out, out_bytes, S, iters, ncopy); }
        };
        launch();
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double copies = double(grid) * 4 * iters * ncopy;
        const double cyc_per_copy_sm = ms * 1e-3 * 1.965e9 / (copies / 148.);
        printf("mode %d (%s) S %5d ncopy %d: %.3f ms  %.1f GB/s  %.1f cycles/copy/SM  err %d\n", mode,
               mode == 0 ? "divergent issue, wait.read each iter" : mode == 1 ? "divergent issue, no wait"
               : mode == 2 ? "lane0 uniform issue, no wait" : "lane0 uniform issue, wait.read each iter",
               S, ncopy, ms, copies * S / ms / 1e6, cyc_per_copy_sm, int(cudaGetLastError()));
      }
  return 0;
}
