// Micro-benchmark: which store path reaches the write-only HBM ceiling (torch fill: 7.5 TB/s) for the fused kernel's
// pattern -- every warp writes 8 chunks of S bytes per iteration, at addresses that advance with the CTA index?
//   mode 0: cp.async.bulk shared->global (UBLKCP), lane 0 issues the 8 copies             (what K2 does)
//   mode 1: same with an L2 evict_first cache hint
//   mode 2: STG.128 from registers, coalesced (512 B per warp instruction), same addresses
//   mode 3: LDS.128 + STG.128 from the staged tile, same addresses
//   mode 4: st.global.cs (streaming) from registers, same addresses
//   mode 5: fill-like grid-stride STG.128 over the whole buffer (sanity: torch fill)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_paths store_paths.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
template <int MODE>
__global__ void __launch_bounds__(128) k(char* out, size_t out_bytes, int S, int iters) {
  extern __shared__ __align__(128) char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* st = sm + size_t(warp) * 8 * S;
  for (int i = lane * 16; i < 8 * S; i += 512) *reinterpret_cast<int4*>(st + i) = make_int4(i, lane, warp, 7);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  uint64_t pol = 0;
  if (MODE == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const size_t gw = size_t(blockIdx.x) * 4 + warp, nw = size_t(gridDim.x) * 4;
  const int4 v = make_int4(lane, warp, 3, 4);
  if (MODE == 5) {
    const size_t n16 = out_bytes / 16, tid = size_t(blockIdx.x) * 128 + threadIdx.x, nt = size_t(gridDim.x) * 128;
    const size_t per = size_t(iters) * 8 * S / 512;   // 16-byte stores per thread: same bytes per warp as the other modes
    for (size_t i = tid, c = 0; c < per; i += nt, ++c) reinterpret_cast<int4*>(out)[i % n16] = v;
    return;
  }
  for (int it = 0; it < iters; ++it) {
    size_t base = ((size_t(it) * nw + gw) * 8) * size_t(S);
    base %= (out_bytes - size_t(8) * S);
    base &= ~size_t(127);
    if (MODE == 0 || MODE == 1) {
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (MODE == 0)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + base + size_t(i) * S),
                         "r"(smem_u32(st + i * S)), "r"(S) : "memory");
          else
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(
                             out + base + size_t(i) * S), "r"(smem_u32(st + i * S)), "r"(S), "l"(pol) : "memory");
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if ((it & 3) == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
      __syncwarp();
    } else {
      for (int i = lane * 16; i < 8 * S; i += 512) {
        if (MODE == 2) *reinterpret_cast<int4*>(out + base + i) = v;
        if (MODE == 3) *reinterpret_cast<int4*>(out + base + i) = *reinterpret_cast<const int4*>(st + i);
        if (MODE == 4) __stcs(reinterpret_cast<int4*>(out + base + i), v);
      }
    }
  }
  if (MODE == 0 || MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int MODE>
float run(char* out, size_t out_bytes, int S, int iters, int grid, size_t smem) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<grid, 128, smem>>>(out, out_bytes, S, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<grid, 128, smem>>>(out, out_bytes, S, iters);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}
int main() {
  size_t out_bytes = size_t(16) << 30;
  char* out;
  cudaMalloc(&out, out_bytes);
  cudaMemset(out, 0, out_bytes);
  const char* names[] = {"TMA s2g", "TMA s2g evict_first", "STG.128 regs", "LDS+STG.128", "st.cs regs", "fill-like"};
  for (int S : {1152, 4608})
    for (int cps : {3, 4, 6})   // CTAs of 4 warps per SM
      for (int mode = 0; mode < 6; ++mode) {
        const int grid = 148 * cps;
        const size_t smem = size_t(4) * 8 * S;
        if (smem * cps > 220 * 1024) continue;
        const int iters = int((size_t(12) << 30) / (size_t(grid) * 4 * 8 * S));
        float ms = 0;
        switch (mode) {
          case 0: ms = run<0>(out, out_bytes, S, iters, grid, smem); break;
          case 1: ms = run<1>(out, out_bytes, S, iters, grid, smem); break;
          case 2: ms = run<2>(out, out_bytes, S, iters, grid, smem); break;
          case 3: ms = run<3>(out, out_bytes, S, iters, grid, smem); break;
          case 4: ms = run<4>(out, out_bytes, S, iters, grid, smem); break;
          case 5: ms = run<5>(out, out_bytes, S, iters, grid, smem); break;
        }
        const double bytes = double(grid) * 4 * iters * 8 * S;
        printf("{\"S\": %d, \"warps_per_sm\": %d, \"mode\": \"%s\", \"ms\": %.3f, \"GBps\": %.1f, \"err\": %d}\n", S,
               cps * 4, names[mode], ms, bytes / ms / 1e6, int(cudaGetLastError()));
      }
  return 0;
}
