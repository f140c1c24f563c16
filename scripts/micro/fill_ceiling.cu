// Micro-benchmark: is the 6.18 TB/s of store_paths.cu a DRAM write ceiling or a property of its launch shape?
// torch's fill_ of 8.6 GB was measured at 7.5 TB/s on the same GPUs.  Variants of a plain fill:
//   A  one-shot grid (one 4 KB chunk per 128-thread CTA, 2 x 16-B stores per thread), constant value    (torch-like)
//   B  the same with a value that depends on the address                                               (rules out any same-value effect)
//   C  one-shot grid, 9216-B chunk per one-warp CTA (K2's shape: one warp = one node pair's COO+CSR share)
//   D  persistent grid, W warps per SM, each warp 9216-B chunks in a grid-stride sweep                 (store_paths.cu's shape)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fill_ceiling fill_ceiling.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128) kA(int4* out, size_t n16, int vary) {
  const size_t i = (size_t(blockIdx.x) * 128 + threadIdx.x) * 2;
  if (i + 1 < n16) {
    int4 v = make_int4(1, 2, 3, 4);
    if (vary) v = make_int4(int(i), int(i >> 3), int(i * 7), int(i ^ 0x5555));
    out[i] = v;
    v.x += vary;
    out[i + 1] = v;
  }
}
__global__ void __launch_bounds__(32) kC(int4* out, size_t nchunks) {
  const size_t c = blockIdx.x;
  if (c >= nchunks) return;
  int4* p = out + c * 576;   // 9216 B = 576 x 16
  const int4 v = make_int4(threadIdx.x, int(c), 3, 4);
#pragma unroll
  for (int i = 0; i < 18; ++i) p[i * 32 + threadIdx.x] = v;
}
__global__ void __launch_bounds__(128) kD(int4* out, size_t nchunks) {
  const size_t gw = size_t(blockIdx.x) * 4 + (threadIdx.x >> 5), nw = size_t(gridDim.x) * 4;
  const int lane = threadIdx.x & 31;
  const int4 v = make_int4(lane, int(gw), 3, 4);
  for (size_t c = gw; c < nchunks; c += nw) {
    int4* p = out + c * 576;
#pragma unroll
    for (int i = 0; i < 18; ++i) p[i * 32 + lane] = v;
  }
}
template <class F>
float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}
int main() {
  const size_t bytes = size_t(12) << 30, n16 = bytes / 16, nchunks = bytes / 9216;
  int4* out;
  cudaMalloc(&out, bytes);
  cudaMemset(out, 0, bytes);
  auto report = [&](const char* name, float ms, double b) {
    printf("{\"mode\": \"%s\", \"ms\": %.3f, \"GBps\": %.1f, \"err\": %d}\n", name, ms, b / ms / 1e6, int(cudaGetLastError()));
  };
  report("A one-shot 4KB/CTA constant", timeit([&] { kA<<<unsigned(n16 / 256), 128>>>(out, n16, 0); }), double(bytes));
  report("B one-shot 4KB/CTA varying", timeit([&] { kA<<<unsigned(n16 / 256), 128>>>(out, n16, 1); }), double(bytes));
  report("C one-shot 9216B per one-warp CTA", timeit([&] { kC<<<unsigned(nchunks), 32>>>(out, nchunks); }), double(nchunks) * 9216);
  for (int cps : {3, 6, 8, 12, 16})
    for (int rep = 0; rep < 1; ++rep) {
      char name[96];
      snprintf(name, sizeof name, "D persistent %d warps/SM 9216B chunks", cps * 4);
      report(name, timeit([&] { kD<<<148 * cps, 128>>>(out, nchunks); }), double(nchunks) * 9216);
    }
  cudaMemsetAsync(out, 0, bytes);
  report("cudaMemset", timeit([&] { cudaMemsetAsync(out, 1, bytes); }), double(bytes));
  return 0;
}
