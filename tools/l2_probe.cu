// L2 -> SM streaming probe: bytes/clk/SM when every SM streams the same L2-resident buffer
// (what the contraction does with the prepared matrices).  nvcc -arch=sm_100a -O3 l2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int UNROLL>
__global__ void stream(const double2* __restrict__ buf, size_t n, int iters, long long* cyc, double* sink) {
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    for (size_t base = 0; base + (size_t)blockDim.x * UNROLL <= n; base += (size_t)blockDim.x * UNROLL) {
      double2 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) v[u] = __ldg(buf + base + (size_t)u * blockDim.x + threadIdx.x);
#pragma unroll
      for (int u = 0; u < UNROLL; u++) acc += v[u].x + v[u].y;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 1.2345) *sink = acc;
}
int main() {
  const size_t n = (size_t)24 << 20 >> 4;  // 24 MiB of double2
  double2* d; long long* c; double* s;
  cudaMalloc(&d, n * 16); cudaMemset(d, 0, n * 16); cudaMalloc(&c, 148 * 8); cudaMalloc(&s, 8);
  for (int threads : {256, 512, 1024}) {
    for (int un : {8, 16}) {
      const int iters = 4;
      if (un == 8) stream<8><<<148, threads>>>(d, n, iters, c, s); else stream<16><<<148, threads>>>(d, n, iters, c, s);
      cudaDeviceSynchronize();
      if (un == 8) stream<8><<<148, threads>>>(d, n, iters, c, s); else stream<16><<<148, threads>>>(d, n, iters, c, s);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (auto x : h) mx = x > mx ? x : mx;
      printf("threads %4d unroll %2d (%3d KiB in flight/SM): %.1f B/clk/SM  (%s)\n", threads, un,
             threads * un * 16 / 1024, (double)n * 16 * iters / mx, cudaGetErrorString(e));
    }
  }
}
