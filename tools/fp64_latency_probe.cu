// FP64 latency / ILP probe: how many independent DFMA chains per warp does the B200 FP64 pipe need?
// One CTA per SM; W warps; every thread runs C independent chains of dependent DFMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency_probe fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void probe(int iters, long long* cyc, double* sink) {
  double x[C];
#pragma unroll
  for (int j = 0; j < C; j++) x[j] = 1.0 + threadIdx.x + j;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int j = 0; j < C; j++) x[j] = fma(x[j], 1.0000001, 0.5);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  double s = 0;
  for (int j = 0; j < C; j++) s += x[j];
  if (s == 1.2345) *sink = s;
}
template <int C>
static void run(int warps, long long* c, double* s) {
  const int iters = 2000;
  probe<C><<<148, 32 * warps>>>(iters, c, s);
  cudaDeviceSynchronize();
  probe<C><<<148, 32 * warps>>>(iters, c, s);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  const double per_step = (double)mx / (iters * 8.0);        // cycles per dependent step (C DFMAs per warp)
  const double rate = (double)warps * C * 32 / per_step;     // FMA lanes per clock per SM
  printf("warps %2d  chains/thread %2d: %6.1f cycles per dependent step, %5.1f FMA/clk/SM (peak 64)\n", warps, C, per_step, rate);
}
int main() {
  long long* c; double* s;
  cudaMalloc(&c, 148 * 8); cudaMalloc(&s, 8);
  for (int w : {4, 8, 16, 32}) {
    run<1>(w, c, s); run<2>(w, c, s); run<4>(w, c, s); run<8>(w, c, s); run<16>(w, c, s);
  }
  return 0;
}
