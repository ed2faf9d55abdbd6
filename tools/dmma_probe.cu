// Does the FP64 tensor-core path (mma.sync m8n8k4 f64, "DMMA") run beside the FP64 FMA pipe on sm_100a?
// Round-2 question: the per-frequency contraction of an external product is a third of its flops; if
// DMMA has its own issue slot / datapath, the contraction of ciphertexts that share a GGSW could move there and
// leave the DFMA pipe to the butterflies.  NOT YET RUN (written after the round's GPU budget was spent).
// One CTA of 512 threads per SM.  Modes:
//   0  every warp: DFMA only (8 independent chains per thread, 64 DFMA per thread and iteration)
//   1  every warp: DMMA only (4 independent accumulator tiles per warp, 16 mma per iteration = 16 x 256 FMA per warp)
//   2  warps 0-7 DFMA only, warps 8-15 DMMA only
//   3  every warp interleaves 64 DFMA per thread with 16 DMMA
// Output: FMA per clock and SM of each mode, and whether mode 2 / 3 take max(a, b) or a + b.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* cyc, double* sink) {
  const int t = threadIdx.x, w = t >> 5;
  double x[8], acc[4][2];
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = 1.0 + 1e-9 * (t + j);
#pragma unroll
  for (int j = 0; j < 4; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
  const double a = 1.0 + 1e-12 * t, b = 1.0 - 1e-12 * t;
  const bool do_f = MODE == 0 || MODE == 3 || (MODE == 2 && w < 8);
  const bool do_m = MODE == 1 || MODE == 3 || (MODE == 2 && w >= 8);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (do_f) {
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
          for (int j = 0; j < 8; j++) x[j] = fma(x[j], 1.0000001, 1e-9);
      }
      if (do_m) {
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(acc[j], a, b);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (t == 0) cyc[blockIdx.x] = t1 - t0;
  double s = 0;
  for (int j = 0; j < 8; j++) s += x[j];
  for (int j = 0; j < 4; j++) s += acc[j][0] + acc[j][1];
  if (s == 1.2345) *sink = s;
}

template <int MODE>
static double run(const char* name, int iters, long long* c, double* s, double fma_per_iter_sm) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  probe<MODE><<<sms, 512>>>(iters, c, s);
  cudaDeviceSynchronize();
  probe<MODE><<<sms, 512>>>(iters, c, s);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256] = {0};
  cudaMemcpy(h, c, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < sms; i++) mx = h[i] > mx ? h[i] : mx;
  const double cyc = (double)mx / iters;
  printf("%-46s %8.1f cycles per iteration, %6.1f FMA/clk/SM (%s)\n", name, cyc, fma_per_iter_sm / cyc, cudaGetErrorString(e));
  return cyc;
}

int main() {
  long long* c; double* s;
  cudaMalloc(&c, 256 * 8); cudaMalloc(&s, 8);
  const int iters = 4000;
  // per iteration and SM: DFMA side 64 per thread x 512 threads (mode 0/3), x 256 (mode 2);
  //                       DMMA side 16 mma x 256 FMA x 16 warps (mode 1/3), x 8 warps (mode 2)
  const double F = 64.0 * 512, M = 16.0 * 256 * 16;
  const double a = run<0>("0: 16 warps DFMA", iters, c, s, F);
  const double b = run<1>("1: 16 warps DMMA m8n8k4", iters, c, s, M);
  const double m = run<2>("2: 8 warps DFMA + 8 warps DMMA", iters, c, s, F / 2 + M / 2);
  const double d = run<3>("3: 16 warps, DFMA and DMMA interleaved", iters, c, s, F + M);
  printf("if the two paths overlap: mode 2 ~ %.0f, mode 3 ~ %.0f cycles; if they share one datapath: %.0f / %.0f\n",
         a / 2 > b / 2 ? a / 2 : b / 2, a > b ? a : b, a / 2 + b / 2, a + b);
  printf("measured: mode 2 %.0f, mode 3 %.0f\n", m, d);
  return 0;
}
