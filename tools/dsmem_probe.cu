// Round-2 question (frequency split of the external product over a cluster of two CTAs): what does the
// one exchange cost?  After the half-size inverse transforms each CTA needs the other's 1 024 complex values per output
// polynomial (16 KiB per output and direction, 8 outputs per external product).
// NOT YET RUN (written after the round's GPU budget was spent).
// Cluster of 2 CTAs x 256 threads, one or two clusters' CTAs per SM (the driver places them).  Per round every thread
//   mode 0  writes 4 double2 into its OWN shared memory, __syncthreads, reads 4 back          (local baseline)
//   mode 1  writes 4 double2 into the PEER's shared memory (st.shared::cluster.v2.f64), cluster barrier, reads 4 local
//   mode 2  writes 4 local, cluster barrier, reads 4 from the PEER (ld.shared::cluster.v2.f64)
// Output: cycles per 16 KiB round and bytes per clock and CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_probe dsmem_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double2 v) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double2 ld_cluster(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 2) probe(int rounds, long long* cyc, double* sink) {
  __shared__ __align__(16) double2 buf[1024];
  const int t = threadIdx.x;
  const uint32_t peer = mapa(smem_u32(buf), ctarank() ^ 1u);
  double2 x[4];
#pragma unroll
  for (int j = 0; j < 4; j++) { x[j] = make_double2(1.0 + t, 0.5 * j); buf[t + 256 * j] = x[j]; }
  cluster_sync();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; r++) {
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; j++) st_cluster(peer + 16u * (uint32_t)(t + 256 * j), x[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) buf[t + 256 * j] = x[j];
    }
    if (MODE == 0) __syncthreads(); else cluster_sync();
    if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 4; j++) x[j] = ld_cluster(peer + 16u * (uint32_t)(((t + 32) & 255) + 256 * j));
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) x[j] = buf[((t + 32) & 255) + 256 * j];
    }
#pragma unroll
    for (int j = 0; j < 4; j++) x[j].x = fma(x[j].x, 1.0000001, x[j].y);
    if (MODE == 0) __syncthreads(); else cluster_sync();  // reads done before the next round's writes
  }
  const long long t1 = clock64();
  cluster_sync();
  if (t == 0) cyc[blockIdx.x] = t1 - t0;
  double s = 0;
  for (int j = 0; j < 4; j++) s += x[j].x + x[j].y;
  if (s == 1.2345) *sink = s;
}

template <int MODE>
static void run(const char* name, int rounds, long long* c, double* s, int grid) {
  probe<MODE><<<grid, 256>>>(rounds, c, s);
  cudaDeviceSynchronize();
  probe<MODE><<<grid, 256>>>(rounds, c, s);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[1024] = {0};
  cudaMemcpy(h, c, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
  printf("%-58s grid %4d: %7.1f cycles per 16 KiB round, %5.1f B/clk/CTA each way (%s)\n", name, grid, (double)mx / rounds,
         16384.0 * rounds / (double)mx, cudaGetErrorString(e));
}

int main() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* c; double* s;
  cudaMalloc(&c, 1024 * 8); cudaMalloc(&s, 8);
  const int rounds = 20000;
  for (int grid : {sms & ~1, 2 * sms}) {  // one CTA per SM, two CTAs per SM
    run<0>("0: local store, __syncthreads, local load", rounds, c, s, grid);
    run<1>("1: remote store (st.shared::cluster), cluster barrier", rounds, c, s, grid);
    run<2>("2: local store, cluster barrier, remote load", rounds, c, s, grid);
  }
  return 0;
}
