// TMEM probe: can tensor memory be used as lane-private scratch (tcgen05.st / tcgen05.ld without
// any MMA), and at what throughput?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tmem_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
         "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
         "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
         "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// every warp owns `cols_per_warp` columns of its lane quadrant; writes a pattern, reads it back
__global__ void probe(int iters, int cols_per_warp, long long* cyc, unsigned* bad) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(&base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(sa) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = base_s;
  const int q = warp & 3, slot = warp >> 2;  // lane quadrant, column slot
  const uint32_t my = base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * cols_per_warp);
  uint32_t r[32];
  unsigned errs = 0;
  for (int c = 0; c < cols_per_warp; c += 32) {
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = (threadIdx.x << 16) ^ ((c + i) * 2654435761u);
    tmem_st32(my + c, r);
  }
  tmem_wait_st();
  for (int c = 0; c < cols_per_warp; c += 32) {
    tmem_ld32(my + c, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; i++) errs += r[i] != ((threadIdx.x << 16) ^ ((c + i) * 2654435761u));
  }
  __syncthreads();
  // throughput: loads
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
    for (int c = 0; c < cols_per_warp; c += 32) {
      tmem_ld32(my + c, r);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i++) acc ^= r[i];
    }
  }
  __syncthreads();
  long long t1 = clock64();
  for (int it = 0; it < iters; it++) {
    for (int c = 0; c < cols_per_warp; c += 32) {
#pragma unroll
      for (int i = 0; i < 32; i++) r[i] = acc + i + it;
      tmem_st32(my + c, r);
    }
    tmem_wait_st();
  }
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = t1 - t0; cyc[blockIdx.x * 2 + 1] = t2 - t1; }
  if (errs || acc == 0x12345) atomicAdd(bad, errs);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}

int main() {
  long long* d_c; unsigned* d_bad;
  cudaMalloc(&d_c, sizeof(long long) * 2 * 148);
  cudaMalloc(&d_bad, 4);
  for (int warps : {4, 8, 16}) {
    cudaMemset(d_bad, 0, 4);
    const int cols = 512 / ((warps + 3) / 4);  // columns per warp
    const int iters = 200;
    probe<<<148, warps * 32>>>(iters, cols, d_c, d_bad);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; unsigned bad;
    cudaMemcpy(h, d_c, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * warps * cols * 32 * 4;
    printf("warps %2d cols/warp %3d: %s mismatches %u | ld %.1f B/clk/SM  st %.1f B/clk/SM\n", warps, cols,
           cudaGetErrorString(e), bad, bytes / h[0], bytes / h[1]);
  }
  return 0;
}
