// Which (lane, column) does each (thread, register) of the 16-lane tcgen05.ld shapes read?  Written to decide
// whether the warp-local exchange of the 16-point transform (kernels_ks7.cuh, exchange 2) can go through tensor
// memory instead of shared memory: store with 32x32b (thread i -> lane i, register k -> column k), load with
// 16x256b / 16x128b / 16x64b at lane offsets 0 and 16, print the mapping.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_shape_probe tmem_shape_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)(w * 32) << 16);
  // value = lane * 256 + column  (lane within the warp's quadrant)
  uint32_t v[16];
#pragma unroll
  for (int k = 0; k < 16; k++) v[k] = (uint32_t)(lane * 256 + k);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                 "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t r[8];
  // 16x256b.x1: 4 registers, lanes base .. base+15, columns 0..7
  for (int half = 0; half < 2; half++) {
    const uint32_t a = base + ((uint32_t)(16 * half) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (w == 0) for (int q = 0; q < 4; q++) out[(0 * 2 + half) * 32 * 8 + lane * 8 + q] = r[q];
    // 16x256b.x2: 8 registers, columns 0..15
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (w == 0) for (int q = 0; q < 8; q++) out[(1 * 2 + half) * 32 * 8 + lane * 8 + q] = r[q];
    // 16x128b.x1: 2 registers
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (w == 0) for (int q = 0; q < 2; q++) out[(2 * 2 + half) * 32 * 8 + lane * 8 + q] = r[q];
    // 16x64b.x1: 1 register
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (w == 0) out[(3 * 2 + half) * 32 * 8 + lane * 8] = r[0];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(slot) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 8 * 32 * 8 * 4); cudaMemset(d, 0xff, 4 * 8 * 32 * 8 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  static uint32_t h[8 * 32 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[4] = {"16x256b.x1", "16x256b.x2", "16x128b.x1", "16x64b.x1"};
  const int nreg[4] = {4, 8, 2, 1};
  for (int s = 0; s < 4; s++)
    for (int half = 0; half < 2; half++) {
      printf("%s lane offset %d: thread -> (lane,col) per register\n", names[s], 16 * half);
      for (int t = 0; t < 32; t++) {
        printf("  T%02d:", t);
        for (int q = 0; q < nreg[s]; q++) { uint32_t v = h[(s * 2 + half) * 256 + t * 8 + q]; printf(" (%u,%u)", v >> 8, v & 255); }
        if (t % 4 == 3) printf("\n");
      }
    }
  return 0;
}
