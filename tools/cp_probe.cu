// Round-2 question: can prepared-matrix tiles reach the threads WITHOUT the LSU / l1tex data pipe (the unit ncu shows
// closest to saturation in k_ext3 / k_ks7)?  Path tried here: cp.async.bulk (L2 -> shared memory, async proxy), then
// tcgen05.cp.128x128b (shared memory -> tensor memory, issued by one thread, no registers), then tcgen05.ld 32x32b
// (tensor memory -> the lane that owns the frequency).  Measured: (1) layout correctness of the 128x128b copy with a
// SWIZZLE_NONE descriptor on plain contiguous data (row t = 16 bytes at base + 16 t), (2) the copy rate alone,
// (3) what a saturating LDS/STS loop on the other warps loses while the copy pipeline runs, against the same bytes
// streamed with LDG.128 by the warps themselves.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cp_probe cp_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(const void* smem) {
  // SWIZZLE_NONE, K-major: core matrix = 8 rows x 16 bytes stored contiguously (128 B); SBO = 128 B between core
  // matrices along the rows; LBO unused (one 16-byte column); version 1 (sm_100)
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;   // LBO (unused)
  d |= (uint64_t)(128 >> 4) << 32;   // SBO
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tc_cp_128x128b(uint32_t taddr, uint64_t desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}

constexpr int kChunk = 8192;          // bytes: 4 frequencies x 128 lanes x 16 B
constexpr int kStages = 4;
constexpr size_t kSet = (size_t)8 << 20;

// MODE 0: LSU loop only; 1: LSU loop (warps 1..15) + copy pipeline (warp 0); 2: copy pipeline only (+ check);
// 3: LSU loop + the same bytes streamed by LDG.128 in the loop
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(const double2* __restrict__ set, int iters, int chunks, long long* cyc,
                                                double* sink, int* bad) {
  extern __shared__ __align__(1024) unsigned char smem[];
  double2* ring = reinterpret_cast<double2*>(smem);                       // kStages x 8 KiB
  double2* priv = reinterpret_cast<double2*>(smem + kStages * kChunk);    // 512 x 9 double2 (padded)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kChunk + 512 * 9 * 16);
  uint64_t* done = full + kStages;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + kStages);
  const int tid = threadIdx.x, w = tid >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(full + s, 1); mbar_init(done + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *slot;
  const size_t first = ((size_t)blockIdx.x * 37 * kChunk) % kSet;
  double2 acc = make_double2(1.0 + tid, 0.5);
  double2* mine = priv + 9 * tid;
  for (int j = 0; j < 8; j++) mine[j] = make_double2(tid, j);
  __syncthreads();
  const long long t0 = clock64();
  long long lsu_cyc = 0, cp_cyc = 0;
  if (MODE == 1 ? (w == 0) : (MODE == 2)) {
    if (tid == 0) {
      // producer: keep kStages TMA copies in flight; when one lands, copy it to tensor memory and commit
      for (int c = 0; c < kStages && c < chunks; c++) {
        mbar_expect_tx(full + c, kChunk);
        bulk_g2s(ring + (size_t)c * 512, (const char*)set + (first + (size_t)c * kChunk) % kSet, kChunk, full + c);
      }
      // the refill of a stage lags its tensor-memory copy by kStages / 2 chunks, so neither latency is exposed
      constexpr int LAG = kStages / 2;
      auto refill = [&](int c) {  // chunk c has been copied to tensor memory: its stage takes chunk c + kStages
        const int s = c % kStages;
        mbar_wait(done + s, (uint32_t)((c / kStages) & 1));
        if (c + kStages < chunks) {
          mbar_expect_tx(full + s, kChunk);
          bulk_g2s(ring + (size_t)s * 512, (const char*)set + (first + (size_t)(c + kStages) * kChunk) % kSet, kChunk, full + s);
        }
      };
      for (int c = 0; c < chunks; c++) {
        const int s = c % kStages;
        mbar_wait(full + s, (uint32_t)((c / kStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int r = 0; r < 4; r++) tc_cp_128x128b(tbase + 16 * s + 4 * r, make_desc(ring + (size_t)s * 512 + 128 * r));
        tc_commit(done + s);
        if (c >= LAG) refill(c - LAG);
      }
      for (int c = chunks - LAG; c < chunks; c++) refill(c);
      cp_cyc = clock64() - t0;
    }
  }
  if (MODE == 0 || MODE == 3 || (MODE == 1 && w != 0)) {
        for (int it = 0; it < iters; it++) {
      double2 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[j].x), "=d"(v[j].y) : "r"(smem_u32(mine + j)) : "memory");
      if (MODE == 3) {
        const double2 g = __ldg(set + (((size_t)it * 512 + (size_t)blockIdx.x * 7919) % (kSet / 16 - 512)) + tid);
        acc.x += g.x; acc.y += g.y;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        v[j].x = fma(v[j].x, 1.0000001, acc.y);
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(smem_u32(mine + j)), "d"(v[j].x), "d"(v[j].y) : "memory");
      }
      acc.x += v[3].x;
    }
    lsu_cyc = clock64() - t0;
  }
  __syncthreads();
  if (tid == 32) cyc[2 * blockIdx.x] = lsu_cyc;
  if (tid == 0) cyc[2 * blockIdx.x + 1] = MODE == 0 || MODE == 3 ? lsu_cyc : cp_cyc;
  if (MODE == 2 && tid < 128) {
    // check the last chunk: lane t, columns 16 s + 4 r .. + 3 must hold set[chunk base + 128 r + t]
    const int c = chunks - 1, s = c % kStages;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int nbad = 0;
    for (int r = 0; r < 4; r++) {
      uint32_t q[4];
      const uint32_t ta = tbase + ((uint32_t)((w & 3) * 32) << 16) + 16 * s + 4 * r;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const double x = __hiloint2double((int)q[1], (int)q[0]), y = __hiloint2double((int)q[3], (int)q[2]);
      const double2 want = set[((first + (size_t)c * kChunk) % kSet) / 16 + 128 * r + tid];
      if (x != want.x || y != want.y) nbad++;
    }
    if (nbad) atomicAdd(bad, nbad);
  }
  if (acc.x + acc.y == 1.2345) *sink = acc.x;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}

template <int MODE>
static void run(const char* name, const double2* set, int iters, int chunks, long long* c, double* s, int* bad, int sms) {
  const size_t smem = kStages * kChunk + 512 * 9 * 16 + 256;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemset(bad, 0, 4);
  probe<MODE><<<sms, 512, smem>>>(set, iters, chunks, c, s, bad);
  cudaDeviceSynchronize();
  cudaMemset(bad, 0, 4);
  probe<MODE><<<sms, 512, smem>>>(set, iters, chunks, c, s, bad);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[1024] = {0};
  cudaMemcpy(h, c, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost);
  int hb = 0;
  cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  long long ml = 0, mc = 0;
  for (int i = 0; i < sms; i++) { if (h[2 * i] > ml) ml = h[2 * i]; if (h[2 * i + 1] > mc) mc = h[2 * i + 1]; }
  // LSU loop: 16 B x 16 accesses per thread and iteration
  printf("%-60s lsu loop %8.1f cyc/iter (%5.1f B/clk/SM)  copy %6.1f B/clk/SM  mismatches %d (%s)\n", name,
         ml ? (double)ml / iters : 0.0, ml ? (MODE == 1 ? 480 : 512) * 256.0 * iters / ml : 0.0,
         (MODE == 1 || MODE == 2) && mc ? (double)chunks * kChunk / mc : 0.0, hb, cudaGetErrorString(e));
}

__global__ void fill(double2* p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_double2((double)i, -(double)i - 0.5);
}

int main() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double2* set; long long* c; double* s; int* bad;
  cudaMalloc(&set, kSet + 65536); cudaMalloc(&c, 1024 * 8); cudaMalloc(&s, 8); cudaMalloc(&bad, 4);
  fill<<<256, 256>>>(set, (kSet + 65536) / 16);
  const int iters = 1000, chunks = 8000;
  run<2>("2: TMA -> smem -> tcgen05.cp -> tmem only", set, iters, chunks, c, s, bad, sms);
  run<0>("0: LDS/STS loop only (16 warps)", set, iters, chunks, c, s, bad, sms);
  run<1>("1: LDS/STS loop (15 warps) + copy pipeline (1 thread)", set, iters, chunks, c, s, bad, sms);
  run<3>("3: LDS/STS loop + one LDG.128 per thread and iteration", set, iters, chunks, c, s, bad, sms);
  return 0;
}
