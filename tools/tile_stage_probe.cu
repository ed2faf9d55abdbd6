// Round-2 question: how fast can a CTA stream prepared-matrix tiles out of L2 when they are staged in
// shared memory by cp.async.bulk (one elected thread, mbarrier completion) instead of LDG.128 into registers?
// The contraction of k_ext3 waits one L2 round trip per tile because its registers hold one tile only.
// NOT YET RUN (written after the round's GPU budget was spent).
// Two CTAs of 256 threads per SM, working set 8 MiB (L2 resident), tiles of 16 KiB, each value consumed by one
// complex FMA per thread (the contraction's ratio is 6 per 16 bytes and stage; the loads are what is measured).
//   mode 0  LDG.128 into registers, one tile in flight (4 loads per thread), consume, next tile
//   mode 1  LDG.128, two tiles in flight (8 loads per thread)
//   mode 2  cp.async.bulk ring of S = 2 stages: thread 0 arms the mbarrier and issues the copy of tile i + S as soon
//           as tile i has been consumed by every thread (__syncthreads), consumers wait on the stage's mbarrier
//   mode 3  same, S = 4
// Output: bytes per clock and SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_stage_probe tile_stage_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kTile = 16384;                 // bytes
constexpr int kTileElems = kTile / 16;       // double2 per tile = 1024 = 4 per thread
constexpr size_t kSetBytes = (size_t)8 << 20;
constexpr int kTiles = (int)(kSetBytes / kTile);

__device__ __forceinline__ double2 ldg_nc(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
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
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(const double2* __restrict__ set, int tiles_per_cta, long long* cyc, double* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int S = MODE == 3 ? 4 : 2;
  double2* ring = reinterpret_cast<double2*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * kTile);
  const int t = threadIdx.x;
  // every CTA walks the whole set from its own offset (so that the tiles come from L2, not from one another's wake)
  const int first = (int)(((long)blockIdx.x * 37) % kTiles);
  double2 acc = make_double2(0.0, 0.0);
  const double2 a = make_double2(1.0 + 1e-9 * t, 1e-9);
  auto consume = [&](double2 g) {
    acc.x = fma(a.x, g.x, fma(-a.y, g.y, acc.x));
    acc.y = fma(a.x, g.y, fma(a.y, g.x, acc.y));
  };
  if (MODE >= 2) {
    if (t == 0) {
      for (int s = 0; s < S; s++) mbar_init(full + s, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  const long long t0 = clock64();
  if (MODE == 0 || MODE == 1) {
    constexpr int IN_FLIGHT = MODE == 0 ? 1 : 2;
    for (int i = 0; i < tiles_per_cta; i += IN_FLIGHT) {
      double2 g[IN_FLIGHT][4];
#pragma unroll
      for (int f = 0; f < IN_FLIGHT; f++) {
        const double2* p = set + (size_t)((first + i + f) % kTiles) * kTileElems + t;
#pragma unroll
        for (int j = 0; j < 4; j++) g[f][j] = ldg_nc(p + 256 * j);
      }
#pragma unroll
      for (int f = 0; f < IN_FLIGHT; f++)
#pragma unroll
        for (int j = 0; j < 4; j++) consume(g[f][j]);
    }
  } else {
    if (t == 0) {
      for (int s = 0; s < S && s < tiles_per_cta; s++) {
        mbar_expect_tx(full + s, kTile);
        bulk_g2s(ring + (size_t)s * kTileElems, set + (size_t)((first + s) % kTiles) * kTileElems, kTile, full + s);
      }
    }
    for (int i = 0; i < tiles_per_cta; i++) {
      const int s = i % S;
      mbar_wait(full + s, (uint32_t)((i / S) & 1));
      const double2* p = ring + (size_t)s * kTileElems + t;
#pragma unroll
      for (int j = 0; j < 4; j++) consume(p[256 * j]);
      __syncthreads();  // the stage is free
      if (t == 0 && i + S < tiles_per_cta) {
        mbar_expect_tx(full + s, kTile);
        bulk_g2s(ring + (size_t)s * kTileElems, set + (size_t)((first + i + S) % kTiles) * kTileElems, kTile, full + s);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (t == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc.x + acc.y == 1.2345) *sink = acc.x;
}

template <int MODE>
static void run(const char* name, const double2* set, int tiles_per_cta, long long* c, double* s, int sms) {
  const int grid = 2 * sms;
  const size_t smem = MODE >= 2 ? (size_t)(MODE == 3 ? 4 : 2) * kTile + 64 : 0;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<MODE><<<grid, 256, smem>>>(set, tiles_per_cta, c, s);
  cudaDeviceSynchronize();
  probe<MODE><<<grid, 256, smem>>>(set, tiles_per_cta, c, s);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[1024] = {0};
  cudaMemcpy(h, c, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
  printf("%-52s %9lld cycles, %6.1f B/clk/SM (%s)\n", name, mx, 2.0 * tiles_per_cta * kTile / (double)mx, cudaGetErrorString(e));
}

int main(int argc, char** argv) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (argc > 1) sms = atoi(argv[1]);  // stream on fewer SMs: per-SM port limit or aggregate L2 limit?
  printf("streaming on %d SMs\n", sms);
  double2* set; long long* c; double* s;
  cudaMalloc(&set, kSetBytes); cudaMalloc(&c, 1024 * 8); cudaMalloc(&s, 8);
  cudaMemset(set, 0, kSetBytes);
  const int tiles = 2048;  // 32 MiB per CTA
  run<0>("0: LDG.128, one 16 KiB tile in flight", set, tiles, c, s, sms);
  run<1>("1: LDG.128, two tiles in flight", set, tiles, c, s, sms);
  run<2>("2: cp.async.bulk ring, 2 stages", set, tiles, c, s, sms);
  run<3>("3: cp.async.bulk ring, 4 stages", set, tiles, c, s, sms);
  return 0;
}
