// kernels_ks2.cuh -- key-switch kernels (trace / one-sided packer level chains, two-sided packer
// combine) built for TWO resident CTAs per SM.
//
// The single-CTA kernel (kernels.cuh: k_vmp) runs its phases back to back: matrix streaming
// (L2 bound), transforms (shared-memory bound) and the integer epilogue (ALU bound) never
// overlap, and 224 KiB of shared memory + 250 registers leave room for only 8 warps per SM.
// Here the per-operation footprint is cut so that two CTAs fit and each other's phases overlap:
//   * the three input spectra are thread-private (the thread that finishes a forward transform
//     is the only one that reads its 8 frequencies back), so they live in TENSOR MEMORY
//     (tcgen05.st / tcgen05.ld as lane-private scratch, 96 columns per thread; no MMA involved),
//   * the late-stage twiddles (32 registers in k_vmp) are parked in tensor memory as well,
//   * the ciphertext buffer is packed 3 limbs -> one 64-bit word (21 bits per limb): 64 KiB,
//   * one 32 KiB exchange buffer,
// i.e. 96 KiB shared memory, 256 TMEM columns and <= 128 registers per CTA.
// Arithmetic, dataflow and results are identical to k_vmp<3,1,4,3,MODE_TRACE|MODE_COMBINE2>.
#pragma once
#include "kernels.cuh"

namespace fheram {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- tensor memory as lane-private scratch ------------------------------------------------
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 16 consecutive 32-bit columns of this thread's TMEM lane <-> 4 double2
__device__ __forceinline__ void tm_st4(uint32_t taddr, const double2 (&v)[4]) {
  uint32_t r[16];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    r[4 * j] = (uint32_t)__double2loint(v[j].x); r[4 * j + 1] = (uint32_t)__double2hiint(v[j].x);
    r[4 * j + 2] = (uint32_t)__double2loint(v[j].y); r[4 * j + 3] = (uint32_t)__double2hiint(v[j].y);
  }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, double2 (&v)[4]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tm_wait_ld();
#pragma unroll
  for (int j = 0; j < 4; j++) {
    v[j].x = __hiloint2double((int)r[4 * j + 1], (int)r[4 * j]);
    v[j].y = __hiloint2double((int)r[4 * j + 3], (int)r[4 * j + 2]);
  }
}

// ---- packed ciphertext words: limb l in bits [21 l, 21 l + 21), two's complement -------------
__device__ __forceinline__ long long pack3(int l0, int l1, int l2) {
  return (long long)(l0 & 0x1fffff) | ((long long)(l1 & 0x1fffff) << 21) | ((long long)(l2 & 0x1fffff) << 42);
}
__device__ __forceinline__ int unpack3(long long wd, int l) {
  long long r;
  asm("bfe.s64 %0, %1, %2, 21;" : "=l"(r) : "l"(wd), "r"(21 * l));
  return (int)r;
}
__device__ __forceinline__ long long repack3(long long wd, int l, int v) {
  long long r;
  asm("bfi.b64 %0, %1, %2, %3, 21;" : "=l"(r) : "l"((long long)v), "l"(wd), "r"(21 * l));
  return r;
}

// passes 2-4 of the forward transform inside the warp's 256-element block of `work`; returns
// the thread's 8 final frequencies in x (position 256w + 32j + lane) instead of storing them
__device__ __forceinline__ void fwd_warp_passes2(double2* work, int w, int lane, uint32_t ttw, double2 (&x)[8]) {
  double2* base = work + 256 * w;
#pragma unroll
  for (int m = 0; m < 8; m++) x[m] = base[S1(lane + 32 * m)];
  radix8_fwd<true>(x, c_tw_lo[8 + w], c_tw_lo[16 + 2 * w], c_tw_lo[32 + 4 * w], c_tw_lo[32 + 4 * w + 2]);
#pragma unroll
  for (int m = 0; m < 8; m++) base[S1(lane + 32 * m)] = x[m];
  __syncwarp();
  const int qr = 32 * (lane >> 2) + (lane & 3);
#pragma unroll
  for (int m = 0; m < 8; m++) x[m] = base[S1(qr + 4 * m)];
  {
    double2 t[4];
    tm_ld4(ttw, t);  // a3, b3, c3, d3
    radix8_fwd<true>(x, t[0], t[1], t[2], t[3]);
  }
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 8; m++) base[S2(qr + 4 * m)] = x[m];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) x[j] = base[S2(8 * lane + j)];
  {
    double2 t[4];
    tm_ld4(ttw + 16, t);  // b4a, b4b, c4a, c4b
    bf(x[0], x[2], t[0]); bf(x[1], x[3], t[0]); bf(x[4], x[6], t[1]); bf(x[5], x[7], t[1]);
    bf(x[0], x[1], t[2]); bf(x[2], x[3], mul_i(t[2]));
    bf(x[4], x[5], t[3]); bf(x[6], x[7], mul_i(t[3]));
  }
}

template <typename F>
__device__ __forceinline__ void inv_transform2(double2 (&x)[8], double2* work, int T, int w, int lane,
                                               uint32_t ttw, F&& pre_sync2) {
  double2* wb = work + 256 * w;
  {
    double2 t[4];
    tm_ld4(ttw + 16, t);
    ibf(x[0], x[1], t[2]); ibf(x[2], x[3], mul_i(t[2]));
    ibf(x[4], x[5], t[3]); ibf(x[6], x[7], mul_i(t[3]));
    ibf(x[0], x[2], t[0]); ibf(x[1], x[3], t[0]); ibf(x[4], x[6], t[1]); ibf(x[5], x[7], t[1]);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) wb[S2(8 * lane + j)] = x[j];
  __syncwarp();
  const int qr = 32 * (lane >> 2) + (lane & 3);
#pragma unroll
  for (int m = 0; m < 8; m++) x[m] = wb[S2(qr + 4 * m)];
  {
    double2 t[4];
    tm_ld4(ttw, t);
    radix8_inv<true>(x, t[0], t[1], t[2], t[3]);
  }
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 8; m++) wb[S1(qr + 4 * m)] = x[m];
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 8; m++) x[m] = wb[S1(lane + 32 * m)];
  radix8_inv<true>(x, c_tw_lo[8 + w], c_tw_lo[16 + 2 * w], c_tw_lo[32 + 4 * w], c_tw_lo[32 + 4 * w + 2]);
#pragma unroll
  for (int m = 0; m < 8; m++) wb[S1(lane + 32 * m)] = x[m];
  __syncthreads();
  pre_sync2();
#pragma unroll
  for (int m = 0; m < 8; m++) x[m] = work[S1(T + 256 * m)];
  __syncthreads();
  radix8_inv<true>(x, c_tw_lo[1], c_tw_lo[2], c_tw_lo[4], c_tw_lo[6]);
}

// ======================================================================================
// k_ext2: external-product chains with TWO CTAs per SM.  Of the six input spectra four live in
// tensor memory (128 columns per thread) and two in shared memory; pass-4 twiddles sit in a
// 16 KiB shared table, pass-3 twiddles in __constant__ (8 distinct addresses per warp).
// 32 (exchange) + 64 (two spectra) + 16 (twiddles) = 112 KiB shared memory, 256 TMEM columns,
// <= 128 registers per CTA.  Same arithmetic and results as k_vmp<3,2,4,3,MODE_EXT>.
// ======================================================================================
__constant__ double2 c_tw3[256];  // zeta(6,B) | zeta(7,2B) | zeta(8,2k)   (Twiddles::tw6/tw7c/tw8c)

struct Tw3 { double2 a, b, c, d; };
__device__ __forceinline__ Tw3 load_tw3(int w, int lane) {
  const int B = 8 * w + (lane >> 2);
  Tw3 t;
  t.a = c_tw3[B]; t.b = c_tw3[64 + B]; t.c = c_tw3[128 + 2 * B]; t.d = c_tw3[128 + 2 * B + 1];
  return t;
}

constexpr size_t kExt2Smem = (size_t)kM * sizeof(double2) * 3 + (size_t)4 * kThreads * sizeof(double2) + 16;

__global__ void __launch_bounds__(kThreads, 2) k_ext2(const VmpArgs A) {
  constexpr int NR = 6, LOUT = 4, LRES = 3, NOUT = 2 * LOUT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* work = reinterpret_cast<double2*>(smem_raw);
  double2* rows_s = work + kM;                 // spectra of rows 4 and 5
  double2* tw4s = rows_s + 2 * kM;             // [4][256] pass-4 twiddles, thread-private columns
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tw4s + 4 * kThreads);

  const int T = threadIdx.x, w = T >> 5, lane = T & 31;
  auto CT = [](int col, int limb) { return (limb * 2 + col) * kN; };

  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const int B4 = 32 * w + lane;
    tw4s[0 * kThreads + T] = __ldg(A.tw.tw9 + 2 * B4);
    tw4s[1 * kThreads + T] = __ldg(A.tw.tw9 + 2 * B4 + 1);
    tw4s[2 * kThreads + T] = __ldg(A.tw.tw10c + 2 * B4);
    tw4s[3 * kThreads + T] = __ldg(A.tw.tw10c + 2 * B4 + 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_s;
  const uint32_t tsp = tmem_base + ((uint32_t)((w & 3) * 32) << 16) + 128 * (w >> 2);  // rows 0..3
  const int P0 = 256 * w + lane;
  long long phase_t0 = A.phase_cycles ? clock64() : 0;

  for (int item = blockIdx.x; item < A.n_items; item += gridDim.x) {
    int* dst = A.dst + (size_t)item * A.ct_stride;
    const int* src;
    {
      long idx = item;
      if (A.src_div > 0) idx = item / A.src_div;
      else if (A.src_mod > 0) { int r = item % A.src_mod; idx = A.src_map ? A.src_map[r] : r; }
      src = A.src + idx * A.ct_stride;
    }
    const size_t mat_off = A.mat_div > 0 ? (size_t)(item / A.mat_div) * A.mat_stride : 0;

    for (int step = 0; step < A.n_steps; step++) {
      const double2* G = A.mat[step] + mat_off;
      const int* xin = step > 0 ? dst : src;
      PHASE_TICK(0);
      // --------------------------- forward transforms ------------------------------
      {
        int nx[16];
        auto load_row = [&](int rho, int (&v)[16]) {
          const int* p = xin + CT(rho & 1, rho >> 1);
#pragma unroll
          for (int m = 0; m < 8; m++) { v[m] = p[T + 256 * m]; v[m + 8] = p[T + 256 * m + kM]; }
        };
        load_row(0, nx);
#pragma unroll 1
        for (int rho = 0; rho < NR; rho++) {
          double2 x[8];
#pragma unroll
          for (int m = 0; m < 8; m++) x[m] = make_double2((double)nx[m], (double)nx[m + 8]);
          if (rho + 1 < NR) load_row(rho + 1, nx);
          fwd_pass1_store(x, work, T);
          __syncthreads();
          {
            double2* base = work + 256 * w;
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = base[S1(lane + 32 * m)];
            radix8_fwd<true>(x, c_tw_lo[8 + w], c_tw_lo[16 + 2 * w], c_tw_lo[32 + 4 * w], c_tw_lo[32 + 4 * w + 2]);
#pragma unroll
            for (int m = 0; m < 8; m++) base[S1(lane + 32 * m)] = x[m];
            __syncwarp();
            const int qr = 32 * (lane >> 2) + (lane & 3);
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = base[S1(qr + 4 * m)];
            {
              const Tw3 t = load_tw3(w, lane);
              radix8_fwd<true>(x, t.a, t.b, t.c, t.d);
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 8; m++) base[S2(qr + 4 * m)] = x[m];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = base[S2(8 * lane + j)];
            {
              const double2 b4a = tw4s[T], b4b = tw4s[kThreads + T], c4a = tw4s[2 * kThreads + T], c4b = tw4s[3 * kThreads + T];
              bf(x[0], x[2], b4a); bf(x[1], x[3], b4a); bf(x[4], x[6], b4b); bf(x[5], x[7], b4b);
              bf(x[0], x[1], c4a); bf(x[2], x[3], mul_i(c4a));
              bf(x[4], x[5], c4b); bf(x[6], x[7], mul_i(c4b));
            }
          }
          if (rho < 4) {
            const double2 lo[4] = {x[0], x[1], x[2], x[3]};
            const double2 hi[4] = {x[4], x[5], x[6], x[7]};
            tm_st4(tsp + 32 * rho, lo);
            tm_st4(tsp + 32 * rho + 16, hi);
          } else {
#pragma unroll
            for (int j = 0; j < 8; j++) rows_s[(size_t)(rho - 4) * kM + P0 + 32 * j] = x[j];
          }
          __syncthreads();  // `work` is reused by the next row
        }
        tm_wait_st();
      }
      PHASE_TICK(2);

      // --------------- contraction + inverse transform + epilogue ------------------
#pragma unroll 1
      for (int co = 0; co < 2; co++) {
        int carry[16];
#pragma unroll
        for (int q = 0; q < 16; q++) carry[q] = 0;
#pragma unroll 1
        for (int l = LOUT - 1; l >= 0; l--) {
          const int o = co * LOUT + l;
          double2 cur[8];
#pragma unroll
          for (int j = 0; j < 8; j++) cur[j] = make_double2(0.0, 0.0);
#pragma unroll 1
          for (int rho = 0; rho < NR; rho++) {
            const double2* gp = G + ((size_t)rho * NOUT + o) * kM + P0;
            double2 g[8];
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] = __ldg(gp + 32 * j);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              double2 a[4];
              if (rho < 4) {
                tm_ld4(tsp + 32 * rho + 16 * h, a);
              } else {
#pragma unroll
                for (int j = 0; j < 4; j++) a[j] = rows_s[(size_t)(rho - 4) * kM + P0 + 32 * (4 * h + j)];
              }
#pragma unroll
              for (int j = 0; j < 4; j++) {
                cur[4 * h + j].x = fma(a[j].x, g[4 * h + j].x, fma(-a[j].y, g[4 * h + j].y, cur[4 * h + j].x));
                cur[4 * h + j].y = fma(a[j].x, g[4 * h + j].y, fma(a[j].y, g[4 * h + j].x, cur[4 * h + j].y));
              }
            }
          }
          PHASE_TICK(3);
          // inverse transform (pass-4 twiddles from shared, pass-3 from constant memory)
          {
            double2 (&x)[8] = cur;
            double2* wb = work + 256 * w;
            {
              const double2 b4a = tw4s[T], b4b = tw4s[kThreads + T], c4a = tw4s[2 * kThreads + T], c4b = tw4s[3 * kThreads + T];
              ibf(x[0], x[1], c4a); ibf(x[2], x[3], mul_i(c4a));
              ibf(x[4], x[5], c4b); ibf(x[6], x[7], mul_i(c4b));
              ibf(x[0], x[2], b4a); ibf(x[1], x[3], b4a); ibf(x[4], x[6], b4b); ibf(x[5], x[7], b4b);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) wb[S2(8 * lane + j)] = x[j];
            __syncwarp();
            const int qr = 32 * (lane >> 2) + (lane & 3);
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = wb[S2(qr + 4 * m)];
            {
              const Tw3 t = load_tw3(w, lane);
              radix8_inv<true>(x, t.a, t.b, t.c, t.d);
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 8; m++) wb[S1(qr + 4 * m)] = x[m];
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = wb[S1(lane + 32 * m)];
            radix8_inv<true>(x, c_tw_lo[8 + w], c_tw_lo[16 + 2 * w], c_tw_lo[32 + 4 * w], c_tw_lo[32 + 4 * w + 2]);
#pragma unroll
            for (int m = 0; m < 8; m++) wb[S1(lane + 32 * m)] = x[m];
            __syncthreads();
#pragma unroll
            for (int m = 0; m < 8; m++) x[m] = work[S1(T + 256 * m)];
            __syncthreads();
            radix8_inv<true>(x, c_tw_lo[1], c_tw_lo[2], c_tw_lo[4], c_tw_lo[6]);
          }
          PHASE_TICK(4);
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const int i = T + 256 * (q & 7) + (q >> 3) * kM;
            const double v = (q < 8) ? cur[q & 7].x : cur[q & 7].y;
            const long long t = __double2ll_rn(v) + (long long)carry[q];
            const int c = (int)((t + 65536) >> kK);
            const int dg = (int)t - (c << kK);
            carry[q] = c;
            if (l < LRES) dst[CT(co, l) + i] = dg;
          }
          PHASE_TICK(5);
        }
      }
    }  // steps
    __syncthreads();
    PHASE_TICK(6);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

constexpr size_t kKs2Smem = (size_t)kM * sizeof(double2) + (size_t)2 * kN * sizeof(long long) + 16;

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) k_ks2(const VmpArgs A) {
  static_assert(MODE == MODE_TRACE || MODE == MODE_COMBINE2, "key-switch modes only");
  constexpr int R = 3, LOUT = 4, LRES = 3, NOUT = 2 * LOUT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* work = reinterpret_cast<double2*>(smem_raw);
  long long* xp = reinterpret_cast<long long*>(work + kM);  // [2 cols][N] packed limbs
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(xp + 2 * kN);

  const int T = threadIdx.x, w = T >> 5, lane = T & 31;
  auto CT = [](int col, int limb) { return (limb * 2 + col) * kN; };

  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_s;
  // this thread's 128 columns: spectra rows at +0/+32/+64, twiddles at +96 (pass 3) / +112 (pass 4)
  const uint32_t tsp = tmem_base + ((uint32_t)((w & 3) * 32) << 16) + 128 * (w >> 2);
  const uint32_t ttw = tsp + 96;
  {
    const Tw34 t = load_tw34(A.tw, w, lane);
    const double2 p3[4] = {t.a3, t.b3, t.c3, t.d3};
    const double2 p4[4] = {t.b4a, t.b4b, t.c4a, t.c4b};
    tm_st4(ttw, p3);
    tm_st4(ttw + 16, p4);
    tm_wait_st();
  }
  const int P0 = 256 * w + lane;
  long long phase_t0 = A.phase_cycles ? clock64() : 0;

  for (int item = blockIdx.x; item < A.n_items; item += gridDim.x) {
    int* dst = A.dst + (size_t)item * A.ct_stride;
    int* scr1 = A.scratch ? A.scratch + (size_t)blockIdx.x * A.ct_stride : nullptr;  // COMBINE2: S
    const int* src;
    {
      long idx = item;
      if (MODE == MODE_COMBINE2) idx = 2L * item;
      else if (A.src_div > 0) idx = item / A.src_div;
      else if (A.src_mod > 0) { int r = item % A.src_mod; idx = A.src_map ? A.src_map[r] : r; }
      src = A.src + idx * A.ct_stride;
    }
    const size_t mat_off = A.mat_div > 0 ? (size_t)(item / A.mat_div) * A.mat_stride : 0;

    for (int step = 0; step < A.n_steps; step++) {
      const double2* G = A.mat[step] + mat_off;
      const int ginv = A.gal_inv[step];

      // ------------------------------ prologue ------------------------------------
      // (fusing the next step's rsh into the digit loop was measured: the extra 64-bit field
      // inserts cost more than the prologue they save)
      if (MODE == MODE_TRACE) {
        int rk = A.rot_const;
        if (A.rot_mod > 0) rk += A.rot_mul * (item % A.rot_mod);
        rk &= (2 * kN - 1);
#pragma unroll 4
        for (int m = 0; m < 16; m++) {
          const int i = T + 256 * m;
#pragma unroll
          for (int col = 0; col < 2; col++) {
            int a0, a1, a2;
            if (step == 0) {
              bool neg;
              const int j = rot_index(i, 2 * kN - rk, neg);
              a0 = src[CT(col, 0) + j]; a1 = src[CT(col, 1) + j]; a2 = src[CT(col, 2) + j];
              if (neg) { a0 = -a0; a1 = -a1; a2 = -a2; }
            } else {
              const long long wd = xp[col * kN + i];
              a0 = unpack3(wd, 0); a1 = unpack3(wd, 1); a2 = unpack3(wd, 2);
            }
            int d0, d1, d2;
            rsh1_3(a0, a1, a2, d0, d1, d2);
            xp[col * kN + i] = pack3(d0, d1, d2);
          }
        }
      } else {
        const int* a = src;
        const int* b = src + A.ct_stride;
        const int tt = A.rot_const;
        // loads staged four positions at a time ahead of the stores (see k_vmp)
#pragma unroll 1
        for (int mc = 0; mc < 16; mc += 4) {
          int av[4][2][3], bv[4][2][3];
          bool ng[4];
#pragma unroll
          for (int mm = 0; mm < 4; mm++) {
            const int i = T + 256 * (mc + mm);
            const int j = rot_index(i, tt, ng[mm]);
#pragma unroll
            for (int col = 0; col < 2; col++)
#pragma unroll
              for (int l = 0; l < 3; l++) { av[mm][col][l] = a[CT(col, l) + j]; bv[mm][col][l] = b[CT(col, l) + i]; }
          }
#pragma unroll
          for (int mm = 0; mm < 4; mm++) {
            const int i = T + 256 * (mc + mm);
#pragma unroll
            for (int col = 0; col < 2; col++) {
              int a0 = av[mm][col][0], a1 = av[mm][col][1], a2 = av[mm][col][2];
              if (ng[mm]) { a0 = -a0; a1 = -a1; a2 = -a2; }
              const int b0 = bv[mm][col][0], b1 = bv[mm][col][1], b2 = bv[mm][col][2];
              int d0, d1, d2;
              rsh1_3(a0 - b0, a1 - b1, a2 - b2, d0, d1, d2);
              xp[col * kN + i] = pack3(d0, d1, d2);
              rsh1_3(a0 + b0, a1 + b1, a2 + b2, d0, d1, d2);
              scr1[CT(col, 0) + i] = d0; scr1[CT(col, 1) + i] = d1; scr1[CT(col, 2) + i] = d2;
            }
          }
        }
      }
      __syncthreads();
      PHASE_TICK(0);

      // --------------------------- forward transforms ------------------------------
      {
        // phi_g(x) mask words for this thread's 16 input positions, gathered once for all limbs
        long long mw[16];
        unsigned sgn = 0;
#pragma unroll
        for (int m = 0; m < 8; m++) {
          const int i = T + 256 * m;
          const int u = (i * ginv) & (2 * kN - 1);
          const int u2 = (u + kM * (ginv & 3)) & (2 * kN - 1);
          mw[m] = xp[kN + (u & (kN - 1))];
          mw[m + 8] = xp[kN + (u2 & (kN - 1))];
          sgn |= (u >= kN ? 1u : 0u) << m;
          sgn |= (u2 >= kN ? 1u : 0u) << (m + 8);
        }
#pragma unroll 1
        for (int rho = 0; rho < R; rho++) {
          double2 x[8];
#pragma unroll
          for (int m = 0; m < 8; m++) {
            const int v = unpack3(mw[m], rho), v2 = unpack3(mw[m + 8], rho);
            x[m] = make_double2((double)((sgn >> m) & 1 ? -v : v), (double)((sgn >> (m + 8)) & 1 ? -v2 : v2));
          }
          fwd_pass1_store(x, work, T);
          __syncthreads();
          fwd_warp_passes2(work, w, lane, ttw, x);
          {
            const double2 lo[4] = {x[0], x[1], x[2], x[3]};
            const double2 hi[4] = {x[4], x[5], x[6], x[7]};
            tm_st4(tsp + 32 * rho, lo);
            tm_st4(tsp + 32 * rho + 16, hi);
          }
          __syncthreads();  // `work` is reused by the next row
        }
        tm_wait_st();
      }
      PHASE_TICK(2);

      // --------------- contraction + inverse transform + epilogue ------------------
#pragma unroll 1
      for (int co = 0; co < 2; co++) {
        int carry[16], carry2[16];
#pragma unroll
        for (int q = 0; q < 16; q++) { carry[q] = 0; carry2[q] = 0; }
#pragma unroll 1
        for (int l = LOUT - 1; l >= 0; l--) {
          const int o = co * LOUT + l;
          double2 cur[8];
#pragma unroll
          for (int j = 0; j < 8; j++) cur[j] = make_double2(0.0, 0.0);
#pragma unroll 1
          for (int rho = 0; rho < R; rho++) {
            const double2* gp = G + ((size_t)rho * NOUT + o) * kM + P0;
            double2 g[8];
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] = __ldg(gp + 32 * j);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              double2 a[4];
              tm_ld4(tsp + 32 * rho + 16 * h, a);
#pragma unroll
              for (int j = 0; j < 4; j++) {
                cur[4 * h + j].x = fma(a[j].x, g[4 * h + j].x, fma(-a[j].y, g[4 * h + j].y, cur[4 * h + j].x));
                cur[4 * h + j].y = fma(a[j].x, g[4 * h + j].y, fma(a[j].y, g[4 * h + j].x, cur[4 * h + j].y));
              }
            }
          }
          PHASE_TICK(3);
          const bool has_small = l < R;
          int xnat[16];
          inv_transform2(cur, work, T, w, lane, ttw, [&]() {
            if (MODE == MODE_TRACE) {
#pragma unroll
              for (int q = 0; q < 16; q++) {
                const int i = T + 256 * (q & 7) + (q >> 3) * kM;
                bool neg;
                const int u = auto_index(i, ginv, neg);
                const int v = (co == 0 && has_small) ? unpack3(xp[u], l) : 0;
                xnat[q] = neg ? -v : v;
              }
            }
          });
          PHASE_TICK(4);
          int sv[16];  // COMBINE2: S digits of this output, requested together
          if (MODE == MODE_COMBINE2 && l < LRES) {
#pragma unroll
            for (int q = 0; q < 16; q++) sv[q] = scr1[CT(co, l) + T + 256 * (q & 7) + (q >> 3) * kM];
          }
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const int i = T + 256 * (q & 7) + (q >> 3) * kM;
            const double v = (q < 8) ? cur[q & 7].x : cur[q & 7].y;
            long long big = __double2ll_rn(v);
            bool neg = false;
            long long wd = 0;
            if (MODE == MODE_TRACE) {
              big += (long long)xnat[q];
              if (A.sign < 0) big = -big;
              if (has_small) {
                wd = xp[co * kN + i];
                big += (long long)unpack3(wd, l);
              }
            } else {
              const int u = auto_index(i, ginv, neg);
              if (neg) big = -big;
              if (co == 0 && has_small) big += (long long)unpack3(xp[u], l);
            }
            const long long t = big + (long long)carry[q];
            const int c = (int)((t + 65536) >> kK);
            const int dg = (int)t - (c << kK);
            carry[q] = c;
            if (l < LRES) {
              if (MODE == MODE_TRACE) {
                xp[co * kN + i] = repack3(wd, l, dg);
              } else {
                const int y = neg ? -dg : dg;
                const int t2 = sv[q] - y + carry2[q];
                const int dg2 = sext17i(t2);
                carry2[q] = (t2 - dg2) >> kK;
                bool rneg;
                const int dd = rot_index(i, A.rot_const, rneg);
                dst[CT(co, l) + dd] = rneg ? -dg2 : dg2;
              }
            }
          }
          PHASE_TICK(5);
        }
      }
      if (MODE == MODE_TRACE) __syncthreads();
    }  // steps

    if (MODE == MODE_TRACE) {
#pragma unroll 4
      for (int m = 0; m < 16; m++) {
        const int i = T + 256 * m;
#pragma unroll
        for (int col = 0; col < 2; col++) {
          const long long wd = xp[col * kN + i];
          dst[CT(col, 0) + i] = unpack3(wd, 0);
          dst[CT(col, 1) + i] = unpack3(wd, 1);
          dst[CT(col, 2) + i] = unpack3(wd, 2);
        }
      }
    }
    __syncthreads();
    PHASE_TICK(6);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

}  // namespace fheram
