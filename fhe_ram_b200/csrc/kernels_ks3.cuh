// kernels_ks3.cuh -- key-switch kernels (trace / one-sided packer chains, two-sided packer combine)
// with the integer side of the operation done on 51-bit WORDS instead of per-limb digits.
//
// A normalised 3-limb coefficient (balanced base-2^17 digits a0,a1,a2 in [-2^16, 2^16)) is kept
// on chip as the biased word
//     U = sum_l (a_l + 2^16) << 17 (2 - l)          (51 bits; the digits are plain bit fields)
// i.e. U = X + bias with X = a0 2^34 + a1 2^17 + a2 and bias = 2^16 + 2^33 + 2^50.  Because the
// balanced digits of a value are unique and the top carry of vec_znx_big_normalize is dropped,
// "normalise" is simply "reduce mod 2^51", and everything Poulpy does limb by limb between two
// transforms collapses to a few 64-bit additions per coefficient:
//   * glwe_rsh(1)          X -> ceil(X / 2)                 (kernels.cuh: rsh1_3 computes its digits)
//   * vec_znx_big_normalize of the 4-limb product r_0..r_3 plus small operands
//                          V = sum_{l<3} r_l << 17 (2 - l)  +  floor((r_3 + 2^16) / 2^17)  +  smalls
//   * the inverse transform's rounding: v + 1.5 2^52 puts round(v) mod 2^51 in the low mantissa
//     bits (one DFMA, no F2I), and digits enter the forward transform as (2^52 + field) - (2^52 +
//     2^16) (one DADD, no I2F).
// This replaces k_ks2's digit carry chains, packed-field inserts and the separate rsh pass (about
// 60 K of its 126 K warp instructions per key switch).  Transforms, contraction, tensor-memory use
// and the two-CTAs-per-SM footprint are those of k_ks2; results are bit-identical (same integers).
// Reference semantics: GLWE trace / GLWEPacker::combine as issued from src/ram.rs:435-457,540,572,
// 616-629 (Poulpy 0.3.2 glwe_trace, glwe_packer; oracle/fheram_oracle.c glwe_trace, pack_combine).
#pragma once
#include "kernels_ks2.cuh"
#include "transform_pad.cuh"

// timing ablations for tools/ablate.sh (results are garbage when non-zero): 1 no matrix loads,
// 2 no inverse transforms, 4 no forward transforms
#ifndef FHERAM_ABL
#define FHERAM_ABL 0
#endif

namespace fheram {

constexpr unsigned long long kBias51 = (1ull << 16) | (1ull << 33) | (1ull << 50);
constexpr unsigned long long kMask51 = (1ull << 51) - 1;
constexpr double kMagic52 = 6755399441055744.0;  // 1.5 * 2^52

// value of three (not necessarily canonical) limbs
__device__ __forceinline__ long long limbs_value(int a0, int a1, int a2) {
  return ((long long)a0 << 34) + ((long long)a1 << 17) + (long long)a2;
}
// biased word of ceil(X / 2) for an arbitrary value X (wraps mod 2^51 like the dropped top carry)
__device__ __forceinline__ unsigned long long rsh1_word(long long X) {
  return ((unsigned long long)((X + 1) >> 1) + kBias51) & kMask51;
}
// the same for a canonical word (value U - bias in the digit range: cannot wrap)
__device__ __forceinline__ unsigned long long rsh1_canon(unsigned long long U) {
  return ((U + 1) >> 1) + (kBias51 >> 1);
}
__device__ __forceinline__ int word_digit(unsigned long long U, int l) {
  return (int)((uint32_t)(U >> (17 * (2 - l))) & 0x1ffffu) - 65536;
}
// +/-(field - 2^16) as a double; negbit = 0 or 0x80000000
__device__ __forceinline__ double field_f64(uint32_t f, uint32_t negbit) {
  const double x = __hiloint2double((int)(0x43300000u | negbit), (int)f);       // +/-(2^52 + f)
  const double c = __hiloint2double((int)(0xC3300000u ^ negbit), 65536);        // -/+(2^52 + 2^16)
  return x + c;
}
// integer bits of a double produced by "+ kMagic52": congruent to round(v) mod 2^51
__device__ __forceinline__ unsigned long long magic_bits(double t) {
  return ((unsigned long long)(uint32_t)__double2hiint(t) << 32) | (uint32_t)__double2loint(t);
}

// ======================================================================================
// k_ext3: k_ext2 with the padded exchange buffer of transform_pad.cuh (addresses are base + immediate
// instead of XOR swizzles), three instead of four pass-4 twiddle columns, and digits converted with
// one DADD instead of I2F.  Same arithmetic and results.
// ======================================================================================
// int32 -> double as (2^52 + 2^31 + v) - (2^52 + 2^31)
__device__ __forceinline__ double int_f64(int v) {
  return __hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)) - 4503601774854144.0;
}

// 36 (padded exchange buffer) + 64 (two spectra) + 8 (pass-4 twiddles b4a, c4a; b4b = i b4a, c4b = e^(i pi/4) c4a)
// + 4 (pass-3 twiddle table) = 112 KiB
constexpr size_t kExt3Smem = (size_t)kWorkPad * sizeof(double2) + (size_t)2 * kM * sizeof(double2) +
                             (size_t)3 * kThreads * sizeof(double2) + 16;
// w * e^(i pi/4), with exactly the roundings the host uses to build the odd entries of tw10c
__device__ __forceinline__ double2 mul_e8(double2 w) {
  const double r = 0.70710678118654757;
  return make_double2(__dmul_rn(__dadd_rn(w.x, -w.y), r), __dmul_rn(__dadd_rn(w.x, w.y), r));
}

__global__ void __launch_bounds__(kThreads, 2) k_ext3(const VmpArgs A) {
  constexpr int NR = 6, LOUT = 4, LRES = 3, NOUT = 2 * LOUT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* work = reinterpret_cast<double2*>(smem_raw);
  double2* rows_s = work + kWorkPad;           // spectra of rows 4 and 5
  double2* tw4s = rows_s + 2 * kM;             // [2][256] pass-4 twiddles b4a, c4a (thread-private columns)
  double2* tw3s = tw4s + 2 * kThreads;         // [256] pass-3 twiddle table zeta(6,B) | zeta(7,2B) | zeta(8,2k)
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(tw3s + kThreads);

  const int T = threadIdx.x, w = T >> 5, lane = T & 31;
  auto CT = [](int col, int limb) { return (limb * 2 + col) * kN; };

  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  BufSync bs{smem_u32(tmem_base_s + 2), 0u};  // exchange-buffer mbarrier in the same 16-byte slot
  if (T == 0) buf_init(bs.mbar);
  {
    const int B4 = 32 * w + lane;
    tw4s[0 * kThreads + T] = __ldg(A.tw.tw9 + 2 * B4);    // zeta(9, 2 B4 + 1) = i * this one
    tw4s[1 * kThreads + T] = __ldg(A.tw.tw10c + 2 * B4);  // zeta(10, 4 B4 + 2) = e^(i pi/4) * this one
    tw3s[T] = __ldg(A.tw.tw6 + T);                        // tw6 | tw7c | tw8c are contiguous
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_s;
  buf_release(bs);  // the buffer starts out free
  const uint32_t tsp = tmem_base + ((uint32_t)((w & 3) * 32) << 16) + 128 * (w >> 2);  // rows 0..3
  const int P0 = 256 * w + lane;
  const PadAddr pa = pad_addr(work, T, w, lane);
  // pass-3 twiddles from a 4 KiB shared table (8 distinct entries per warp, broadcast within quads):
  // k_ext2 read them from the constant bank, where 8 distinct addresses per warp serialise in the
  // address-divergence unit (ncu: ADU pipe 37 %)
  auto tw3 = [&]() {
    const int B = 8 * w + (lane >> 2);
    return Tw4x{tw3s[B], tw3s[64 + B], tw3s[128 + 2 * B], tw3s[128 + 2 * B + 1]};
  };
  auto tw4 = [&]() {
    const double2 b4a = tw4s[T], c4a = tw4s[kThreads + T];
    return Tw4x{b4a, mul_i(b4a), c4a, mul_e8(c4a)};
  };
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
          // per-thread pointer made opaque: with a (uniform base + thread offset) split the compiler
          // re-materialises the base in uniform registers for every access (two R2UR per load)
          const int* p = xin + CT(rho & 1, rho >> 1) + T;
          asm volatile("" : "+l"(p));
#pragma unroll
          for (int m = 0; m < 8; m++) { v[m] = p[256 * m]; v[m + 8] = p[256 * m + kM]; }
        };
        load_row(0, nx);
#pragma unroll 1
        for (int rho = 0; rho < NR; rho++) {
          double2 x[8];
#pragma unroll
          for (int m = 0; m < 8; m++) x[m] = make_double2(int_f64(nx[m]), int_f64(nx[m + 8]));
          if (rho + 1 < NR) load_row(rho + 1, nx);
          fwd_pass1_store_p(x, pa, bs);
          __syncthreads();
          fwd_warp_passes_p(pa, w, tw3, tw4, x, bs);
          if (rho < 4) {
            const double2 lo[4] = {x[0], x[1], x[2], x[3]};
            const double2 hi[4] = {x[4], x[5], x[6], x[7]};
            tm_st4(tsp + 32 * rho, lo);
            tm_st4(tsp + 32 * rho + 16, hi);
          } else {
#pragma unroll
            for (int j = 0; j < 8; j++) rows_s[(size_t)(rho - 4) * kM + P0 + 32 * j] = x[j];
          }
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
            for (int j = 0; j < 8; j++) g[j] = ldg_pinned(gp + 32 * j);
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
          inv_transform_p(cur, pa, w, tw3, tw4, bs);
          PHASE_TICK(4);
          int* dp = dst + CT(co, l < LRES ? l : 0) + T;
          asm volatile("" : "+l"(dp));
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const double v = (q < 8) ? cur[q & 7].x : cur[q & 7].y;
            const long long t = __double2ll_rn(v) + (long long)carry[q];
            const int c = (int)((t + 65536) >> kK);
            const int dg = (int)t - (c << kK);
            carry[q] = c;
            if (l < LRES) dp[256 * (q & 7) + (q >> 3) * kM] = dg;
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


constexpr size_t kKs3Smem = (size_t)kWorkPad * sizeof(double2) + (size_t)2 * kN * sizeof(long long) + 16;

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) k_ks3(const VmpArgs A) {
  static_assert(MODE == MODE_TRACE || MODE == MODE_COMBINE2, "key-switch modes only");
  constexpr int R = 3, LOUT = 4, NOUT = 2 * LOUT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* work = reinterpret_cast<double2*>(smem_raw);  // padded exchange buffer (transform_pad.cuh)
  unsigned long long* xp = reinterpret_cast<unsigned long long*>(work + kWorkPad);  // [2 cols][N] words
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(xp + 2 * kN);

  const int T = threadIdx.x, w = T >> 5, lane = T & 31;
  auto CT = [](int col, int limb) { return (limb * 2 + col) * kN; };

  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  BufSync bs{smem_u32(tmem_base_s + 2), 0u};  // exchange-buffer mbarrier in the same 16-byte slot
  if (T == 0) buf_init(bs.mbar);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_s;
  buf_release(bs);  // the buffer starts out free
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
  const PadAddr pa = pad_addr(work, T, w, lane);
  auto tw3 = [&]() { double2 t[4]; tm_ld4(ttw, t); return Tw4x{t[0], t[1], t[2], t[3]}; };
  auto tw4 = [&]() { double2 t[4]; tm_ld4(ttw + 16, t); return Tw4x{t[0], t[1], t[2], t[3]}; };
  const double sgn_d = (MODE == MODE_TRACE && A.sign < 0) ? -1.0 : 1.0;
  const uint32_t sgn_bit = (MODE == MODE_TRACE && A.sign < 0) ? 1u : 0u;
  long long phase_t0 = A.phase_cycles ? clock64() : 0;

  for (int item = blockIdx.x; item < A.n_items; item += gridDim.x) {
    int* dst = A.dst + (size_t)item * A.ct_stride;
    // COMBINE2: words of S = rsh1(a X^-t + b), [2 cols][N], in this CTA's global scratch
    unsigned long long* sw = A.scratch
        ? reinterpret_cast<unsigned long long*>(A.scratch + (size_t)blockIdx.x * A.ct_stride) : nullptr;
    const int* src;
    {
      long idx = item;
      if (MODE == MODE_COMBINE2) idx = 2L * item;
      else if (A.src_div > 0) idx = item / A.src_div;
      else if (A.src_mod > 0) { int r = item % A.src_mod; idx = A.src_map ? A.src_map[r] : r; }
      src = A.src + idx * A.ct_stride;
    }
    const size_t mat_off = A.mat_div > 0 ? (size_t)(item / A.mat_div) * A.mat_stride : 0;

    // ------------------------------ prologue (once per item) ---------------------------
    if (MODE == MODE_TRACE) {
      // x = rsh1(src * X^rk)
      int rk = A.rot_const;
      if (A.rot_mod > 0) rk += A.rot_mul * (item % A.rot_mod);
      rk &= (2 * kN - 1);
#pragma unroll 4
      for (int m = 0; m < 16; m++) {
        const int i = T + 256 * m;
        bool neg;
        const int j = rot_index(i, 2 * kN - rk, neg);
#pragma unroll
        for (int col = 0; col < 2; col++) {
          long long X = limbs_value(src[CT(col, 0) + j], src[CT(col, 1) + j], src[CT(col, 2) + j]);
          if (neg) X = -X;
          xp[col * kN + i] = rsh1_word(X);
        }
      }
    } else {
      // a1 = a X^-t;  D = rsh1(a1 - b) -> xp;  S = rsh1(a1 + b) -> sw
      const int* a = src;
      const int* b = src + A.ct_stride;
      const int tt = A.rot_const;
#pragma unroll 1
      for (int mc = 0; mc < 16; mc += 4) {
        int av[4][2][3], bv[4][2][3];
        bool ng[4];
#pragma unroll
        for (int mm = 0; mm < 4; mm++) {
          const int i = T + 256 * (mc + mm);
          const int j = rot_index(i, tt, ng[mm]);  // (a X^-t)[i] = +/- a[(i + t) mod 2N]
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
            long long Xa = limbs_value(av[mm][col][0], av[mm][col][1], av[mm][col][2]);
            if (ng[mm]) Xa = -Xa;
            const long long Xb = limbs_value(bv[mm][col][0], bv[mm][col][1], bv[mm][col][2]);
            xp[col * kN + i] = rsh1_word(Xa - Xb);
            sw[col * kN + i] = rsh1_word(Xa + Xb);
          }
        }
      }
    }
    __syncthreads();
    PHASE_TICK(0);

    for (int step = 0; step < A.n_steps; step++) {
      const double2* G = A.mat[step] + mat_off;
      const int ginv = A.gal_inv[step];
      const bool last = step + 1 == A.n_steps;
      // automorphism source of this thread's positions i_q = T + 256 (q & 7) + 2048 (q >> 3):
      // e_q = i_q * ginv mod 2N (index e_q mod N, sign e_q >= N); e_q = e0 + (q & 7) d1 + (q >> 3) d2
      const int e0 = (T * ginv) & (2 * kN - 1);
      const int d1 = (256 * ginv) & (2 * kN - 1);
      const int d2 = kM * (ginv & 3);
      unsigned sgn = 0;  // bit q: phi_g flips the sign at position i_q

      // --------------------------- forward transforms ------------------------------
      {
        // phi_g(x) mask digits of the 16 input positions; the words are gathered again for every
        // limb (holding them across the three transforms costs 32 registers and spills)
#pragma unroll 1
        for (int rho = 0; rho < R; rho++) {
          double2 x[8];
          // digit rho = bits [17 (2 - rho), +17) of the word = (funnel(lo, hi, s1) >> s2) & (2^17 - 1)
          // (shift amounts kept loop-variant on purpose: with per-rho code the compiler hoists all
          // 48 digits out of the loop and spills them)
          const int s1 = rho == 0 ? 31 : (rho == 1 ? 17 : 0);
          const int s2 = rho == 0 ? 3 : 0;
          unsigned sg = 0;
#pragma unroll
          for (int m = 0; m < 8; m++) {
            const int ea = (e0 + m * d1) & (2 * kN - 1);
            const int eb = (ea + d2) & (2 * kN - 1);
            const unsigned long long wa = xp[kN + (ea & (kN - 1))];
            const unsigned long long wb = xp[kN + (eb & (kN - 1))];
            const uint32_t na = ea >= kN ? 0x80000000u : 0u, nb = eb >= kN ? 0x80000000u : 0u;
            sg |= (na >> (31 - m)) | (nb >> (23 - m));
            x[m] = make_double2(
                field_f64((__funnelshift_r((uint32_t)wa, (uint32_t)(wa >> 32), s1) >> s2) & 0x1ffffu, na),
                field_f64((__funnelshift_r((uint32_t)wb, (uint32_t)(wb >> 32), s1) >> s2) & 0x1ffffu, nb));
          }
          sgn = sg;
          if (!(FHERAM_ABL & 4)) {
            fwd_pass1_store_p(x, pa, bs);
            __syncthreads();
            fwd_warp_passes_p(pa, w, tw3, tw4, x, bs);
          }
          {
            const double2 lo[4] = {x[0], x[1], x[2], x[3]};
            const double2 hi[4] = {x[4], x[5], x[6], x[7]};
            tm_st4(tsp + 32 * rho, lo);
            tm_st4(tsp + 32 * rho + 16, hi);
          }
        }
        tm_wait_st();
      }
      PHASE_TICK(2);

      // --------------- contraction + inverse transform + word accumulation ---------
      // mask column first: its words are rewritten mid-step, many barriers before the next
      // step gathers them; the body column needs one barrier between its gather and its stores.
#pragma unroll 1
      for (int cc = 0; cc < 2; cc++) {
        const int co = 1 - cc;
        unsigned long long acc[16];
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
            for (int j = 0; j < 8; j++) g[j] = (FHERAM_ABL & 1) ? make_double2(1.0 + j, 0.5) : __ldg(gp + 32 * j);
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
          if (!(FHERAM_ABL & 2)) inv_transform_p(cur, pa, w, tw3, tw4, bs);
          PHASE_TICK(4);
          // cur[m] = phi_g(vmp)[T + 256 m] (+ i * [.. + 2048]); round and accumulate into the word
          if (l == 3) {
            // floor((s r + 2^16) / 2^17) with s the sign frame of the carry chain: the uniform
            // sign of the trace step, or (COMBINE2) the per-position automorphism sign, for which
            // floor((-r + 2^16) / 2^17) = -floor((r + 2^16 - 1) / 2^17)
#pragma unroll
            for (int q = 0; q < 16; q++) {
              const double v = (q < 8) ? cur[q & 7].x : cur[q & 7].y;
              double t;
              if (MODE == MODE_TRACE) t = fma(v, sgn_d, kMagic52 + 65536.0);
              else t = v + __hiloint2double(0x43380000, (int)(65536u - ((sgn >> q) & 1u)));
              const int c3 = (int)__funnelshift_r((uint32_t)__double2loint(t), (uint32_t)__double2hiint(t), 17);
              acc[q] = (unsigned long long)(long long)c3;
            }
          } else if (l == 2) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
              const double v = (q < 8) ? cur[q & 7].x : cur[q & 7].y;
              acc[q] += magic_bits(fma(v, sgn_d, kMagic52));
            }
          } else if (l == 1) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
              const double v = (q < 8) ? cur[q & 7].x : cur[q & 7].y;
              acc[q] += magic_bits(fma(v, sgn_d, kMagic52)) << 17;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 16; q++) {
              const double v = (q < 8) ? cur[q & 7].x : cur[q & 7].y;
              const double t = fma(v, sgn_d, kMagic52);
              acc[q] += (unsigned long long)((uint32_t)__double2loint(t) << 2) << 32;
            }
          }
          PHASE_TICK(5);
        }
        // ------------------------ combine with the small operands ------------------------
        if (MODE == MODE_TRACE) {
          // x <- x + s phi_g(KS(x)):  V = s (R + phi(x_body)) + c3 + x   (acc = s R + c3)
          if (co == 1) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
              const int i = T + 256 * (q & 7) + (q >> 3) * kM;
              const unsigned long long U = (acc[q] + xp[kN + i]) & kMask51;
              xp[kN + i] = last ? U : rsh1_canon(U);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 16; q++) {
              const int i = T + 256 * (q & 7) + (q >> 3) * kM;
              const int e = (e0 + (q & 7) * d1 + (q >> 3) * d2) & (2 * kN - 1);
              const unsigned long long b = xp[e & (kN - 1)] - kBias51;
              const bool ng = (((sgn >> q) & 1u) ^ sgn_bit) != 0;
              acc[q] = (acc[q] + (ng ? 0ull - b : b) + xp[i]) & kMask51;
            }
            __syncthreads();  // every gather of the old body column precedes its stores
#pragma unroll
            for (int q = 0; q < 16; q++) {
              const int i = T + 256 * (q & 7) + (q >> 3) * kM;
              xp[i] = last ? acc[q] : rsh1_canon(acc[q]);
            }
          }
        } else {
          // y = phi_g(normalize(KS(D)));  out = normalize(S - y) X^t
          //   word = S - (R + k3) - sigma (D_body[u] - bias)      (acc = R + k3)
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const int i = T + 256 * (q & 7) + (q >> 3) * kM;
            unsigned long long U = sw[co * kN + i] - acc[q];
            if (co == 0) {
              const int e = (e0 + (q & 7) * d1 + (q >> 3) * d2) & (2 * kN - 1);
              const unsigned long long b = xp[e & (kN - 1)] - kBias51;
              U -= ((sgn >> q) & 1u) ? 0ull - b : b;
            }
            U &= kMask51;
            bool rneg;
            const int dd = rot_index(i, A.rot_const, rneg);  // a' * X^t
#pragma unroll
            for (int l = 0; l < 3; l++) {
              const int dg = word_digit(U, l);
              dst[CT(co, l) + dd] = rneg ? -dg : dg;
            }
          }
        }
        PHASE_TICK(5);
      }
    }  // steps

    if (MODE == MODE_TRACE) {
      // own positions only: no barrier needed before reading the words back
#pragma unroll 4
      for (int m = 0; m < 16; m++) {
        const int i = T + 256 * m;
#pragma unroll
        for (int col = 0; col < 2; col++) {
          const unsigned long long U = xp[col * kN + i];
          dst[CT(col, 0) + i] = word_digit(U, 0);
          dst[CT(col, 1) + i] = word_digit(U, 1);
          dst[CT(col, 2) + i] = word_digit(U, 2);
        }
      }
    }
    __syncthreads();  // xp reuse by the next item
    PHASE_TICK(6);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
}

}  // namespace fheram
