/*
 * fheram_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See the header
 * of fheram_oracle.h and oracle/SPEC.md.  PARITY UNPINNED at the limb level (no golden
 * ciphertexts exist in the reference; Poulpy 0.3.2 is absent from /root/reference).
 *
 * Every function cites the reference file:line it restates, or "Poulpy [spec]" when the
 * logic lives in the un-vendored dependency and is restated from its published algorithm.
 */
#include "fheram_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __GLIBC__
#include <malloc.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef uint64_t u64;
typedef unsigned __int128 u128;

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288
#endif

/* ======================================================================================
 * Parameters (src/parameters.rs:11-21) and digit layout (src/base.rs)
 * ==================================================================================== */
void orc_params_snapshot(orc_params *p) {
  memset(p, 0, sizeof(*p));
  p->log_n = 12;                 /* parameters.rs:11 */
  p->base2k = 17;                /* :12 */
  p->k_pt = 3;                   /* :14 */
  p->k_ct = 17 * 3;              /* :15 */
  p->k_addr = 17 * 4;            /* :16 */
  p->k_evk_trace = 17 * 4;       /* :17 */
  p->k_evk_ggsw_inv = 17 * 5;    /* :18 */
  p->n_decomp = 4;               /* :19 */
  for (int i = 0; i < 4; i++) p->decomp_n[i] = 3;
  p->word_size = 4;              /* :20 */
  p->max_addr = 1u << 14;        /* :21 */
}

void orc_params_readme(orc_params *p) { /* README.md:17-34 */
  orc_params_snapshot(p);
  p->k_pt = 9;
  p->max_addr = 1u << 18;
}

uint32_t orc_base1d_max(const int32_t *b, int n) { /* base.rs:10-14 */
  uint32_t m = 1;
  for (int i = 0; i < n; i++) m <<= b[i];
  return m;
}
uint32_t orc_base1d_gap(const int32_t *b, int n, int log_n) { /* base.rs:17-21 */
  uint32_t gap = (uint32_t)log_n;
  for (int i = 0; i < n; i++) gap >>= b[i];
  return 1u << gap;
}
void orc_base1d_decomp(const int32_t *b, int n, uint32_t v, uint8_t *o) { /* base.rs:24-33 */
  int sum = 0;
  for (int i = 0; i < n; i++) {
    o[i] = (uint8_t)((v >> sum) & ((1u << b[i]) - 1));
    sum += b[i];
  }
}
uint32_t orc_base1d_recomp(const int32_t *b, int n, const uint8_t *d) { /* base.rs:36-44 */
  uint32_t v = 0;
  int sum = 0;
  for (int i = 0; i < n; i++) {
    v |= (uint32_t)d[i] << sum;
    sum += b[i];
  }
  return v;
}

int orc_get_base_2d(uint32_t value, const int32_t *base, int n_base, int32_t *lens,
                    int32_t *digits) { /* base.rs:84-108 */
  int n_out = 0;
  uint32_t vm1 = value - 1;
  int bits = vm1 == 0 ? 0 : 32 - __builtin_clz(vm1);
  while (bits != 0) {
    int len = 0;
    for (int i = 0; i < n_base; i++) {
      int b = base[i];
      if (b <= bits) {
        digits[n_out * 8 + len++] = b;
        bits -= b;
      } else {
        if (bits != 0) {
          digits[n_out * 8 + len++] = bits;
          bits = 0;
        }
        break;
      }
    }
    lens[n_out++] = len;
    if (n_out >= 8) break;
  }
  return n_out;
}

uint64_t orc_reverse_bits_msb(uint64_t x, uint32_t n) { /* src/lib.rs:23-26 */
  uint64_t r = 0;
  for (uint32_t i = 0; i < n; i++) r |= ((x >> i) & 1ull) << (n - 1 - i);
  return r;
}

int64_t orc_cast_u8_to_signed(uint8_t v, int bits) { /* examples/fhe-ram.rs:25-32 */
  int shift = 8 - bits;
  return (int64_t)((int8_t)(uint8_t)(v << shift)) >> shift;
}

/* ======================================================================================
 * Source: ChaCha20 keystream (stand-in for poulpy_hal::source::Source, which wraps
 * rand_chacha 0.9 -- Cargo.lock:479-506; stream compatibility with it is NOT claimed).
 * ==================================================================================== */
struct orc_source {
  uint32_t key[8];
  u64 counter;
  uint32_t block[16];
  int pos; /* next unread 32-bit word in block, 16 = empty */
};
#define ROTL32(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define QR(a, b, c, d)                                                                   \
  a += b; d ^= a; d = ROTL32(d, 16); c += d; b ^= c; b = ROTL32(b, 12);                  \
  a += b; d ^= a; d = ROTL32(d, 8);  c += d; b ^= c; b = ROTL32(b, 7);
static void chacha_block(orc_source *s) {
  uint32_t st[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
  for (int i = 0; i < 8; i++) st[4 + i] = s->key[i];
  st[12] = (uint32_t)s->counter;
  st[13] = (uint32_t)(s->counter >> 32);
  st[14] = 0;
  st[15] = 0;
  uint32_t x[16];
  memcpy(x, st, sizeof(x));
  for (int r = 0; r < 10; r++) {
    QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
    QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
    QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
    QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; i++) s->block[i] = x[i] + st[i];
  s->counter++;
  s->pos = 0;
}
orc_source *orc_source_new(const uint8_t seed[32]) {
  orc_source *s = (orc_source *)calloc(1, sizeof(*s));
  for (int i = 0; i < 8; i++)
    s->key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) |
                ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
  s->pos = 16;
  return s;
}
void orc_source_free(orc_source *s) { free(s); }
uint32_t orc_source_next_u32(orc_source *s) {
  if (s->pos >= 16) chacha_block(s);
  return s->block[s->pos++];
}
uint64_t orc_source_next_u64(orc_source *s) {
  u64 lo = orc_source_next_u32(s);
  u64 hi = orc_source_next_u32(s);
  return lo | (hi << 32);
}
void orc_source_fill_bytes(orc_source *s, uint8_t *out, size_t n) {
  size_t i = 0;
  while (i < n) {
    uint32_t w = orc_source_next_u32(s);
    for (int b = 0; b < 4 && i < n; b++, i++) out[i] = (uint8_t)(w >> (8 * b));
  }
}
static double source_f64(orc_source *s) { /* uniform in (0,1] */
  return ((double)(orc_source_next_u64(s) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
}
/* rounded Gaussian sigma, truncated at bound (Poulpy [spec]: add_normal(sigma, 6 sigma)) */
static i64 source_gauss(orc_source *s, double sigma, double bound) {
  for (;;) {
    double u1 = source_f64(s), u2 = source_f64(s);
    double z = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2) * sigma;
    if (fabs(z) <= bound) return (i64)llround(z);
  }
}

/* ======================================================================================
 * Negacyclic transform backends.  A transformed polynomial is n "tfe" words.
 *   exact: NTT modulo a 62-bit prime P == 1 mod 2^14 (results are < 2^50 in magnitude, so
 *          the centred lift is the exact integer negacyclic product).
 *   fft64: z_j = a_j + i a_{j+n/2}, evaluated at the n/2 roots psi^(4k+1), psi = e^{i pi/n}
 *          (SURVEY.md A.2 "Forward DFT of a limb"); inverse divides by n/2 and rounds.
 * ==================================================================================== */
#ifdef ORC_FFT64
typedef double tfe;
const char *orc_backend_name(void) { return "fft64"; }
#else
typedef u64 tfe;
const char *orc_backend_name(void) { return "exact"; }
#endif

struct orc_ctx {
  orc_params p;
  int n, log_n, k;                     /* k = base2k */
  int size_ct, dnum_ct, size_addr, size_evk_trace, dnum_ggsw, size_evk_inv;
  int n_coord, coord_len[8], coord_digits[8][8], n_ggsw, n_glwe;
  i64 gal[32];
  u64 counters[2];
#ifdef ORC_FFT64
  double *tw_re, *tw_im; /* zeta for block b of stage s at index (1<<s)+b */
#else
  u64 P, NPINV, R2, NINV_M;
  u64 *psi_rev_m, *psi_inv_rev_m; /* Montgomery form */
#endif
};

#ifndef ORC_FFT64
static inline u64 mont_mul(const orc_ctx *c, u64 a, u64 b) {
  u128 t = (u128)a * b;
  u64 m = (u64)t * c->NPINV;
  u64 r = (u64)((t + (u128)m * c->P) >> 64);
  return r >= c->P ? r - c->P : r;
}
static u64 mulmod(u64 a, u64 b, u64 p) { return (u64)((u128)a * b % p); }
static u64 powmod(u64 a, u64 e, u64 p) {
  u64 r = 1;
  while (e) {
    if (e & 1) r = mulmod(r, a, p);
    a = mulmod(a, a, p);
    e >>= 1;
  }
  return r;
}
static int is_prime_u64(u64 n) {
  static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return 0;
  for (int i = 0; i < 12; i++) {
    if (n == bases[i]) return 1;
    if (n % bases[i] == 0) return 0;
  }
  u64 d = n - 1;
  int r = 0;
  while ((d & 1) == 0) { d >>= 1; r++; }
  for (int i = 0; i < 12; i++) {
    u64 x = powmod(bases[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int comp = 1;
    for (int j = 1; j < r; j++) {
      x = mulmod(x, x, n);
      if (x == n - 1) { comp = 0; break; }
    }
    if (comp) return 0;
  }
  return 1;
}
static void tf_init(orc_ctx *c) {
  u64 p = (1ull << 62) + 1;
  do { p -= 1ull << 14; } while (!is_prime_u64(p));
  c->P = p;
  u64 inv = 1; /* p^-1 mod 2^64 by Newton */
  for (int i = 0; i < 6; i++) inv *= 2 - p * inv;
  c->NPINV = (u64)0 - inv;
  u64 R = (u64)(((u128)1 << 64) % p);
  c->R2 = mulmod(R, R, p);
  int n = c->n;
  u64 psi = 0;
  for (u64 g = 2;; g++) {
    psi = powmod(g, (p - 1) / (2 * (u64)n), p);
    if (powmod(psi, (u64)n, p) == p - 1) break;
  }
  u64 psi_inv = powmod(psi, p - 2, p);
  c->psi_rev_m = (u64 *)malloc(sizeof(u64) * n);
  c->psi_inv_rev_m = (u64 *)malloc(sizeof(u64) * n);
  for (int i = 0; i < n; i++) {
    u64 r = orc_reverse_bits_msb((u64)i, (uint32_t)c->log_n);
    c->psi_rev_m[i] = mulmod(powmod(psi, r, p), R, p);
    c->psi_inv_rev_m[i] = mulmod(powmod(psi_inv, r, p), R, p);
  }
  c->NINV_M = mulmod(powmod((u64)n, p - 2, p), R, p);
}
static void tf_free(orc_ctx *c) { free(c->psi_rev_m); free(c->psi_inv_rev_m); }

static void ntt_inplace(const orc_ctx *c, u64 *a) {
  const int n = c->n;
  const u64 P = c->P;
  int t = n;
  for (int m = 1; m < n; m <<= 1) {
    t >>= 1;
    for (int i = 0; i < m; i++) {
      u64 S = c->psi_rev_m[m + i];
      u64 *x = a + 2 * i * t, *y = x + t;
      for (int j = 0; j < t; j++) {
        u64 U = x[j], V = mont_mul(c, y[j], S);
        u64 s = U + V;
        x[j] = s >= P ? s - P : s;
        y[j] = U >= V ? U - V : U + P - V;
      }
    }
  }
}
/* mont != 0: output scaled by R (prepared-matrix form, so mont_mul(a, b_m) = a*b) */
static void tf_forward(const orc_ctx *c, const i64 *a, tfe *out, int mont) {
  const u64 P = c->P;
  for (int i = 0; i < c->n; i++) {
    i64 v = a[i];
    u64 x = v >= 0 ? (u64)v % P : P - ((u64)(-v) % P);
    if (x == P) x = 0;
    out[i] = mont ? mont_mul(c, x, c->R2) : x;
  }
  ntt_inplace(c, out);
}
static void tf_zero(const orc_ctx *c, tfe *acc) { memset(acc, 0, sizeof(tfe) * c->n); }
static void tf_mac(const orc_ctx *c, tfe *acc, const tfe *a, const tfe *bm) {
  const u64 P = c->P;
  for (int i = 0; i < c->n; i++) {
    u64 s = acc[i] + mont_mul(c, a[i], bm[i]);
    acc[i] = s >= P ? s - P : s;
  }
}
static void tf_inverse(const orc_ctx *c, tfe *a, i64 *out) {
  const int n = c->n;
  const u64 P = c->P;
  int t = 1;
  for (int m = n; m > 1; m >>= 1) {
    int h = m >> 1;
    for (int i = 0; i < h; i++) {
      u64 S = c->psi_inv_rev_m[h + i];
      u64 *x = a + 2 * i * t, *y = x + t;
      for (int j = 0; j < t; j++) {
        u64 U = x[j], V = y[j];
        u64 s = U + V;
        x[j] = s >= P ? s - P : s;
        u64 d = U >= V ? U - V : U + P - V;
        y[j] = mont_mul(c, d, S);
      }
    }
    t <<= 1;
  }
  for (int i = 0; i < n; i++) {
    u64 v = mont_mul(c, a[i], c->NINV_M);
    out[i] = v > P / 2 ? -(i64)(P - v) : (i64)v;
  }
}
#else /* ---------------------------------- fft64 ---------------------------------- */
static void tf_init(orc_ctx *c) {
  int m = c->n / 2;
  c->tw_re = (double *)malloc(sizeof(double) * m);
  c->tw_im = (double *)malloc(sizeof(double) * m);
  c->tw_re[0] = c->tw_im[0] = 0;
  for (int s = 0; (1 << s) < m; s++)
    for (int b = 0; b < (1 << s); b++) {
      /* shift theta = (1/4 + bitrev_s(b)) / 2^s ; zeta = e^{i pi theta} */
      u64 rb = orc_reverse_bits_msb((u64)b, (uint32_t)s);
      long double th = (0.25L + (long double)rb) / (long double)(1ull << s);
      c->tw_re[(1 << s) + b] = (double)cosl(3.14159265358979323846264338327950288L * th);
      c->tw_im[(1 << s) + b] = (double)sinl(3.14159265358979323846264338327950288L * th);
    }
}
static void tf_free(orc_ctx *c) { free(c->tw_re); free(c->tw_im); }
/* layout: out[0..m) real parts, out[m..2m) imaginary parts, frequency order bit-reversed
 * (private to the backend).
 * Written so that gcc vectorises it (AVX2 / AVX-512 clones picked at load time): stages are fused in pairs
 * (radix 4: half the passes over the 32 KiB of data), the last two stages (distance 2 and 1) work on blocks
 * of four, and the int64 <-> double conversions use the 2^52 + 2^51 trick instead of cvtsi2sd / llrint
 * (exact for |x| < 2^51: inputs are limbs, outputs are < 2^47; same round-to-nearest-even). */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(ORC_NO_CLONES)
#define ORC_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define ORC_CLONES
#endif
#ifdef __AVX2__
#include <immintrin.h>
#endif
#define ORC_MAGIC 6755399441055744.0 /* 2^52 + 2^51 */
static inline double i64_to_f64(i64 v) {
  union { double d; i64 i; } u;
  u.i = v + 0x4338000000000000LL;
  return u.d - ORC_MAGIC;
}
static inline i64 f64_round_i64(double v) {
  union { double d; i64 i; } u;
  u.d = v + ORC_MAGIC;
  return u.i - 0x4338000000000000LL;
}
/* forward butterfly pair (x, y) <- (x + w y, x - w y) on arrays of length t */
static inline void fwd_bf(double *xr, double *xi, double *yr, double *yi, double wr, double wi, int t) {
  for (int j = 0; j < t; j++) {
    double vr = yr[j] * wr - yi[j] * wi, vi = yr[j] * wi + yi[j] * wr;
    double ur = xr[j], ui = xi[j];
    xr[j] = ur + vr; xi[j] = ui + vi;
    yr[j] = ur - vr; yi[j] = ui - vi;
  }
}
ORC_CLONES
static void tf_forward(const orc_ctx *c, const i64 *a, tfe *out, int mont) {
  (void)mont;
  const int m = c->n / 2;
  double *re = out, *im = out + m;
  int t = m, s = 1;
  if ((t >> 2) >= 4) {
    /* first pair of stages straight from the integer limbs: no separate conversion pass over the data (that pass
     * alone cost more than a radix-4 pass: loads of a[] right behind stores to out[] at the same page offsets) */
    const int q = t >> 2;
    const double w1r = c->tw_re[1], w1i = c->tw_im[1];
    const double w2r = c->tw_re[2], w2i = c->tw_im[2], w3r = c->tw_re[3], w3i = c->tw_im[3];
    const i64 *restrict ar = a, *restrict ai = a + m;
    double *restrict r0 = re, *restrict i0 = im;
    double *restrict r1 = r0 + q, *restrict i1 = i0 + q, *restrict r2 = r0 + 2 * q, *restrict i2 = i0 + 2 * q,
           *restrict r3 = r0 + 3 * q, *restrict i3 = i0 + 3 * q;
#pragma GCC ivdep
    for (int j = 0; j < q; j++) {
      const double y0r = i64_to_f64(ar[j]), y0i = i64_to_f64(ai[j]);
      const double y1r = i64_to_f64(ar[j + q]), y1i = i64_to_f64(ai[j + q]);
      const double y2r = i64_to_f64(ar[j + 2 * q]), y2i = i64_to_f64(ai[j + 2 * q]);
      const double y3r = i64_to_f64(ar[j + 3 * q]), y3i = i64_to_f64(ai[j + 3 * q]);
      double pr = y2r * w1r - y2i * w1i, pi = y2r * w1i + y2i * w1r;
      double br = y3r * w1r - y3i * w1i, bi = y3r * w1i + y3i * w1r;
      double x0r = y0r + pr, x0i = y0i + pi, x2r = y0r - pr, x2i = y0i - pi;
      double x1r = y1r + br, x1i = y1i + bi, x3r = y1r - br, x3i = y1i - bi;
      double cr = x1r * w2r - x1i * w2i, ci = x1r * w2i + x1i * w2r;
      double dr = x3r * w3r - x3i * w3i, di = x3r * w3i + x3i * w3r;
      r0[j] = x0r + cr; i0[j] = x0i + ci; r1[j] = x0r - cr; i1[j] = x0i - ci;
      r2[j] = x2r + dr; i2[j] = x2i + di; r3[j] = x2r - dr; i3[j] = x2i - di;
    }
    t = q; s = 4;
  } else {
    for (int i = 0; i < m; i++) { re[i] = i64_to_f64(a[i]); im[i] = i64_to_f64(a[i + m]); }
  }
  /* fused pairs of stages (s, 2s) while the second stage still has distance >= 4 */
  for (; (t >> 2) >= 4; s <<= 2) {
    const int q = t >> 2; /* distance of the second stage; the block of 4q elements splits into quarters */
    for (int b = 0; b < s; b++) {
      const double w1r = c->tw_re[s + b], w1i = c->tw_im[s + b];
      const double w2r = c->tw_re[2 * s + 2 * b], w2i = c->tw_im[2 * s + 2 * b];
      const double w3r = c->tw_re[2 * s + 2 * b + 1], w3i = c->tw_im[2 * s + 2 * b + 1];
      /* the quarters do not overlap: without restrict + ivdep gcc leaves this loop scalar */
      double *restrict r0 = re + (size_t)b * t, *restrict i0 = im + (size_t)b * t;
      double *restrict r1 = r0 + q, *restrict i1 = i0 + q, *restrict r2 = r0 + 2 * q, *restrict i2 = i0 + 2 * q,
             *restrict r3 = r0 + 3 * q, *restrict i3 = i0 + 3 * q;
#pragma GCC ivdep
      for (int j = 0; j < q; j++) {
        /* stage s: (0,2) and (1,3) with w1 */
        double ar = r2[j] * w1r - i2[j] * w1i, ai = r2[j] * w1i + i2[j] * w1r;
        double br = r3[j] * w1r - i3[j] * w1i, bi = r3[j] * w1i + i3[j] * w1r;
        double x0r = r0[j] + ar, x0i = i0[j] + ai, x2r = r0[j] - ar, x2i = i0[j] - ai;
        double x1r = r1[j] + br, x1i = i1[j] + bi, x3r = r1[j] - br, x3i = i1[j] - bi;
        /* stage 2s: (0,1) with w2, (2,3) with w3 */
        double cr = x1r * w2r - x1i * w2i, ci = x1r * w2i + x1i * w2r;
        double dr = x3r * w3r - x3i * w3i, di = x3r * w3i + x3i * w3r;
        r0[j] = x0r + cr; i0[j] = x0i + ci; r1[j] = x0r - cr; i1[j] = x0i - ci;
        r2[j] = x2r + dr; i2[j] = x2i + di; r3[j] = x2r - dr; i3[j] = x2i - di;
      }
    }
    t = q;
  }
#ifdef __AVX2__
  if (t == 8 && s * 8 == m) {
    /* last three stages (distances 4, 2, 1) in registers, one block of 8 points per iteration */
    const __m256d sg2 = _mm256_set_pd(-1.0, -1.0, 1.0, 1.0), sg1 = _mm256_set_pd(-1.0, 1.0, -1.0, 1.0);
    const double *w4r = c->tw_re + s, *w4i = c->tw_im + s;          /* stage s: one twiddle per block */
    const double *w2r = c->tw_re + 2 * s, *w2i = c->tw_im + 2 * s;  /* stage 2s: one per half block */
    const double *w1r = c->tw_re + 4 * s, *w1i = c->tw_im + 4 * s;  /* stage 4s: one per pair */
    for (int b = 0; b < s; b++) {
      __m256d r0 = _mm256_loadu_pd(re + 8 * b), r1 = _mm256_loadu_pd(re + 8 * b + 4);
      __m256d i0 = _mm256_loadu_pd(im + 8 * b), i1 = _mm256_loadu_pd(im + 8 * b + 4);
      { /* distance 4 */
        const __m256d wr = _mm256_broadcast_sd(w4r + b), wi = _mm256_broadcast_sd(w4i + b);
        const __m256d vr = _mm256_fnmadd_pd(i1, wi, _mm256_mul_pd(r1, wr));
        const __m256d vi = _mm256_fmadd_pd(i1, wr, _mm256_mul_pd(r1, wi));
        r1 = _mm256_sub_pd(r0, vr); r0 = _mm256_add_pd(r0, vr);
        i1 = _mm256_sub_pd(i0, vi); i0 = _mm256_add_pd(i0, vi);
      }
#define ORC_FWD_D2(R, I, K)                                                                                     \
      {                                                                                                         \
        const __m256d wr = _mm256_broadcast_sd(w2r + 2 * b + (K)), wi = _mm256_broadcast_sd(w2i + 2 * b + (K)); \
        const __m256d ur = _mm256_permute2f128_pd(R, R, 0x00), yr = _mm256_permute2f128_pd(R, R, 0x11);         \
        const __m256d ui = _mm256_permute2f128_pd(I, I, 0x00), yi = _mm256_permute2f128_pd(I, I, 0x11);         \
        const __m256d vr = _mm256_fnmadd_pd(yi, wi, _mm256_mul_pd(yr, wr));                                     \
        const __m256d vi = _mm256_fmadd_pd(yi, wr, _mm256_mul_pd(yr, wi));                                      \
        R = _mm256_fmadd_pd(vr, sg2, ur); I = _mm256_fmadd_pd(vi, sg2, ui);                                     \
      }
      ORC_FWD_D2(r0, i0, 0)
      ORC_FWD_D2(r1, i1, 1)
#define ORC_FWD_D1(R, I, K)                                                                                          \
      {                                                                                                              \
        const __m256d wr = _mm256_permute4x64_pd(_mm256_castpd128_pd256(_mm_loadu_pd(w1r + 4 * b + 2 * (K))), 0x50); \
        const __m256d wi = _mm256_permute4x64_pd(_mm256_castpd128_pd256(_mm_loadu_pd(w1i + 4 * b + 2 * (K))), 0x50); \
        const __m256d ur = _mm256_permute_pd(R, 0x0), yr = _mm256_permute_pd(R, 0xF);                                \
        const __m256d ui = _mm256_permute_pd(I, 0x0), yi = _mm256_permute_pd(I, 0xF);                                \
        const __m256d vr = _mm256_fnmadd_pd(yi, wi, _mm256_mul_pd(yr, wr));                                          \
        const __m256d vi = _mm256_fmadd_pd(yi, wr, _mm256_mul_pd(yr, wi));                                           \
        R = _mm256_fmadd_pd(vr, sg1, ur); I = _mm256_fmadd_pd(vi, sg1, ui);                                          \
      }
      ORC_FWD_D1(r0, i0, 0)
      ORC_FWD_D1(r1, i1, 1)
      _mm256_storeu_pd(re + 8 * b, r0); _mm256_storeu_pd(re + 8 * b + 4, r1);
      _mm256_storeu_pd(im + 8 * b, i0); _mm256_storeu_pd(im + 8 * b + 4, i1);
    }
    return;
  }
#endif
  /* remaining single stages */
  for (; s < m; s <<= 1) {
    t >>= 1;
    if (t >= 4) {
      for (int b = 0; b < s; b++)
        fwd_bf(re + 2 * b * t, im + 2 * b * t, re + 2 * b * t + t, im + 2 * b * t + t, c->tw_re[s + b], c->tw_im[s + b], t);
    } else {
      for (int b = 0; b < s; b++) {
        const double wr = c->tw_re[s + b], wi = c->tw_im[s + b];
        double *xr = re + 2 * b * t, *xi = im + 2 * b * t;
        for (int j = 0; j < t; j++) {
          double vr = xr[t + j] * wr - xi[t + j] * wi, vi = xr[t + j] * wi + xi[t + j] * wr;
          double ur = xr[j], ui = xi[j];
          xr[j] = ur + vr; xi[j] = ui + vi;
          xr[t + j] = ur - vr; xi[t + j] = ui - vi;
        }
      }
    }
  }
}
static void tf_zero(const orc_ctx *c, tfe *acc) { memset(acc, 0, sizeof(tfe) * c->n); }
ORC_CLONES
static void tf_mac(const orc_ctx *c, tfe *acc, const tfe *a, const tfe *b) {
  const int m = c->n / 2;
  for (int i = 0; i < m; i++) {
    double ar = a[i], ai = a[i + m], br = b[i], bi = b[i + m];
    acc[i] += ar * br - ai * bi;
    acc[i + m] += ar * bi + ai * br;
  }
}
/* acc = sum_k a[k] * b[k] (complex, pointwise), terms added in index order starting from zero: one pass over the
 * accumulator instead of tf_zero + cnt x tf_mac */
static void tf_dot(const orc_ctx *c, tfe *restrict acc, const tfe *const *a, const tfe *const *b, int cnt) {
  const int m = c->n / 2;
#ifdef __AVX2__
  for (int i = 0; i < m; i += 4) {
    __m256d sr = _mm256_setzero_pd(), si = _mm256_setzero_pd();
    for (int k = 0; k < cnt; k++) {
      const __m256d ar = _mm256_loadu_pd(a[k] + i), ai = _mm256_loadu_pd(a[k] + i + m);
      const __m256d br = _mm256_loadu_pd(b[k] + i), bi = _mm256_loadu_pd(b[k] + i + m);
      sr = _mm256_add_pd(sr, _mm256_fnmadd_pd(ai, bi, _mm256_mul_pd(ar, br)));
      si = _mm256_add_pd(si, _mm256_fmadd_pd(ai, br, _mm256_mul_pd(ar, bi)));
    }
    _mm256_storeu_pd(acc + i, sr);
    _mm256_storeu_pd(acc + i + m, si);
  }
#else
  tf_zero(c, acc);
  for (int k = 0; k < cnt; k++) tf_mac(c, acc, a[k], b[k]);
#endif
}
ORC_CLONES
static void tf_inverse(const orc_ctx *c, tfe *a, i64 *out) {
  const int m = c->n / 2;
  double *re = a, *im = a + m;
  int t = 1, s = m >> 1;
#ifdef __AVX2__
  if (m >= 16) {
    /* first three stages (distances 1, 2, 4) in registers, one block of 8 points per iteration */
    const double *w1r = c->tw_re + s, *w1i = c->tw_im + s;          /* one twiddle per pair */
    const double *w2r = c->tw_re + s / 2, *w2i = c->tw_im + s / 2;  /* one per block of four */
    const double *w4r = c->tw_re + s / 4, *w4i = c->tw_im + s / 4;  /* one per block of eight */
#define ORC_INV_D1(R, I, K)                                                                                          \
      { /* distance 1: (x, y) <- (x + y, (x - y) conj(w)) */                                                         \
        const __m256d wr = _mm256_permute4x64_pd(_mm256_castpd128_pd256(_mm_loadu_pd(w1r + 4 * b + 2 * (K))), 0x50); \
        const __m256d wi = _mm256_permute4x64_pd(_mm256_castpd128_pd256(_mm_loadu_pd(w1i + 4 * b + 2 * (K))), 0x50); \
        const __m256d ur = _mm256_permute_pd(R, 0x0), yr = _mm256_permute_pd(R, 0xF);                                \
        const __m256d ui = _mm256_permute_pd(I, 0x0), yi = _mm256_permute_pd(I, 0xF);                                \
        const __m256d sr = _mm256_add_pd(ur, yr), si = _mm256_add_pd(ui, yi);                                        \
        const __m256d dr = _mm256_sub_pd(ur, yr), di = _mm256_sub_pd(ui, yi);                                        \
        const __m256d pr = _mm256_fmadd_pd(di, wi, _mm256_mul_pd(dr, wr));  /* dr wr + di wi */                      \
        const __m256d pi = _mm256_fnmadd_pd(dr, wi, _mm256_mul_pd(di, wr)); /* di wr - dr wi */                      \
        R = _mm256_blend_pd(sr, pr, 0xA); I = _mm256_blend_pd(si, pi, 0xA);                                          \
      }
#define ORC_INV_D2(R, I, K)                                                                                          \
      {                                                                                                              \
        const __m256d wr = _mm256_broadcast_sd(w2r + 2 * b + (K)), wi = _mm256_broadcast_sd(w2i + 2 * b + (K));      \
        const __m256d ur = _mm256_permute2f128_pd(R, R, 0x00), yr = _mm256_permute2f128_pd(R, R, 0x11);              \
        const __m256d ui = _mm256_permute2f128_pd(I, I, 0x00), yi = _mm256_permute2f128_pd(I, I, 0x11);              \
        const __m256d sr = _mm256_add_pd(ur, yr), si = _mm256_add_pd(ui, yi);                                        \
        const __m256d dr = _mm256_sub_pd(ur, yr), di = _mm256_sub_pd(ui, yi);                                        \
        const __m256d pr = _mm256_fmadd_pd(di, wi, _mm256_mul_pd(dr, wr));                                           \
        const __m256d pi = _mm256_fnmadd_pd(dr, wi, _mm256_mul_pd(di, wr));                                          \
        R = _mm256_blend_pd(sr, pr, 0xC); I = _mm256_blend_pd(si, pi, 0xC);                                          \
      }
    for (int b = 0; b < m / 8; b++) {
      __m256d R0 = _mm256_loadu_pd(re + 8 * b), R1 = _mm256_loadu_pd(re + 8 * b + 4);
      __m256d I0 = _mm256_loadu_pd(im + 8 * b), I1 = _mm256_loadu_pd(im + 8 * b + 4);
      ORC_INV_D1(R0, I0, 0)
      ORC_INV_D1(R1, I1, 1)
      ORC_INV_D2(R0, I0, 0)
      ORC_INV_D2(R1, I1, 1)
      { /* distance 4 */
        const __m256d wr = _mm256_broadcast_sd(w4r + b), wi = _mm256_broadcast_sd(w4i + b);
        const __m256d dr = _mm256_sub_pd(R0, R1), di = _mm256_sub_pd(I0, I1);
        R0 = _mm256_add_pd(R0, R1); I0 = _mm256_add_pd(I0, I1);
        R1 = _mm256_fmadd_pd(di, wi, _mm256_mul_pd(dr, wr));
        I1 = _mm256_fnmadd_pd(dr, wi, _mm256_mul_pd(di, wr));
      }
      _mm256_storeu_pd(re + 8 * b, R0); _mm256_storeu_pd(re + 8 * b + 4, R1);
      _mm256_storeu_pd(im + 8 * b, I0); _mm256_storeu_pd(im + 8 * b + 4, I1);
    }
    t = 8; s >>= 3;
  }
#endif
  /* single stages while the distance is below 4 */
  for (; s >= 1 && t < 4; s >>= 1) {
    for (int b = 0; b < s; b++) {
      const double wr = c->tw_re[s + b], wi = -c->tw_im[s + b];
      double *xr = re + 2 * b * t, *xi = im + 2 * b * t;
      for (int j = 0; j < t; j++) {
        double ur = xr[j], ui = xi[j], vr = xr[t + j], vi = xi[t + j];
        xr[j] = ur + vr; xi[j] = ui + vi;
        double dr = ur - vr, di = ui - vi;
        xr[t + j] = dr * wr - di * wi; xi[t + j] = dr * wi + di * wr;
      }
    }
    t <<= 1;
  }
  const double sc = 1.0 / (double)m;
  /* fused pairs of stages (s, s / 2): distance t then 2t */
  for (; s >= 2; s >>= 2) {
    const int h = s >> 1; /* blocks of the second stage */
    if (s == 2) {
      /* the last pair: scaled and rounded straight into the integer result (no separate pass) */
      const double w2r = c->tw_re[2], w2i = -c->tw_im[2], w3r = c->tw_re[3], w3i = -c->tw_im[3];
      const double w1r = c->tw_re[1], w1i = -c->tw_im[1];
      const double *restrict r0 = re, *restrict i0 = im;
      const double *restrict r1 = r0 + t, *restrict i1 = i0 + t, *restrict r2 = r0 + 2 * t, *restrict i2 = i0 + 2 * t,
                   *restrict r3 = r0 + 3 * t, *restrict i3 = i0 + 3 * t;
      i64 *restrict o = out;
#pragma GCC ivdep
      for (int j = 0; j < t; j++) {
        double x0r = r0[j] + r1[j], x0i = i0[j] + i1[j], d0r = r0[j] - r1[j], d0i = i0[j] - i1[j];
        double x2r = r2[j] + r3[j], x2i = i2[j] + i3[j], d1r = r2[j] - r3[j], d1i = i2[j] - i3[j];
        double x1r = d0r * w2r - d0i * w2i, x1i = d0r * w2i + d0i * w2r;
        double x3r = d1r * w3r - d1i * w3i, x3i = d1r * w3i + d1i * w3r;
        double e0r = x0r - x2r, e0i = x0i - x2i, e1r = x1r - x3r, e1i = x1i - x3i;
        o[j] = f64_round_i64((x0r + x2r) * sc); o[j + m] = f64_round_i64((x0i + x2i) * sc);
        o[j + t] = f64_round_i64((x1r + x3r) * sc); o[j + t + m] = f64_round_i64((x1i + x3i) * sc);
        o[j + 2 * t] = f64_round_i64((e0r * w1r - e0i * w1i) * sc); o[j + 2 * t + m] = f64_round_i64((e0r * w1i + e0i * w1r) * sc);
        o[j + 3 * t] = f64_round_i64((e1r * w1r - e1i * w1i) * sc); o[j + 3 * t + m] = f64_round_i64((e1r * w1i + e1i * w1r) * sc);
      }
      return;
    }
    for (int b = 0; b < h; b++) {
      const double w2r = c->tw_re[s + 2 * b], w2i = -c->tw_im[s + 2 * b];
      const double w3r = c->tw_re[s + 2 * b + 1], w3i = -c->tw_im[s + 2 * b + 1];
      const double w1r = c->tw_re[h + b], w1i = -c->tw_im[h + b];
      double *restrict r0 = re + (size_t)b * 4 * t, *restrict i0 = im + (size_t)b * 4 * t;
      double *restrict r1 = r0 + t, *restrict i1 = i0 + t, *restrict r2 = r0 + 2 * t, *restrict i2 = i0 + 2 * t,
             *restrict r3 = r0 + 3 * t, *restrict i3 = i0 + 3 * t;
#pragma GCC ivdep
      for (int j = 0; j < t; j++) {
        /* stage s: (0,1) with w2, (2,3) with w3 */
        double x0r = r0[j] + r1[j], x0i = i0[j] + i1[j], d0r = r0[j] - r1[j], d0i = i0[j] - i1[j];
        double x2r = r2[j] + r3[j], x2i = i2[j] + i3[j], d1r = r2[j] - r3[j], d1i = i2[j] - i3[j];
        double x1r = d0r * w2r - d0i * w2i, x1i = d0r * w2i + d0i * w2r;
        double x3r = d1r * w3r - d1i * w3i, x3i = d1r * w3i + d1i * w3r;
        /* stage s / 2: (0,2) and (1,3) with w1 */
        r0[j] = x0r + x2r; i0[j] = x0i + x2i;
        r1[j] = x1r + x3r; i1[j] = x1i + x3i;
        double e0r = x0r - x2r, e0i = x0i - x2i, e1r = x1r - x3r, e1i = x1i - x3i;
        r2[j] = e0r * w1r - e0i * w1i; i2[j] = e0r * w1i + e0i * w1r;
        r3[j] = e1r * w1r - e1i * w1i; i3[j] = e1r * w1i + e1i * w1r;
      }
    }
    t <<= 2;
  }
  if (s == 1) { /* one stage left: fused with the scaling and the rounding to integers */
    const double wr = c->tw_re[1], wi = -c->tw_im[1];
    const double *restrict xr = re, *restrict xi = im, *restrict yr = re + t, *restrict yi = im + t;
    i64 *restrict o = out;
#pragma GCC ivdep
    for (int j = 0; j < t; j++) {
      const double ur = xr[j], ui = xi[j], vr = yr[j], vi = yi[j];
      const double dr = ur - vr, di = ui - vi;
      o[j] = f64_round_i64((ur + vr) * sc);
      o[j + m] = f64_round_i64((ui + vi) * sc);
      o[j + t] = f64_round_i64((dr * wr - di * wi) * sc);
      o[j + t + m] = f64_round_i64((dr * wi + di * wr) * sc);
    }
    return;
  }
  for (int i = 0; i < m; i++) {
    out[i] = f64_round_i64(re[i] * sc);
    out[i + m] = f64_round_i64(im[i] * sc);
  }
}
#endif

/* ======================================================================================
 * Context
 * ==================================================================================== */
static int ceil_div(int a, int b) { return (a + b - 1) / b; }

static i64 mod_pow_i64(i64 b, u64 e, i64 m) {
  i64 r = 1;
  b %= m;
  while (e) {
    if (e & 1) r = (i64)((u128)r * b % m);
    b = (i64)((u128)b * b % m);
    e >>= 1;
  }
  return r;
}

orc_ctx *orc_ctx_new(const orc_params *p) {
#if defined(__GLIBC__) && !defined(ORC_NO_MALLOPT)
  /* the per-operation temporaries (192 - 256 KiB) are above glibc's default mmap threshold: every operation would map,
   * fault in and unmap its buffers (a sixth of an external product).  Keep them on the heap. */
  mallopt(M_MMAP_THRESHOLD, 64 << 20);
  mallopt(M_TRIM_THRESHOLD, 256 << 20);
#endif
  orc_ctx *c = (orc_ctx *)calloc(1, sizeof(*c));
  c->p = *p;
  c->log_n = p->log_n;
  c->n = 1 << p->log_n;
  c->k = p->base2k;
  c->size_ct = ceil_div(p->k_ct, p->base2k);
  c->dnum_ct = ceil_div(p->k_ct, p->base2k);           /* parameters.rs:138-140 */
  c->size_addr = ceil_div(p->k_addr, p->base2k);
  c->size_evk_trace = ceil_div(p->k_evk_trace, p->base2k);
  c->dnum_ggsw = ceil_div(p->k_addr, p->base2k);       /* parameters.rs:142-144 */
  c->size_evk_inv = ceil_div(p->k_evk_ggsw_inv, p->base2k);
  int32_t lens[8], digits[64];
  c->n_coord = orc_get_base_2d((uint32_t)p->max_addr, p->decomp_n, p->n_decomp, lens, digits);
  c->n_ggsw = 0;
  for (int i = 0; i < c->n_coord; i++) {
    c->coord_len[i] = lens[i];
    for (int j = 0; j < lens[i]; j++) c->coord_digits[i][j] = digits[i * 8 + j];
    c->n_ggsw += lens[i];
  }
  c->n_glwe = (int)((p->max_addr + (u64)c->n - 1) / (u64)c->n);
  /* Poulpy [spec] GLWE::trace_galois_elements: [-1, 5^(2^0), 5^(2^1), ...] mod 2n */
  for (int i = 0; i < c->log_n; i++)
    c->gal[i] = i == 0 ? -1 : mod_pow_i64(5, 1ull << (i - 1), 2 * (i64)c->n);
  tf_init(c);
  return c;
}
void orc_ctx_free(orc_ctx *c) {
  if (!c) return;
  tf_free(c);
  free(c);
}
size_t orc_n(const orc_ctx *c) { return (size_t)c->n; }
size_t orc_glwe_len(const orc_ctx *c) { return (size_t)2 * c->size_ct * c->n; }
size_t orc_ggsw_len(const orc_ctx *c) { return (size_t)c->dnum_ct * 2 * 2 * c->size_addr * c->n; }
size_t orc_atk_len(const orc_ctx *c) { return (size_t)c->dnum_ct * 2 * c->size_evk_trace * c->n; }
size_t orc_evk_inv_len(const orc_ctx *c) { return (size_t)c->dnum_ggsw * 2 * c->size_evk_inv * c->n; }
int orc_n_gal(const orc_ctx *c) { return c->log_n; }
int orc_n_ggsw(const orc_ctx *c) { return c->n_ggsw; }
int orc_n_glwe_per_subram(const orc_ctx *c) { return c->n_glwe; }
int64_t orc_gal_el(const orc_ctx *c, int i) { return c->gal[i]; }
void orc_op_counters(const orc_ctx *c, uint64_t out[2]) { out[0] = c->counters[0]; out[1] = c->counters[1]; }
static void count_op(const orc_ctx *c, int which) {
  __atomic_fetch_add(&((orc_ctx *)c)->counters[which], 1, __ATOMIC_RELAXED);
}

/* ======================================================================================
 * VecZnx helpers.  v = base pointer, cols, n; at(col, limb) = v + ((limb*cols)+col)*n
 * ==================================================================================== */
#define AT(v, cols, n, col, limb) ((v) + ((size_t)(limb) * (cols) + (col)) * (size_t)(n))

static inline i64 get_digit(int k, i64 x) { return (i64)((u64)x << (64 - k)) >> (64 - k); }
static inline i64 get_carry(int k, i64 x, i64 d) { return (x - d) >> k; }
/* the same two functions with logical shifts only, so that gcc vectorises the loops below for AVX2 (no 64-bit
 * arithmetic right shift there): balanced digit = ((x + 2^(k-1)) mod 2^k) - 2^(k-1); x - d is a multiple of 2^k and
 * |x - d| < 2^62 for every value on the path (limbs and their sums stay below 2^52) */
static inline i64 digit_l(int k, i64 x) {
  const u64 h = (u64)1 << (k - 1), mask = ((u64)1 << k) - 1;
  return (i64)((((u64)x + h) & mask) - h);
}
static inline i64 carry_l(int k, i64 x, i64 d) {
  return (i64)((((u64)(x - d) + ((u64)1 << 62)) >> k) - ((u64)1 << (62 - k)));
}

/* Poulpy [spec] vec_znx_normalize / vec_znx_big_normalize, same base2k on both sides.
 * Walks limbs from least to most significant; limbs of `a` beyond res_size contribute their
 * carry only; the carry out of limb 0 is dropped (torus wrap).  Poulpy's two-stage step
 * (digit(x), carry(x), then digit(digit+c), carry += ...) equals the one-stage form used
 * here exactly: digit(digit(x)+c) = digit(x+c) and the carries sum to (x+c-digit)>>k. */
static void vz_normalize(int n, int k, i64 *res, int rcols, int rcol, int rsize, const i64 *a,
                         int acols, int acol, int asize) {
  /* limb by limb over all coefficients (the carries of one limb in an array): same integers as the
   * coefficient-by-coefficient walk, written so that the inner loops vectorise */
  i64 *restrict cy = (i64 *)malloc(sizeof(i64) * (size_t)n);
  memset(cy, 0, sizeof(i64) * (size_t)n);
  for (int j = asize - 1; j >= 0; j--) {
    const i64 *restrict aj = AT(a, acols, n, acol, j);
    if (j < rsize) {
      i64 *restrict rj = AT(res, rcols, n, rcol, j);
      if (rj == aj) {  /* in place */
        for (int i = 0; i < n; i++) {
          const i64 t = rj[i] + cy[i], d = digit_l(k, t);
          cy[i] = carry_l(k, t, d);
          rj[i] = d;
        }
      } else {
        for (int i = 0; i < n; i++) {
          const i64 t = aj[i] + cy[i], d = digit_l(k, t);
          cy[i] = carry_l(k, t, d);
          rj[i] = d;
        }
      }
    } else {
      for (int i = 0; i < n; i++) {
        const i64 t = aj[i] + cy[i];
        cy[i] = carry_l(k, t, digit_l(k, t));
      }
    }
  }
  free(cy);
  for (int j = asize; j < rsize; j++) memset(AT(res, rcols, n, rcol, j), 0, sizeof(i64) * n);
}

/* Poulpy [spec] vec_znx_rsh_inplace(base2k, k): torus value / 2^k.  For k not a multiple of
 * base2k it shifts by one extra limb and left-shifts by lsh = base2k - k%base2k inside the
 * normalisation steps; net effect per coefficient: the integer X = sum limb_j 2^(K(size-1-j))
 * becomes ceil(X / 2^k) written in balanced digits (top carry dropped). */
static void vz_rsh_inplace(int n, int K, int k, i64 *v, int cols, int col, int size) {
  if (k == 0) return;
  int steps = k / K, k_rem = k % K;
  if (steps >= size) {
    for (int j = 0; j < size; j++) memset(AT(v, cols, n, col, j), 0, sizeof(i64) * n);
    return;
  }
  if (k_rem == 0) {
    for (int j = size - 1; j >= steps; j--)
      memcpy(AT(v, cols, n, col, j), AT(v, cols, n, col, j - steps), sizeof(i64) * n);
    for (int j = 0; j < steps; j++) memset(AT(v, cols, n, col, j), 0, sizeof(i64) * n);
    return;
  }
  steps += 1;
  const int lsh = K - k_rem, bl = K - lsh;
  /* the same walk as Poulpy's (per coefficient: least significant limb first), limb by limb over all coefficients
   * with the carries in an array so that the inner loops vectorise.  The limbs are read before they are
   * overwritten: limb j is written from limb j - steps, walking j downwards. */
  i64 *restrict cy = (i64 *)malloc(sizeof(i64) * (size_t)n);
  /* limbs shifted out: carry only */
  for (int j = size - 1; j >= size - steps; j--) {
    const i64 *restrict vj = AT(v, cols, n, col, j);
    if (j == size - 1) {
      for (int i = 0; i < n; i++) {
        const i64 x = vj[i];
        cy[i] = carry_l(bl, x, digit_l(bl, x));
      }
    } else {
      for (int i = 0; i < n; i++) {
        const i64 x = vj[i], d = digit_l(bl, x), c0 = carry_l(bl, x, d);
        const i64 dpc = (i64)((u64)d << lsh) + cy[i];
        cy[i] = c0 + carry_l(K, dpc, digit_l(K, dpc));
      }
    }
  }
  /* shifted normalisation: res limb j from source limb j - steps */
  for (int j = size - 1; j >= steps; j--) {
    const i64 *restrict src = AT(v, cols, n, col, j - steps);
    i64 *restrict dst = AT(v, cols, n, col, j);
    for (int i = 0; i < n; i++) {
      const i64 x = src[i], d = digit_l(bl, x), c0 = carry_l(bl, x, d);
      const i64 dpc = (i64)((u64)d << lsh) + cy[i];
      const i64 r = digit_l(K, dpc);
      dst[i] = r;
      cy[i] = c0 + carry_l(K, dpc, r);
    }
  }
  for (int j = steps - 1; j >= 0; j--) {
    i64 *restrict dst = AT(v, cols, n, col, j);
    for (int i = 0; i < n; i++) {
      const i64 r = digit_l(K, cy[i]);
      dst[i] = r;
      cy[i] = carry_l(K, cy[i], r);
    }
  }
  free(cy);
}

/* Poulpy [spec] vec_znx_rotate: res = a * X^k in Z[X]/(X^n+1) (limb-wise, no renormalise) */
static void poly_rotate(int n, i64 k, const i64 *a, i64 *res) {
  i64 two_n = 2 * (i64)n;
  i64 kk = ((k % two_n) + two_n) % two_n;
  i64 e = kk;  /* (i + kk) mod 2n, 2n a power of two */
  for (int i = 0; i < n; i++) {
    if (e >= n) res[e - n] = -a[i]; else res[e] = a[i];
    e = (e + 1) & (two_n - 1);
  }
}
/* Poulpy [spec] vec_znx_automorphism: X^i -> X^(i*p) */
static void poly_automorphism(int n, i64 p, const i64 *a, i64 *res) {
  i64 two_n = 2 * (i64)n;
  i64 pp = ((p % two_n) + two_n) % two_n;
  i64 e = 0;  /* i * pp mod 2n, 2n a power of two */
  for (int i = 0; i < n; i++) {
    if (e >= n) res[e - n] = -a[i]; else res[e] = a[i];
    e = (e + pp) & (two_n - 1);
  }
}

/* ---- GLWE-level small ops (cols = 2) ---- */
static void glwe_rotate(const orc_ctx *c, i64 k, const i64 *in, i64 *out, int size) {
  for (int l = 0; l < size; l++)
    for (int col = 0; col < 2; col++)
      poly_rotate(c->n, k, AT(in, 2, c->n, col, l), AT(out, 2, c->n, col, l));
}
static void glwe_rotate_inplace(const orc_ctx *c, i64 k, i64 *a, int size) {
  size_t len = (size_t)2 * size * c->n;
  i64 *tmp = (i64 *)malloc(sizeof(i64) * len);
  memcpy(tmp, a, sizeof(i64) * len);
  glwe_rotate(c, k, tmp, a, size);
  free(tmp);
}
static void glwe_small_automorphism(const orc_ctx *c, i64 p, const i64 *in, i64 *out, int size) {
  for (int l = 0; l < size; l++)
    for (int col = 0; col < 2; col++)
      poly_automorphism(c->n, p, AT(in, 2, c->n, col, l), AT(out, 2, c->n, col, l));
}
static void glwe_add_inplace(const orc_ctx *c, i64 *a, const i64 *b, int size) {
  size_t len = (size_t)2 * size * c->n;
  for (size_t i = 0; i < len; i++) a[i] += b[i];
}
static void glwe_sub_inplace(const orc_ctx *c, i64 *a, const i64 *b, int size) { /* a -= b */
  size_t len = (size_t)2 * size * c->n;
  for (size_t i = 0; i < len; i++) a[i] -= b[i];
}
static void glwe_normalize_inplace(const orc_ctx *c, i64 *a, int size) {
  for (int col = 0; col < 2; col++) vz_normalize(c->n, c->k, a, 2, col, size, a, 2, col, size);
}
static void glwe_rsh(const orc_ctx *c, int k, i64 *a, int size) {
  for (int col = 0; col < 2; col++) vz_rsh_inplace(c->n, c->k, k, a, 2, col, size);
}
void orc_glwe_normalize(const orc_ctx *c, int64_t *g) { glwe_normalize_inplace(c, g, c->size_ct); }
void orc_glwe_rsh(const orc_ctx *c, int k, int64_t *g) { glwe_rsh(c, k, g, c->size_ct); }
void orc_glwe_rotate(const orc_ctx *c, int64_t k, const int64_t *in, int64_t *out) {
  glwe_rotate(c, k, in, out, c->size_ct);
}
void orc_glwe_small_automorphism(const orc_ctx *c, int64_t p, const int64_t *in, int64_t *out) {
  glwe_small_automorphism(c, p, in, out, c->size_ct);
}

/* ======================================================================================
 * Prepared matrices (Poulpy [spec] VmpPMat / vmp_prepare) and the vector-matrix product
 * (vmp_apply_dft_to_dft): res[co][l] = sum_{r < min(a_size, rows)} sum_{ci} a[ci][r] * M[r][ci][co][l]
 * ==================================================================================== */
typedef struct {
  int rows, cols_in, cols_out, size;
  tfe *d; /* [row][ci][co][limb][n] */
} pmat;

static pmat *pmat_prepare(const orc_ctx *c, const i64 *raw, int rows, int cols_in, int cols_out,
                          int size) {
  pmat *m = (pmat *)malloc(sizeof(pmat));
  m->rows = rows; m->cols_in = cols_in; m->cols_out = cols_out; m->size = size;
  size_t cnt = (size_t)rows * cols_in * cols_out * size;
  m->d = (tfe *)malloc(sizeof(tfe) * cnt * c->n);
  for (int r = 0; r < rows; r++)
    for (int ci = 0; ci < cols_in; ci++) {
      /* raw MatZnx: at(row, col_in) is a VecZnx(cols_out, size) */
      const i64 *vz = raw + ((size_t)r * cols_in + ci) * (size_t)cols_out * size * c->n;
      for (int co = 0; co < cols_out; co++)
        for (int l = 0; l < size; l++) {
          size_t idx = (((size_t)r * cols_in + ci) * cols_out + co) * size + l;
          tf_forward(c, AT(vz, cols_out, c->n, co, l), m->d + idx * c->n, 1);
        }
    }
  return m;
}
static void pmat_free(pmat *m) {
  if (!m) return;
  free(m->d);
  free(m);
}

/* a: VecZnx(a_cols, a_size); uses columns col0 .. col0+cols_in-1 as the cols_in inputs.
 * big: VecZnx(cols_out, m->size) of exact i64 results. */
static void vmp_apply(const orc_ctx *c, const i64 *a, int a_cols, int a_size, int col0,
                      const pmat *m, i64 *big) {
  const int n = c->n;
  int rows = a_size < m->rows ? a_size : m->rows;
  tfe *atf = (tfe *)malloc(sizeof(tfe) * (size_t)rows * m->cols_in * n);
  for (int r = 0; r < rows; r++)
    for (int ci = 0; ci < m->cols_in; ci++)
      tf_forward(c, AT(a, a_cols, n, col0 + ci, r), atf + ((size_t)r * m->cols_in + ci) * n, 0);
  tfe *acc = (tfe *)malloc(sizeof(tfe) * n);
  for (int co = 0; co < m->cols_out; co++)
    for (int l = 0; l < m->size; l++) {
#ifdef ORC_FFT64
      const tfe *pa[32], *pb[32];
      int cnt = 0;
      for (int r = 0; r < rows; r++)
        for (int ci = 0; ci < m->cols_in; ci++) {
          size_t idx = (((size_t)r * m->cols_in + ci) * m->cols_out + co) * m->size + l;
          pa[cnt] = atf + ((size_t)r * m->cols_in + ci) * n;
          pb[cnt] = m->d + idx * n;
          cnt++;
        }
      tf_dot(c, acc, pa, pb, cnt);
#else
      tf_zero(c, acc);
      for (int r = 0; r < rows; r++)
        for (int ci = 0; ci < m->cols_in; ci++) {
          size_t idx = (((size_t)r * m->cols_in + ci) * m->cols_out + co) * m->size + l;
          tf_mac(c, acc, atf + ((size_t)r * m->cols_in + ci) * n, m->d + idx * n);
        }
#endif
      tf_inverse(c, acc, AT(big, m->cols_out, n, co, l));
    }
  free(acc);
  free(atf);
}

/* Poulpy [spec] glwe_external_product (rank 1, dsize 1): DFT all limbs of both columns,
 * vmp with the prepared GGSW, IDFT to big, normalise each column to res size. */
static void glwe_external_product(const orc_ctx *c, i64 *res, int res_size, const i64 *a,
                                  int a_size, const pmat *g) {
  const int n = c->n;
  i64 *big = (i64 *)malloc(sizeof(i64) * (size_t)2 * g->size * n);
  vmp_apply(c, a, 2, a_size, 0, g, big);
  for (int col = 0; col < 2; col++)
    vz_normalize(n, c->k, res, 2, col, res_size, big, 2, col, g->size);
  free(big);
  count_op(c, 0);
}

/* Poulpy [spec] glwe_keyswitch_internal: DFT the mask column, vmp with the key, IDFT, add the
 * body limbs to column 0 of the big result. */
static i64 *keyswitch_big(const orc_ctx *c, const i64 *a, int a_size, const pmat *key) {
  const int n = c->n;
  i64 *big = (i64 *)malloc(sizeof(i64) * (size_t)2 * key->size * n);
  vmp_apply(c, a, 2, a_size, 1, key, big);
  int lim = a_size < key->size ? a_size : key->size;
  for (int l = 0; l < lim; l++) {
    i64 *b0 = AT(big, 2, n, 0, l);
    const i64 *a0 = AT(a, 2, n, 0, l);
    for (int i = 0; i < n; i++) b0[i] += a0[i];
  }
  count_op(c, 1);
  return big;
}

/* Poulpy [spec] glwe_automorphism: key-switch, normalise, then X -> X^p on the small result */
static void glwe_automorphism(const orc_ctx *c, i64 *res, int res_size, const i64 *a, int a_size,
                              const pmat *key, i64 p) {
  const int n = c->n;
  i64 *big = keyswitch_big(c, a, a_size, key);
  i64 *tmp = (i64 *)malloc(sizeof(i64) * (size_t)2 * res_size * n);
  for (int col = 0; col < 2; col++)
    vz_normalize(n, c->k, tmp, 2, col, res_size, big, 2, col, key->size);
  glwe_small_automorphism(c, p, tmp, res, res_size);
  free(tmp);
  free(big);
}
/* Poulpy [spec] glwe_automorphism_add (sign=+1): res = a + phi(KS(a));
 * glwe_automorphism_sub_negate (sign=-1): res = a - phi(KS(a)).  The automorphism acts on the
 * big (un-normalised) key-switch output, the small input is added limb-wise, then one
 * normalisation. */
static void glwe_automorphism_addsub(const orc_ctx *c, i64 *res, int res_size, const i64 *a,
                                     int a_size, const pmat *key, i64 p, int sign) {
  const int n = c->n;
  i64 *big = keyswitch_big(c, a, a_size, key);
  i64 *pb = (i64 *)malloc(sizeof(i64) * (size_t)2 * key->size * n);
  int lim = a_size < key->size ? a_size : key->size;
  for (int col = 0; col < 2; col++)
    for (int l = 0; l < key->size; l++) {
      i64 *d = AT(pb, 2, n, col, l);
      poly_automorphism(n, p, AT(big, 2, n, col, l), d);
      if (sign < 0) for (int i = 0; i < n; i++) d[i] = -d[i];
      if (l < lim) {
        const i64 *s = AT(a, 2, n, col, l);
        for (int i = 0; i < n; i++) d[i] += s[i];
      }
    }
  for (int col = 0; col < 2; col++)
    vz_normalize(n, c->k, res, 2, col, res_size, pb, 2, col, key->size);
  free(pb);
  free(big);
}

/* ======================================================================================
 * Evaluation keys (src/keys.rs:27-71)
 * ==================================================================================== */
struct orc_keys {
  const orc_ctx *c;
  pmat *atk[32];   /* atk_glwe[gal_el(i)], keys.rs:39-49 */
  pmat *atk_inv;   /* keys.rs:50 */
  pmat *tsk;       /* keys.rs:51 */
};
orc_keys *orc_keys_prepare(const orc_ctx *c, const int64_t *atk_glwe, const int64_t *tsk,
                           const int64_t *atk_inv) { /* keys.rs:57-71 */
  orc_keys *k = (orc_keys *)calloc(1, sizeof(*k));
  k->c = c;
  for (int i = 0; i < c->log_n; i++)
    k->atk[i] = pmat_prepare(c, atk_glwe + (size_t)i * orc_atk_len(c), c->dnum_ct, 1, 2,
                             c->size_evk_trace);
  k->atk_inv = pmat_prepare(c, atk_inv, c->dnum_ggsw, 1, 2, c->size_evk_inv);
  k->tsk = pmat_prepare(c, tsk, c->dnum_ggsw, 1, 2, c->size_evk_inv);
  return k;
}
void orc_keys_free(orc_keys *k) {
  if (!k) return;
  for (int i = 0; i < 32; i++) pmat_free(k->atk[i]);
  pmat_free(k->atk_inv);
  pmat_free(k->tsk);
  free(k);
}

void orc_automorphism(const orc_ctx *c, const orc_keys *k, int gi, int mode, const int64_t *in,
                      int64_t *out) {
  size_t len = orc_glwe_len(c);
  i64 *tmp = (i64 *)malloc(sizeof(i64) * len);
  memcpy(tmp, in, sizeof(i64) * len);
  if (mode == 0) glwe_automorphism(c, out, c->size_ct, tmp, c->size_ct, k->atk[gi], c->gal[gi]);
  else glwe_automorphism_addsub(c, out, c->size_ct, tmp, c->size_ct, k->atk[gi], c->gal[gi],
                                mode == 1 ? 1 : -1);
  free(tmp);
}

/* Poulpy [spec] glwe_trace_inplace(start, end): for i in start..end { rsh(1); res += phi_i(res) } */
static void glwe_trace_inplace(const orc_ctx *c, const orc_keys *k, int start, int end, i64 *res) {
  size_t len = orc_glwe_len(c);
  i64 *tmp = (i64 *)malloc(sizeof(i64) * len);
  for (int i = start; i < end; i++) {
    glwe_rsh(c, 1, res, c->size_ct);
    memcpy(tmp, res, sizeof(i64) * len);
    glwe_automorphism_addsub(c, res, c->size_ct, tmp, c->size_ct, k->atk[i], c->gal[i], 1);
  }
  free(tmp);
}
void orc_trace(const orc_ctx *c, const orc_keys *k, int start, int end, const int64_t *in,
               int64_t *out) {
  if (out != in) memcpy(out, in, sizeof(i64) * orc_glwe_len(c));
  glwe_trace_inplace(c, k, start, end, out);
}

void orc_external_product(const orc_ctx *c, const int64_t *in, const int64_t *ggsw, int64_t *out) {
  pmat *g = pmat_prepare(c, ggsw, c->dnum_ct, 2, 2, c->size_addr);
  size_t len = orc_glwe_len(c);
  i64 *tmp = (i64 *)malloc(sizeof(i64) * len);
  memcpy(tmp, in, sizeof(i64) * len);
  glwe_external_product(c, out, c->size_ct, tmp, c->size_ct, g);
  free(tmp);
  pmat_free(g);
}

/* n ciphertexts against ONE GGSW (BASELINE.json config 2), the GGSW prepared once; one ciphertext per thread */
void orc_external_product_many(const orc_ctx *c, const int64_t *in, int n, const int64_t *ggsw, int64_t *out,
                               int threads) {
  pmat *g = pmat_prepare(c, ggsw, c->dnum_ct, 2, 2, c->size_addr);
  const size_t len = orc_glwe_len(c);
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(threads > 0 ? threads : 1)
#endif
  for (int i = 0; i < n; i++) {
    i64 *tmp = (i64 *)malloc(sizeof(i64) * len);
    memcpy(tmp, in + (size_t)i * len, sizeof(i64) * len);
    glwe_external_product(c, out + (size_t)i * len, c->size_ct, tmp, c->size_ct, g);
    free(tmp);
  }
  pmat_free(g);
}

/* ======================================================================================
 * GLWEPacker (Poulpy [spec], log_batch = 0): binary counter of log_n accumulators.
 * ==================================================================================== */
typedef struct { i64 *data; int value, control; } accum;
struct orc_packer {
  const orc_ctx *c;
  accum acc[32];
  int counter;
};
orc_packer *orc_packer_new(const orc_ctx *c) {
  orc_packer *p = (orc_packer *)calloc(1, sizeof(*p));
  p->c = c;
  for (int i = 0; i < c->log_n; i++) p->acc[i].data = (i64 *)calloc(orc_glwe_len(c), sizeof(i64));
  return p;
}
void orc_packer_free(orc_packer *p) {
  if (!p) return;
  for (int i = 0; i < 32; i++) free(p->acc[i].data);
  free(p);
}

/* combine at level i: a = a + b X^t + phi(a - b X^t), t = 2^(log_n-i-1), g = gal(i) */
static void packer_combine(const orc_ctx *c, const orc_keys *k, accum *acc, const i64 *b, int i) {
  const int S = c->size_ct;
  const size_t len = orc_glwe_len(c);
  i64 *a = acc->data;
  const i64 t = (i64)1 << (c->log_n - i - 1);
  if (acc->value) {
    if (b) {
      i64 *tmp_b = (i64 *)malloc(sizeof(i64) * len);
      glwe_rotate_inplace(c, -t, a, S);              /* a = a X^-t */
      memcpy(tmp_b, a, sizeof(i64) * len);
      glwe_sub_inplace(c, tmp_b, b, S);              /* tmp_b = a X^-t - b */
      glwe_rsh(c, 1, tmp_b, S);
      glwe_add_inplace(c, a, b, S);                  /* a = a X^-t + b */
      glwe_rsh(c, 1, a, S);
      glwe_normalize_inplace(c, tmp_b, S);
      i64 *in = (i64 *)malloc(sizeof(i64) * len);
      memcpy(in, tmp_b, sizeof(i64) * len);
      glwe_automorphism(c, tmp_b, S, in, S, k->atk[i], c->gal[i]); /* tmp_b = phi(tmp_b) */
      free(in);
      glwe_sub_inplace(c, a, tmp_b, S);              /* a = a X^-t + b - phi(a X^-t - b) */
      glwe_normalize_inplace(c, a, S);
      glwe_rotate_inplace(c, t, a, S);               /* a = a + b X^t + phi(a - b X^t) */
      free(tmp_b);
    } else {
      glwe_rsh(c, 1, a, S);
      i64 *in = (i64 *)malloc(sizeof(i64) * len);
      memcpy(in, a, sizeof(i64) * len);
      glwe_automorphism_addsub(c, a, S, in, S, k->atk[i], c->gal[i], 1); /* a = a + phi(a) */
      free(in);
    }
  } else if (b) {
    i64 *tmp_b = (i64 *)malloc(sizeof(i64) * len);
    glwe_rotate(c, t, b, tmp_b, S);                  /* tmp_b = b X^t */
    glwe_rsh(c, 1, tmp_b, S);
    glwe_automorphism_addsub(c, a, S, tmp_b, S, k->atk[i], c->gal[i], -1); /* a = b X^t - phi(b X^t) */
    free(tmp_b);
    acc->value = 1;
  }
}
static void pack_core(const orc_ctx *c, const orc_keys *k, const i64 *a, accum *accs, int i) {
  if (i == c->log_n) return;
  accum *prev = &accs[i];
  if (!prev->control) {
    if (a) {
      memcpy(prev->data, a, sizeof(i64) * orc_glwe_len(c));
      prev->value = 1;
    } else {
      prev->value = 0;
    }
    prev->control = 1;
  } else {
    packer_combine(c, k, prev, a, i);
    prev->control = 0;
    pack_core(c, k, prev->value ? prev->data : NULL, accs, i + 1);
  }
}
/* one GLWEPacker::combine at tree level `level` on an accumulator that holds a value (a, updated
 * in place) and an optional second operand b: exposed so tests can restate the level-parallel
 * schedule of the GPU path (and its sharded variant) with oracle arithmetic. */
void orc_packer_combine(const orc_ctx *c, const orc_keys *k, int level, int64_t *a, const int64_t *b) {
  accum acc;
  acc.data = a;
  acc.value = 1;
  acc.control = 1;
  packer_combine(c, k, &acc, b, level);
}
void orc_packer_add(orc_packer *p, const orc_keys *k, const int64_t *g) {
  assert(p->counter < p->c->n);
  pack_core(p->c, k, g, p->acc, 0);
  p->counter += 1;
}
void orc_packer_flush(orc_packer *p, int64_t *out) {
  assert(p->counter == p->c->n);
  memcpy(out, p->acc[p->c->log_n - 1].data, sizeof(i64) * orc_glwe_len(p->c));
  for (int i = 0; i < p->c->log_n; i++) p->acc[i].value = p->acc[i].control = 0;
  p->counter = 0;
}

/* ======================================================================================
 * GGSW(X^i) -> GGSW(X^-i)  (src/coordinate_prepared.rs:121-142 -> Poulpy [spec]
 * GGSW::automorphism = key-switch-automorphism of every row's column-0 GLWE with
 * atk_ggsw_inv, then ggsw_expand_row with the GGLWE->GGSW key)
 * ==================================================================================== */
void orc_ggsw_automorphism_inv(const orc_ctx *c, const orc_keys *k, const int64_t *in,
                               int64_t *out) {
  const int n = c->n, S = c->size_addr;
  const size_t glwe = (size_t)2 * S * n;
  for (int r = 0; r < c->dnum_ct; r++) {
    const i64 *src = in + ((size_t)r * 2 + 0) * glwe;
    i64 *dst0 = out + ((size_t)r * 2 + 0) * glwe;
    i64 *dst1 = out + ((size_t)r * 2 + 1) * glwe;
    glwe_automorphism(c, dst0, S, src, S, k->atk_inv, -1);
    /* expand: col1 = sum_l DFT(mask_l) * tsk[l]  + (0, body) */
    i64 *big = (i64 *)malloc(sizeof(i64) * (size_t)2 * k->tsk->size * n);
    vmp_apply(c, dst0, 2, S, 1, k->tsk, big);
    count_op(c, 1);
    for (int l = 0; l < S && l < k->tsk->size; l++) {
      i64 *m1 = AT(big, 2, n, 1, l);
      const i64 *b0 = AT(dst0, 2, n, 0, l);
      for (int i = 0; i < n; i++) m1[i] += b0[i];
    }
    for (int col = 0; col < 2; col++)
      vz_normalize(n, c->k, dst1, 2, col, S, big, 2, col, k->tsk->size);
    free(big);
  }
}

/* ======================================================================================
 * Client side: encryption / decryption (Poulpy [spec]; examples/fhe-ram.rs)
 * ==================================================================================== */
void orc_secret_gen(const orc_ctx *c, orc_source *xs, int64_t *sk) {
  /* GLWESecret::fill_ternary_prob(0.5): P(nonzero) = 0.5, sign uniform */
  for (int i = 0; i < c->n; i++) {
    u64 r = orc_source_next_u64(xs);
    sk[i] = (r & 1) ? ((r & 2) ? 1 : -1) : 0;
  }
}

/* GLWE encryption of (optional) plaintext limbs under `sk`:
 *   mask  a_l uniform in [-2^(K-1), 2^(K-1)) for every limb,
 *   body  = normalise( -a*s + e*2^-k_noise + (pt if pt_col == 0) ),
 *   if pt_col == 1 the plaintext is added to the mask AFTER the body was computed, so the
 *   ciphertext decrypts (b + a*s) to pt*s + e  (GGSW column 1).
 * pt: VecZnx(1, pt_size) or NULL. */
static void glwe_encrypt_sk(const orc_ctx *c, i64 *ct, int size, int k_noise, const i64 *pt,
                            int pt_size, int pt_col, const tfe *sk_tf, orc_source *xa,
                            orc_source *xe) {
  const int n = c->n, K = c->k;
  i64 *big = (i64 *)calloc((size_t)size * n, sizeof(i64));
  tfe *atf = (tfe *)malloc(sizeof(tfe) * n), *acc = (tfe *)malloc(sizeof(tfe) * n);
  for (int l = 0; l < size; l++) {
    i64 *a = AT(ct, 2, n, 1, l);
    for (int i = 0; i < n; i++) a[i] = get_digit(K, (i64)orc_source_next_u64(xa));
    tf_forward(c, a, atf, 0);
    tf_zero(c, acc);
    tf_mac(c, acc, atf, sk_tf);
    tf_inverse(c, acc, big + (size_t)l * n);
    for (int i = 0; i < n; i++) big[(size_t)l * n + i] = -big[(size_t)l * n + i];
  }
  if (pt && pt_col == 0)
    for (int l = 0; l < pt_size && l < size; l++)
      for (int i = 0; i < n; i++) big[(size_t)l * n + i] += pt[(size_t)l * n + i];
  int nl = ceil_div(k_noise, K) - 1, sh = (nl + 1) * K - k_noise;
  for (int i = 0; i < n; i++) big[(size_t)nl * n + i] += source_gauss(xe, 3.2, 19.2) << sh;
  vz_normalize(n, K, ct, 2, 0, size, big, 1, 0, size);
  if (pt && pt_col == 1)
    for (int l = 0; l < pt_size && l < size; l++) {
      i64 *a = AT(ct, 2, n, 1, l);
      for (int i = 0; i < n; i++) a[i] += pt[(size_t)l * n + i];
    }
  free(acc); free(atf); free(big);
}
static tfe *secret_prepare(const orc_ctx *c, const i64 *sk) {
  tfe *s = (tfe *)malloc(sizeof(tfe) * c->n);
  tf_forward(c, sk, s, 1);
  return s;
}
static void negacyclic_mul(const orc_ctx *c, const i64 *a, const i64 *b, i64 *out) {
  tfe *x = (tfe *)malloc(sizeof(tfe) * c->n), *y = (tfe *)malloc(sizeof(tfe) * c->n);
  tfe *acc = (tfe *)malloc(sizeof(tfe) * c->n);
  tf_forward(c, a, x, 0);
  tf_forward(c, b, y, 1);
  tf_zero(c, acc);
  tf_mac(c, acc, x, y);
  tf_inverse(c, acc, out);
  free(x); free(y); free(acc);
}

/* Poulpy [spec] GGSW::encrypt_sk: row r, column ci is a GLWE of m * 2^-(r+1)K placed in
 * component ci.  ggsw layout [row][ci] GLWE(size). */
static void ggsw_encrypt_sk(const orc_ctx *c, i64 *ggsw, const i64 *scalar, const tfe *sk_tf,
                            orc_source *xa, orc_source *xe) {
  const int n = c->n, S = c->size_addr;
  const size_t glwe = (size_t)2 * S * n;
  i64 *pt = (i64 *)calloc((size_t)S * n, sizeof(i64));
  for (int r = 0; r < c->dnum_ct; r++) {
    memset(pt, 0, sizeof(i64) * (size_t)S * n);
    memcpy(pt + (size_t)r * n, scalar, sizeof(i64) * n);
    for (int ci = 0; ci < 2; ci++)
      glwe_encrypt_sk(c, ggsw + ((size_t)r * 2 + ci) * glwe, S, c->p.k_addr, pt, r + 1, ci,
                      sk_tf, xa, xe);
  }
  free(pt);
}

/* Poulpy [spec] GGLWE key-switching key: row r encrypts msg * 2^-(r+1)K under sk_out */
static void gglwe_encrypt_sk(const orc_ctx *c, i64 *key, int rows, int size, int k_noise,
                             const i64 *msg, const tfe *sk_out_tf, orc_source *xa,
                             orc_source *xe) {
  const int n = c->n;
  const size_t glwe = (size_t)2 * size * n;
  i64 *pt = (i64 *)calloc((size_t)size * n, sizeof(i64));
  for (int r = 0; r < rows; r++) {
    memset(pt, 0, sizeof(i64) * (size_t)size * n);
    memcpy(pt + (size_t)r * n, msg, sizeof(i64) * n);
    glwe_encrypt_sk(c, key + (size_t)r * glwe, size, k_noise, pt, r + 1, 0, sk_out_tf, xa, xe);
  }
  free(pt);
}
static i64 mod_inverse_2n(i64 p, i64 two_n) { /* p odd */
  i64 pp = ((p % two_n) + two_n) % two_n;
  /* phi(2n) = n for n a power of two: p^(n-1) */
  return mod_pow_i64(pp, (u64)(two_n / 2 - 1), two_n);
}
/* Poulpy [spec] GLWEAutomorphismKey::encrypt_sk(p): key switches s -> phi_{p^-1}(s), so that
 * phi_p applied afterwards lands back under s. */
static void atk_encrypt_sk(const orc_ctx *c, i64 *key, int rows, int size, int k_noise, i64 p,
                           const i64 *sk, orc_source *xa, orc_source *xe) {
  i64 *sk_out = (i64 *)malloc(sizeof(i64) * c->n);
  poly_automorphism(c->n, mod_inverse_2n(p, 2 * (i64)c->n), sk, sk_out);
  tfe *so = secret_prepare(c, sk_out);
  gglwe_encrypt_sk(c, key, rows, size, k_noise, sk, so, xa, xe);
  free(so);
  free(sk_out);
}

void orc_keygen(const orc_ctx *c, const int64_t *sk, orc_source *xa, orc_source *xe,
                int64_t *atk_glwe, int64_t *tsk, int64_t *atk_inv) { /* keys.rs:135-180 */
  for (int i = 0; i < c->log_n; i++) /* keys.rs:158-165 */
    atk_encrypt_sk(c, atk_glwe + (size_t)i * orc_atk_len(c), c->dnum_ct, c->size_evk_trace,
                   c->p.k_evk_trace, c->gal[i], sk, xa, xe);
  { /* keys.rs:167-169 GGLWEToGGSWKey: encrypts s*s under s */
    i64 *s2 = (i64 *)malloc(sizeof(i64) * c->n);
    negacyclic_mul(c, sk, sk, s2);
    tfe *st = secret_prepare(c, sk);
    gglwe_encrypt_sk(c, tsk, c->dnum_ggsw, c->size_evk_inv, c->p.k_evk_ggsw_inv, s2, st, xa, xe);
    free(st);
    free(s2);
  }
  /* keys.rs:171-173 */
  atk_encrypt_sk(c, atk_inv, c->dnum_ggsw, c->size_evk_inv, c->p.k_evk_ggsw_inv, -1, sk, xa, xe);
}

/* Poulpy [spec] encode at precision k into `size` limbs: value * 2^-k, wrapped to the torus */
static void encode_coeff(const orc_ctx *c, i64 *pt, int size, int idx, i64 v, int k) {
  const int K = c->k, n = c->n;
  int l = ceil_div(k, K) - 1, sh = (l + 1) * K - k;
  i64 carry = 0, t = v << sh;
  for (int j = l; j >= 0; j--) {
    t += carry;
    i64 d = get_digit(K, t);
    carry = get_carry(K, t, d);
    if (j < size) pt[(size_t)j * n + idx] = d;
    t = 0;
  }
}

void orc_ram_encrypt(const orc_ctx *c, const uint8_t *data, const int64_t *sk, orc_source *xa,
                     orc_source *xe, int64_t *out) { /* ram.rs:129-167, 334-380 */
  const int n = c->n, ws = c->p.word_size;
  const u64 max_addr = c->p.max_addr;
  tfe *st = secret_prepare(c, sk);
  int pt_size = ceil_div(c->p.k_pt, c->k);
  i64 *pt = (i64 *)malloc(sizeof(i64) * (size_t)pt_size * n);
  for (int i = 0; i < ws; i++) {                     /* ram.rs:161-166 */
    for (int h = 0; h < c->n_glwe; h++) {            /* ram.rs:358-379 */
      memset(pt, 0, sizeof(i64) * (size_t)pt_size * n);
      for (int j = 0; j < n; j++) {
        u64 addr = (u64)h * n + j;
        i64 v = addr < max_addr ? (i64)(int8_t)data[addr * ws + i] : 0; /* ram.rs:364,367 */
        encode_coeff(c, pt, pt_size, j, v, c->p.k_pt);                  /* ram.rs:368 */
      }
      glwe_encrypt_sk(c, out + ((size_t)i * c->n_glwe + h) * orc_glwe_len(c), c->size_ct,
                      c->p.k_ct, pt, pt_size, 0, st, xa, xe);
    }
  }
  free(pt);
  free(st);
}

void orc_address_encrypt(const orc_ctx *c, uint32_t value, const int64_t *sk, orc_source *xa,
                         orc_source *xe, int64_t *out) { /* address.rs:86-109 */
  const int n = c->n;
  tfe *st = secret_prepare(c, sk);
  i64 *scalar = (i64 *)calloc(n, sizeof(i64));
  u64 remain2d = value;
  int g = 0;
  for (int ci = 0; ci < c->n_coord; ci++) {
    u64 max = 1;
    for (int d = 0; d < c->coord_len[ci]; d++) max <<= c->coord_digits[ci][d];
    i64 kval = (i64)(remain2d & (max - 1));
    i64 v = -kval;                                   /* address.rs:106 */
    /* coordinate.rs:121-180 */
    assert(llabs(v) < n);
    int sign = v > 0 ? 1 : (v < 0 ? -1 : 0);
    u64 remain = (u64)llabs(v);
    int tot_base = 0;
    for (int d = 0; d < c->coord_len[ci]; d++) {
      int base = c->coord_digits[ci][d];
      u64 mask = (1ull << base) - 1;
      u64 chunk = (remain & mask) << tot_base;       /* gap = 1, coordinate.rs:143,154 */
      size_t pos;
      if (sign < 0 && chunk != 0) { pos = n - chunk; scalar[pos] = -1; } /* :156-157 */
      else { pos = chunk; scalar[pos] = 1; }                             /* :159 */
      ggsw_encrypt_sk(c, out + (size_t)g * orc_ggsw_len(c), scalar, st, xa, xe);
      scalar[pos] = 0;
      remain >>= base;
      tot_base += base;
      g++;
    }
    remain2d /= max;                                 /* address.rs:107 */
  }
  free(scalar);
  free(st);
}

void orc_encrypt_byte(const orc_ctx *c, uint8_t value, const int64_t *sk, orc_source *xa,
                      orc_source *xe, int64_t *out) { /* examples/fhe-ram.rs:179-210 */
  const int n = c->n;
  tfe *st = secret_prepare(c, sk);
  int pt_size = ceil_div(c->p.k_pt, c->k);
  i64 *pt = (i64 *)calloc((size_t)pt_size * n, sizeof(i64));
  encode_coeff(c, pt, pt_size, 0, (i64)value, c->p.k_pt); /* :197 */
  glwe_encrypt_sk(c, out, c->size_ct, c->p.k_ct, pt, pt_size, 0, st, xa, xe);
  free(pt);
  free(st);
}

static void glwe_decrypt_generic(const orc_ctx *c, const i64 *ct, int size, const tfe *st,
                                 i64 *pt) {
  const int n = c->n;
  i64 *big = (i64 *)malloc(sizeof(i64) * (size_t)size * n);
  tfe *atf = (tfe *)malloc(sizeof(tfe) * n), *acc = (tfe *)malloc(sizeof(tfe) * n);
  for (int l = 0; l < size; l++) {
    tf_forward(c, AT(ct, 2, n, 1, l), atf, 0);
    tf_zero(c, acc);
    tf_mac(c, acc, atf, st);
    tf_inverse(c, acc, big + (size_t)l * n);
    const i64 *b = AT(ct, 2, n, 0, l);
    for (int i = 0; i < n; i++) big[(size_t)l * n + i] += b[i];
  }
  vz_normalize(n, c->k, pt, 1, 0, size, big, 1, 0, size);
  free(acc); free(atf); free(big);
}
void orc_glwe_decrypt(const orc_ctx *c, const int64_t *glwe, const int64_t *sk, int64_t *pt) {
  tfe *st = secret_prepare(c, sk);
  glwe_decrypt_generic(c, glwe, c->size_ct, st, pt);
  free(st);
}
void orc_ggsw_decrypt_row(const orc_ctx *c, const int64_t *ggsw, int row, int col_in,
                          const int64_t *sk, int64_t *pt) {
  tfe *st = secret_prepare(c, sk);
  const size_t glwe = (size_t)2 * c->size_addr * c->n;
  glwe_decrypt_generic(c, ggsw + ((size_t)row * 2 + col_in) * glwe, c->size_addr, st, pt);
  free(st);
}
void orc_decrypt_glwe(const orc_ctx *c, const int64_t *glwe, const int64_t *sk, int64_t want,
                      int64_t *value, double *noise) { /* examples/fhe-ram.rs:212-237 */
  const int n = c->n, S = c->size_ct;
  i64 *pt = (i64 *)malloc(sizeof(i64) * (size_t)S * n);
  orc_glwe_decrypt(c, glwe, sk, pt);
  int k = c->p.k_ct;
  int log_scale = k - c->p.k_pt;                      /* :229 */
  /* decode_coeff_i64(k, 0): sum limb_j << K*(S-1-j), k = S*K here */
  i64 v = 0;
  for (int j = 0; j < S; j++) v += pt[(size_t)j * n] << (c->k * (S - 1 - j));
  v >>= (S * c->k - k);
  i64 diff = v - want * ((i64)1 << log_scale);        /* :231 */
  *noise = log2((double)llabs(diff)) - (double)k;     /* :232 */
  *value = (i64)llround((double)v / exp2((double)log_scale)); /* :233-234 */
  free(pt);
}

/* ======================================================================================
 * Coordinate / Address / Ram (src/coordinate_prepared.rs, src/ram.rs)
 * ==================================================================================== */
/* CoordinatePrepared::product / product_inplace (coordinate_prepared.rs:147-177) */
static void coordinate_product(const orc_ctx *c, pmat **g, int ng, i64 *res, const i64 *a) {
  const size_t len = orc_glwe_len(c);
  i64 *tmp = (i64 *)malloc(sizeof(i64) * len);
  for (int i = 0; i < ng; i++) {
    memcpy(tmp, i == 0 ? a : res, sizeof(i64) * len);
    glwe_external_product(c, res, c->size_ct, tmp, c->size_ct, g[i]);
  }
  free(tmp);
}
void orc_coordinate_product(const orc_ctx *c, const int64_t *in, const int64_t *ggsws, int n,
                            int64_t *out) {
  pmat *g[16];
  for (int i = 0; i < n; i++)
    g[i] = pmat_prepare(c, ggsws + (size_t)i * orc_ggsw_len(c), c->dnum_ct, 2, 2, c->size_addr);
  coordinate_product(c, g, n, out, in);
  for (int i = 0; i < n; i++) pmat_free(g[i]);
}
/* CoordinatePrepared::prepare (coordinate_prepared.rs:104-116) */
static void coordinate_prepare(const orc_ctx *c, const i64 *addr, int coord, pmat **g) {
  int first = 0;
  for (int i = 0; i < coord; i++) first += c->coord_len[i];
  for (int d = 0; d < c->coord_len[coord]; d++)
    g[d] = pmat_prepare(c, addr + (size_t)(first + d) * orc_ggsw_len(c), c->dnum_ct, 2, 2,
                        c->size_addr);
}
/* CoordinatePrepared::prepare_inv (coordinate_prepared.rs:121-142) */
static void coordinate_prepare_inv(const orc_ctx *c, const orc_keys *k, const i64 *addr,
                                   int coord, pmat **g) {
  int first = 0;
  for (int i = 0; i < coord; i++) first += c->coord_len[i];
  i64 *tmp = (i64 *)malloc(sizeof(i64) * orc_ggsw_len(c));
  for (int d = 0; d < c->coord_len[coord]; d++) {
    orc_ggsw_automorphism_inv(c, k, addr + (size_t)(first + d) * orc_ggsw_len(c), tmp);
    g[d] = pmat_prepare(c, tmp, c->dnum_ct, 2, 2, c->size_addr);
  }
  free(tmp);
}
static void coordinate_free(const orc_ctx *c, int coord, pmat **g) {
  for (int d = 0; d < c->coord_len[coord]; d++) pmat_free(g[d]);
}

typedef struct {
  i64 **data;    /* n_glwe GLWE (ram.rs:299) */
  int n_tree;    /* ram.rs:300: levels; level sizes */
  int tree_size[8];
  i64 **tree[8];
  orc_packer *packer;
  int state;     /* ram.rs:302 */
  int loaded;
} subram;

struct orc_ram {
  const orc_ctx *c;
  subram *sub;
};

orc_ram *orc_ram_new(const orc_ctx *c) { /* ram.rs:59-69, 306-332 */
  orc_ram *r = (orc_ram *)calloc(1, sizeof(*r));
  r->c = c;
  r->sub = (subram *)calloc(c->p.word_size, sizeof(subram));
  for (int s = 0; s < c->p.word_size; s++) {
    subram *sr = &r->sub[s];
    sr->data = (i64 **)calloc(c->n_glwe, sizeof(i64 *));
    for (int h = 0; h < c->n_glwe; h++) sr->data[h] = (i64 *)calloc(orc_glwe_len(c), sizeof(i64));
    u64 n = (u64)c->n, mas = c->p.max_addr;
    if (mas > n) { /* ram.rs:315-324 */
      u64 size = (mas + n - 1) / n;
      while (size != 1) {
        size = (size + n - 1) / n;
        int lv = sr->n_tree++;
        sr->tree_size[lv] = (int)size;
        sr->tree[lv] = (i64 **)calloc(size, sizeof(i64 *));
        for (u64 t = 0; t < size; t++) sr->tree[lv][t] = (i64 *)calloc(orc_glwe_len(c), sizeof(i64));
      }
    }
    sr->packer = orc_packer_new(c);
  }
  return r;
}
void orc_ram_free(orc_ram *r) {
  if (!r) return;
  for (int s = 0; s < r->c->p.word_size; s++) {
    subram *sr = &r->sub[s];
    for (int h = 0; h < r->c->n_glwe; h++) free(sr->data[h]);
    free(sr->data);
    for (int lv = 0; lv < sr->n_tree; lv++) {
      for (int t = 0; t < sr->tree_size[lv]; t++) free(sr->tree[lv][t]);
      free(sr->tree[lv]);
    }
    orc_packer_free(sr->packer);
  }
  free(r->sub);
  free(r);
}
void orc_ram_load(orc_ram *r, const int64_t *cts) {
  const orc_ctx *c = r->c;
  for (int s = 0; s < c->p.word_size; s++) {
    for (int h = 0; h < c->n_glwe; h++)
      memcpy(r->sub[s].data[h], cts + ((size_t)s * c->n_glwe + h) * orc_glwe_len(c),
             sizeof(i64) * orc_glwe_len(c));
    r->sub[s].loaded = 1;
    r->sub[s].state = 0;
  }
}
void orc_ram_store(const orc_ram *r, int64_t *cts) {
  const orc_ctx *c = r->c;
  for (int s = 0; s < c->p.word_size; s++)
    for (int h = 0; h < c->n_glwe; h++)
      memcpy(cts + ((size_t)s * c->n_glwe + h) * orc_glwe_len(c), r->sub[s].data[h],
             sizeof(i64) * orc_glwe_len(c));
}
void orc_ram_tree_store(const orc_ram *r, int64_t *cts) {
  const orc_ctx *c = r->c;
  for (int s = 0; s < c->p.word_size; s++) {
    const subram *sr = &r->sub[s];
    const i64 *src = sr->n_tree ? sr->tree[sr->n_tree - 1][0] : sr->data[0];
    memcpy(cts + (size_t)s * orc_glwe_len(c), src, sizeof(i64) * orc_glwe_len(c));
  }
}
int orc_ram_state(const orc_ram *r) { return r->sub[0].state; }

/* SubRam::read (ram.rs:382-459).  `packer` may be a private one (read_many). */
static int subram_read(const orc_ctx *c, const subram *sr, orc_packer *packer, const i64 *addr,
                       const orc_keys *k, i64 *out) {
  if (sr->state) return -2; /* ram.rs:393-396 */
  const size_t len = orc_glwe_len(c);
  const int n = c->n, n2 = c->n_coord;
  if (n2 > 2) return -5; /* the reference's loop is only meaningful for n2 <= 2 */
  i64 *tmp_ct = (i64 *)calloc(len, sizeof(i64));
  i64 *result0 = (i64 *)calloc(len, sizeof(i64));
  pmat *g[16];
  for (int i = 0; i < n2; i++) {                                     /* :411 */
    coordinate_prepare(c, addr, i, g);                               /* :416-419 */
    if (i < n2 - 1) {                                                /* :421 */
      /* i == 0 here because n2 <= 2: res_prev = data, a single chunk since n_glwe <= n */
      for (int ch = 0; ch < c->n_glwe; ch += n) {                    /* :424 */
        int clen = c->n_glwe - ch < n ? c->n_glwe - ch : n;
        for (int j = 0; j < n; j++) {                                /* :425 */
          int j_rev = (int)orc_reverse_bits_msb((u64)j, (uint32_t)c->log_n); /* :426 */
          if (j_rev < clen) {
            coordinate_product(c, g, c->coord_len[i], tmp_ct, sr->data[ch + j_rev]); /* :429 */
            orc_packer_add(packer, k, tmp_ct);                       /* :435 */
          } else {
            orc_packer_add(packer, k, NULL);                         /* :437-443 */
          }
        }
      }
      orc_packer_flush(packer, tmp_ct);                              /* :448 */
      memcpy(result0, tmp_ct, sizeof(i64) * len);                    /* :449 */
    } else if (i == 0) {
      coordinate_product(c, g, c->coord_len[i], tmp_ct, sr->data[0]); /* :451 */
    } else {
      coordinate_product(c, g, c->coord_len[i], tmp_ct, result0);    /* :454 */
    }
    coordinate_free(c, i, g);
  }
  glwe_trace_inplace(c, k, 0, c->log_n, tmp_ct);                     /* :457 */
  memcpy(out, tmp_ct, sizeof(i64) * len);
  free(result0);
  free(tmp_ct);
  return 0;
}

/* Test-infrastructure knob: the sub-RAMs of one read / read_prepare_write / write are independent
 * (ram.rs:187-190 maps over them sequentially), so the checker may run them on several threads to keep the
 * BASELINE-size parity tests short.  Default 1 = the reference's sequential order; results do not depend on it. */
static int g_ram_threads = 1;
void orc_set_ram_threads(int n) { g_ram_threads = n > 0 ? n : 1; }

int orc_ram_read(orc_ram *r, const int64_t *addr, const orc_keys *k, int64_t *out) {
  const orc_ctx *c = r->c; /* ram.rs:172-191 */
  for (int s = 0; s < c->p.word_size; s++)
    if (!r->sub[s].loaded) return -1; /* :182-185 */
  int rc_all = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_ram_threads)
#endif
  for (int s = 0; s < c->p.word_size; s++) {
    int rc = subram_read(c, &r->sub[s], r->sub[s].packer, addr, k, out + (size_t)s * orc_glwe_len(c));
    if (rc) rc_all = rc;
  }
  return rc_all;
}

int orc_ram_read_many(orc_ram *r, const int64_t *addrs, int n_reads, const orc_keys *k,
                      int64_t *out, int threads) {
  const orc_ctx *c = r->c;
  const int ws = c->p.word_size;
  const size_t addr_len = (size_t)c->n_ggsw * orc_ggsw_len(c);
  int rc_all = 0;
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
#endif
  for (int t = 0; t < n_reads * ws; t++) {
    int q = t / ws, s = t % ws;
    orc_packer *pk = orc_packer_new(c);
    int rc = subram_read(c, &r->sub[s], pk, addrs + (size_t)q * addr_len, k,
                         out + ((size_t)q * ws + s) * orc_glwe_len(c));
    if (rc) rc_all = rc;
    orc_packer_free(pk);
  }
  return rc_all;
}

/* SubRam::read_prepare_write (ram.rs:461-542) */
static int subram_rpw(const orc_ctx *c, subram *sr, const i64 *addr, const orc_keys *k, i64 *out) {
  if (sr->state) return -2; /* :472-475 */
  const size_t len = orc_glwe_len(c);
  const int n = c->n, n2 = c->n_coord;
  if (n2 > 2) return -5;
  i64 *tmp_ct = (i64 *)calloc(len, sizeof(i64));
  pmat *g[16];
  for (int i = 0; i < n2; i++) {                                      /* :487 */
    i64 **res_prev = i == 0 ? sr->data : sr->tree[i - 1];            /* :490-494 */
    int n_prev = i == 0 ? c->n_glwe : sr->tree_size[i - 1];
    coordinate_prepare(c, addr, i, g);                               /* :496-499 */
    for (int h = 0; h < n_prev; h++)                                 /* :502-504 product_inplace */
      coordinate_product(c, g, c->coord_len[i], res_prev[h], res_prev[h]);
    if (i < n2 - 1) {                                                /* :506 */
      int t_idx = 0;
      for (int ch = 0; ch < n_prev; ch += n) {                       /* :510 */
        int clen = n_prev - ch < n ? n_prev - ch : n;
        for (int j = 0; j < n; j++) {                                /* :511 */
          int j_rev = (int)orc_reverse_bits_msb((u64)j, (uint32_t)c->log_n);
          orc_packer_add(sr->packer, k, j_rev < clen ? res_prev[ch + j_rev] : NULL); /* :513-517 */
        }
      }
      orc_packer_flush(sr->packer, tmp_ct);                          /* :521 */
      memcpy(sr->tree[i][t_idx], tmp_ct, sizeof(i64) * len);         /* :525-527 */
    }
    coordinate_free(c, i, g);
  }
  sr->state = 1;                                                     /* :533 */
  memcpy(out, n2 != 1 ? sr->tree[sr->n_tree - 1][0] : sr->data[0], sizeof(i64) * len); /* :534-538 */
  glwe_trace_inplace(c, k, 0, c->log_n, out);                        /* :540 */
  free(tmp_ct);
  return 0;
}
int orc_ram_read_prepare_write(orc_ram *r, const int64_t *addr, const orc_keys *k, int64_t *out) {
  const orc_ctx *c = r->c; /* ram.rs:196-222 */
  for (int s = 0; s < c->p.word_size; s++)
    if (!r->sub[s].loaded) return -1;
  int rc_all = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_ram_threads)
#endif
  for (int s = 0; s < c->p.word_size; s++) {
    int rc = subram_rpw(c, &r->sub[s], addr, k, out + (size_t)s * orc_glwe_len(c));
    if (rc) rc_all = rc;
  }
  return rc_all;
}

int orc_ram_write(orc_ram *r, const int64_t *w, const int64_t *addr, const orc_keys *k) {
  const orc_ctx *c = r->c; /* ram.rs:226-294 */
  const size_t len = orc_glwe_len(c);
  const int n = c->n, n2 = c->n_coord, S = c->size_ct, ws = c->p.word_size;
  if (n2 > 2) return -5;
  for (int s = 0; s < ws; s++) if (!r->sub[s].state) return -3; /* ram.rs:555-558 */
  /* write_first_step (ram.rs:544-577) */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_ram_threads)
#endif
  for (int s = 0; s < ws; s++) {
    subram *sr = &r->sub[s];
    i64 *tmp_a = (i64 *)malloc(sizeof(i64) * len);
    i64 *to = n2 != 1 ? sr->tree[sr->n_tree - 1][0] : sr->data[0];   /* :565-569 */
    orc_trace(c, k, 0, c->log_n, to, tmp_a);                         /* :572 */
    glwe_sub_inplace(c, to, tmp_a, S);                               /* :574 */
    glwe_add_inplace(c, to, w + (size_t)s * len, S);                 /* :575 */
    glwe_normalize_inplace(c, to, S);                                /* :576 */
    free(tmp_a);
  }
  pmat *g[16];
  for (int i = n2 - 2; i >= 0; i--) {                                /* :258 */
    coordinate_prepare_inv(c, k, addr, i + 1, g);                    /* :260-271 */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_ram_threads)
#endif
    for (int s = 0; s < ws; s++) {                                   /* write_mid_step :579-632 */
      subram *sr = &r->sub[s];
      i64 *tmp_a = (i64 *)malloc(sizeof(i64) * len);
      i64 **tree_hi = i == 0 ? sr->data : sr->tree[i - 1];           /* :599-604 */
      int n_hi = i == 0 ? c->n_glwe : sr->tree_size[i - 1];
      i64 **tree_lo = sr->tree[i];
      for (int ch = 0, j = 0; ch < n_hi; ch += n, j++) {             /* :606 */
        int clen = n_hi - ch < n ? n_hi - ch : n;
        i64 *ct_lo = tree_lo[j];                                     /* :608 */
        coordinate_product(c, g, c->coord_len[i + 1], ct_lo, ct_lo); /* :610 */
        for (int h = 0; h < clen; h++) {                             /* :612 */
          i64 *ct_hi = tree_hi[ch + h];
          orc_trace(c, k, 0, c->log_n, ct_hi, tmp_a);                /* :616 */
          glwe_sub_inplace(c, ct_hi, tmp_a, S);                      /* :617 */
          orc_trace(c, k, 0, c->log_n, ct_lo, tmp_a);                /* :621 */
          glwe_add_inplace(c, ct_hi, tmp_a, S);                      /* :625 */
          glwe_normalize_inplace(c, ct_hi, S);                       /* :626 */
          glwe_rotate_inplace(c, -1, ct_lo, S);                      /* :629 */
        }
      }
      free(tmp_a);
    }
    coordinate_free(c, i + 1, g);
  }
  coordinate_prepare_inv(c, k, addr, 0, g);                          /* :278-289 */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(g_ram_threads)
#endif
  for (int s = 0; s < ws; s++) {                                     /* write_last_step :634-649 */
    subram *sr = &r->sub[s];
    for (int h = 0; h < c->n_glwe; h++)
      coordinate_product(c, g, c->coord_len[0], sr->data[h], sr->data[h]); /* :644-646 */
    sr->state = 0;                                                   /* :648 */
  }
  coordinate_free(c, 0, g);
  return 0;
}
