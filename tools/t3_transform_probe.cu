// Round-2 probe t3: the 16-point transform with ONE shared-memory exchange; the warp-local exchange (exchange 2 of
// t2 / kernels_ks7.cuh) goes through TENSOR MEMORY instead: tcgen05.st 32x32b (thread i -> lane i) followed by
// tcgen05.ld 16x256b swaps two index bits between lanes and registers (tools/tmem_shape_probe.cu), so the passes
// become 4 + 4 | trip | 2 | trip | 1 stages.  Same correctness check and timing as t2_transform_probe.cu.
// (t2 header follows)
// Round-2 design probe: the 2048-point shifted transform with TWO shared-memory exchanges instead of three.
// 16 complex points per thread, passes of 4 + 4 + 3 stages, 128 threads per polynomial, two polynomials per
// CTA of 256 threads.  The probe (a) checks the transform against a schoolbook
// negacyclic product, (b) times forward + pointwise product + inverse per polynomial with 1 and 2 CTAs per SM,
// to compare with the 8-points-per-thread transform of kernels.cuh (about 2.1 K cycles per transform and SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o t2_transform_probe t2_transform_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int kN = 4096, kM = 2048;
constexpr int kPad = 2208;  // 16 blocks of 138 slots: P(e) = 138 (e >> 7) + (e & 127) + ((e & 127) >> 4)

__constant__ double2 c_tw1[16];  // pass 1: zeta(s, b) at (1 << s) + b, s < 4

__device__ __forceinline__ double2 mul_i(double2 w) { return make_double2(-w.y, w.x); }
__device__ __forceinline__ void bf(double2& x, double2& y, const double2 w) {
  double xr = fma(w.x, y.x, fma(-w.y, y.y, x.x));
  double xi = fma(w.x, y.y, fma(w.y, y.x, x.y));
  y.x = fma(2.0, x.x, -xr);
  y.y = fma(2.0, x.y, -xi);
  x.x = xr;
  x.y = xi;
}
__device__ __forceinline__ void ibf(double2& x, double2& y, const double2 w) {
  double dr = x.x - y.x, di = x.y - y.y;
  x.x += y.x;
  x.y += y.y;
  y.x = fma(dr, w.x, di * w.y);
  y.y = fma(di, w.x, -(dr * w.y));
}
__device__ __forceinline__ unsigned brev(unsigned b, int s) { return s ? __brev(b) >> (32 - s) : 0u; }
__device__ double2 zeta(int s, unsigned b) {
  double sn, cs;
  sincospi((0.25 + (double)brev(b, s)) / (double)(1u << s), &sn, &cs);
  return make_double2(cs, sn);
}
__device__ __forceinline__ void gsync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// stages sl = SL0 .. 3 of a 16-point register block; tw(sl, bl) = twiddle of local block bl at local stage sl
template <int SL0, typename F>
__device__ __forceinline__ void fwd16(double2 (&x)[16], F&& tw) {
#pragma unroll
  for (int sl = SL0; sl < 4; sl++) {
    const int half = 8 >> sl;
#pragma unroll
    for (int i = 0; i < 16; i++)
      if (!(i & half)) bf(x[i], x[i + half], tw(sl, i >> (4 - sl)));
  }
}
template <int SL0, typename F>
__device__ __forceinline__ void inv16(double2 (&x)[16], F&& tw) {
#pragma unroll
  for (int sl = 3; sl >= SL0; sl--) {
    const int half = 8 >> sl;
#pragma unroll
    for (int i = 0; i < 16; i++)
      if (!(i & half)) ibf(x[i], x[i + half], tw(sl, i >> (4 - sl)));
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#define R16(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
#define W16(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
// one plane (16 doubles d[v], v = column pair) of a thread: 32 columns of its own lane
__device__ __forceinline__ void st_plane_32(uint32_t ta, const double (&d)[16]) {
  uint32_t r[32];
#pragma unroll
  for (int v = 0; v < 16; v++) { r[2 * v] = (uint32_t)__double2loint(d[v]); r[2 * v + 1] = (uint32_t)__double2hiint(d[v]); }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta), R16(r) : "memory");
  uint32_t* q = r + 16;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta + 16), R16(q) : "memory");
}
__device__ __forceinline__ void ld_plane_32(uint32_t ta, double (&d)[16]) {
  uint32_t r[16], q[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : W16(r) : "r"(ta) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : W16(q) : "r"(ta + 16) : "memory");
  tm_wait_ld();
#pragma unroll
  for (int v = 0; v < 8; v++) { d[v] = __hiloint2double((int)r[2 * v + 1], (int)r[2 * v]); d[8 + v] = __hiloint2double((int)q[2 * v + 1], (int)q[2 * v]); }
}
// 16x256b.x4 at lane offset 16 h: register 4 xr + 2 q1 + half <-> lane 16 h + 8 q1 + (T >> 2), column 8 xr + 2 (T & 3) + half,
// i.e. thread T exchanges the doubles n = 4 xr + 2 h + q1 of its NEW register file with column pair 4 xr + (T & 3)
__device__ __forceinline__ void ld_plane_16x256(uint32_t ta, double (&d)[16]) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : W16(r) : "r"(ta + ((uint32_t)(16 * h) << 16)) : "memory");
    tm_wait_ld();
#pragma unroll
    for (int xr = 0; xr < 4; xr++)
#pragma unroll
      for (int q1 = 0; q1 < 2; q1++) d[4 * xr + 2 * h + q1] = __hiloint2double((int)r[4 * xr + 2 * q1 + 1], (int)r[4 * xr + 2 * q1]);
  }
}
__device__ __forceinline__ void st_plane_16x256(uint32_t ta, const double (&d)[16]) {
#pragma unroll
  for (int h = 0; h < 2; h++) {
    uint32_t r[16];
#pragma unroll
    for (int xr = 0; xr < 4; xr++)
#pragma unroll
      for (int q1 = 0; q1 < 2; q1++) {
        r[4 * xr + 2 * q1] = (uint32_t)__double2loint(d[4 * xr + 2 * h + q1]);
        r[4 * xr + 2 * q1 + 1] = (uint32_t)__double2hiint(d[4 * xr + 2 * h + q1]);
      }
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(ta + ((uint32_t)(16 * h) << 16)), R16(r) : "memory");
  }
}
// forward trip: register v of every lane -> (lanes' top two bits into registers, column-pair bits 1..0 into lanes);
// PERM: column pair of register v (a compile-time permutation choosing which register bits go to the lanes)
// loads without the wait (the caller waits once for both planes)
__device__ __forceinline__ void ld_raw_16x256(uint32_t ta, uint32_t (&r)[32]) {
  uint32_t* lo = r;
  uint32_t* hi = r + 16;
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : W16(lo) : "r"(ta) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : W16(hi) : "r"(ta + (16u << 16)) : "memory");
}
__device__ __forceinline__ void ld_raw_32(uint32_t ta, uint32_t (&r)[32]) {
  uint32_t* lo = r;
  uint32_t* hi = r + 16;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : W16(lo) : "r"(ta) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : W16(hi) : "r"(ta + 16) : "memory");
}
// both planes in flight: real parts through columns [scr, scr + 32), imaginary parts through [scr + 32, scr + 64)
template <typename P>
__device__ __forceinline__ void trip_fwd(double2 (&x)[16], uint32_t scr, P&& perm) {
  double d[16];
#pragma unroll
  for (int plane = 0; plane < 2; plane++) {
#pragma unroll
    for (int v = 0; v < 16; v++) d[perm(v)] = plane ? x[v].y : x[v].x;
    st_plane_32(scr + 32 * plane, d);
  }
  tm_wait_st();
  uint32_t re[32], im[32];
  ld_raw_16x256(scr, re);
  ld_raw_16x256(scr + 32, im);
  tm_wait_ld();
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int xr = 0; xr < 4; xr++)
#pragma unroll
      for (int q1 = 0; q1 < 2; q1++) {
        const int n = 4 * xr + 2 * h + q1, k = 16 * h + 4 * xr + 2 * q1;
        x[n] = make_double2(__hiloint2double((int)re[k + 1], (int)re[k]), __hiloint2double((int)im[k + 1], (int)im[k]));
      }
}
template <typename P>
__device__ __forceinline__ void trip_inv(double2 (&x)[16], uint32_t scr, P&& perm) {
  double d[16];
#pragma unroll
  for (int plane = 0; plane < 2; plane++) {
#pragma unroll
    for (int n = 0; n < 16; n++) d[n] = plane ? x[n].y : x[n].x;
    st_plane_16x256(scr + 32 * plane, d);
  }
  tm_wait_st();
  uint32_t re[32], im[32];
  ld_raw_32(scr, re);
  ld_raw_32(scr + 32, im);
  tm_wait_ld();
#pragma unroll
  for (int v = 0; v < 16; v++) {
    const int k = 2 * perm(v);
    x[v] = make_double2(__hiloint2double((int)re[k + 1], (int)re[k]), __hiloint2double((int)im[k + 1], (int)im[k]));
  }
}

// twiddles of the two short passes of one thread: W9[j] = zeta(9, 2 block8(e6 e5 = j)) (stage 8 uses its square,
// the odd stage-9 blocks i times it), W10[k] = zeta(10, block10 with (a1, e2) = k, e1 = 0) (odd blocks: i times it)
struct Tw3 { double2 w9[4], w10[4]; };
__device__ __forceinline__ double2 csq(double2 w) { return make_double2(fma(w.x, w.x, -w.y * w.y), 2.0 * w.x * w.y); }

// index bookkeeping (element index e, 11 bits; thread = warp w of the group, lane l):
//   pass 1   registers e[10:7], thread t = e[6:0]
//   pass 2   registers e[6:3], w = e[10:9], l = (e2 e1 e0 e8 e7)            [after the shared-memory exchange]
//   pass 3a  registers (e6 e5 e2 e1), l = (e0 e8 e7 e4 e3)                  [after trip 1]: stages 8, 9
//   pass 3b  registers (e2 e1 e0 e8), l = (e7 e4 e3 e6 e5)                  [after trip 2]: stage 10
__device__ __forceinline__ void forward(double2 (&x)[16], double2* buf, const double2* tw2, const Tw3& t3, uint32_t scr, int t, int g) {
  fwd16<0>(x, [&](int sl, int bl) { return c_tw1[(1 << sl) + bl]; });
#pragma unroll
  for (int m = 0; m < 16; m++) buf[138 * m + t + (t >> 4)] = x[m];
  gsync(g);
  const int w = t >> 5, l = t & 31;
  const int a = 4 * w + (l & 3), c = l >> 2;
#pragma unroll
  for (int m = 0; m < 16; m++) x[m] = buf[138 * a + 8 * m + c + (m >> 1)];
  gsync(g);  // the buffer may be overwritten by the next transform's stores (nothing else uses it)
  const double2* ta = tw2 + 16 * a;
  fwd16<0>(x, [&](int sl, int bl) {
    const double2 wv = ta[(1 << sl) + (bl & ~1)];
    return (sl > 0 && (bl & 1)) ? mul_i(wv) : wv;
  });
  trip_fwd(x, scr, [](int v) { return v; });                   // (e4 e3) to the lanes, (e2 e1) into the registers
  // registers n = (e6 e5 e2 e1): stage 8 pairs n ^ 2 (e2), stage 9 pairs n ^ 1 (e1)
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double2 w8 = csq(t3.w9[j]);
    bf(x[4 * j], x[4 * j + 2], w8);
    bf(x[4 * j + 1], x[4 * j + 3], w8);
    bf(x[4 * j], x[4 * j + 1], t3.w9[j]);
    bf(x[4 * j + 2], x[4 * j + 3], mul_i(t3.w9[j]));
  }
  trip_fwd(x, scr, [](int n) { return ((n & 3) << 2) | (n >> 2); });  // store order (e2 e1 e6 e5): (e6 e5) to the lanes
  // registers n' = (e2 e1 e0 e8): stage 10 pairs n' ^ 2 (e0); block = (.. e2 e1): twiddle W10[(e8, e2)], times i if e1
#pragma unroll
  for (int n = 0; n < 16; n++)
    if (!(n & 2)) {
      const int e2 = n >> 3, e1 = (n >> 2) & 1, a1 = n & 1;
      const double2 wv = t3.w10[2 * a1 + e2];
      bf(x[n], x[n + 2], e1 ? mul_i(wv) : wv);
    }
}
__device__ __forceinline__ void inverse(double2 (&x)[16], double2* buf, const double2* tw2, const Tw3& t3, uint32_t scr, int t, int g) {
#pragma unroll
  for (int n = 0; n < 16; n++)
    if (!(n & 2)) {
      const int e2 = n >> 3, e1 = (n >> 2) & 1, a1 = n & 1;
      const double2 wv = t3.w10[2 * a1 + e2];
      ibf(x[n], x[n + 2], e1 ? mul_i(wv) : wv);
    }
  trip_inv(x, scr, [](int n) { return ((n & 3) << 2) | (n >> 2); });
#pragma unroll
  for (int j = 0; j < 4; j++) {
    ibf(x[4 * j], x[4 * j + 1], t3.w9[j]);
    ibf(x[4 * j + 2], x[4 * j + 3], mul_i(t3.w9[j]));
    const double2 w8 = csq(t3.w9[j]);
    ibf(x[4 * j], x[4 * j + 2], w8);
    ibf(x[4 * j + 1], x[4 * j + 3], w8);
  }
  trip_inv(x, scr, [](int v) { return v; });
  const int w = t >> 5, l = t & 31;
  const int a = 4 * w + (l & 3), c = l >> 2;
  const double2* ta = tw2 + 16 * a;
  inv16<0>(x, [&](int sl, int bl) {
    const double2 wv = ta[(1 << sl) + (bl & ~1)];
    return (sl > 0 && (bl & 1)) ? mul_i(wv) : wv;
  });
#pragma unroll
  for (int m = 0; m < 16; m++) buf[138 * a + 8 * m + c + (m >> 1)] = x[m];
  gsync(g);
#pragma unroll
  for (int m = 0; m < 16; m++) x[m] = buf[138 * m + t + (t >> 4)];
  gsync(g);
  inv16<0>(x, [&](int sl, int bl) { return c_tw1[(1 << sl) + bl]; });
}

// out = a (*) b negacyclic, computed ITERS times per group; spectra of b prepared by the same forward
__global__ void __launch_bounds__(256, 2) probe(const int* A, const int* B, double2* S, long long* out, int iters, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem[];
  double2* tw2 = reinterpret_cast<double2*>(smem);  // [16][16]: pass-2 twiddles of block a at (1 << sl) + bl
  double2* bufs = tw2 + 256;
  const int g = threadIdx.x >> 7, t = threadIdx.x & 127;
  double2* buf = bufs + g * kPad;
  {
    const int a = threadIdx.x >> 4, k = threadIdx.x & 15;  // 256 threads fill the 16 x 16 table
    if (k >= 1) {
      int sl = 31 - __clz(k), bl = k - (1 << sl);
      tw2[16 * a + k] = zeta(4 + sl, ((unsigned)a << sl) + bl);
    }
  }
  uint32_t* slot = reinterpret_cast<uint32_t*>(bufs + 2 * kPad);
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  Tw3 t3;
  {
    const unsigned w = t >> 5, l = t & 31;
    // pass 3a: lane = (e0 e8 e7 e4 e3): block8 = (e10 e9 e8 e7 | e6 e5 | e4 e3)
    const unsigned a = 4 * w + ((l >> 2) & 3);
#pragma unroll
    for (unsigned j = 0; j < 4; j++) t3.w9[j] = zeta(9, 2u * ((a << 4) | (j << 2) | (l & 3)));
    // pass 3b: lane = (e7 e4 e3 e6 e5): block10 = (e10 e9 | e8 | e7 | e6 e5 | e4 e3 | e2 | e1)
#pragma unroll
    for (unsigned k = 0; k < 4; k++) {
      const unsigned a1 = k >> 1, e2 = k & 1;
      const unsigned blk = (w << 8) | (a1 << 7) | ((l >> 4) << 6) | ((l & 3) << 4) | (((l >> 2) & 3) << 2) | (e2 << 1);
      t3.w10[k] = zeta(10, blk);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t scr = *slot + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16) + 64 * g;  // 64 columns per group
  const int poly = blockIdx.x * 2 + g;
  const int* pa = A + (size_t)poly * kN;
  const int* pb = B + (size_t)poly * kN;
  double2 x[16];
  double2* sp = S + (size_t)poly * kM + t;  // spectrum of b at r * 128 + t: 512 B contiguous per warp access
#pragma unroll
  for (int m = 0; m < 16; m++) x[m] = make_double2((double)pb[t + 128 * m], (double)pb[t + 128 * m + kM]);
  forward(x, buf, tw2, t3, scr, t, g);
#pragma unroll
  for (int r = 0; r < 16; r++) sp[128 * r] = x[r];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int m = 0; m < 16; m++) x[m] = make_double2((double)pa[t + 128 * m], (double)pa[t + 128 * m + kM]);
    forward(x, buf, tw2, t3, scr, t, g);
#pragma unroll
    for (int r = 0; r < 16; r++) {
      const double2 u = x[r], v = __ldg(sp + 128 * r);
      x[r] = make_double2(u.x * v.x - u.y * v.y, u.x * v.y + u.y * v.x);
    }
    inverse(x, buf, tw2, t3, scr, t, g);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(*slot) : "memory");
#pragma unroll
  for (int m = 0; m < 16; m++) {
    out[(size_t)poly * kN + t + 128 * m] = __double2ll_rn(x[m].x * (1.0 / kM));
    out[(size_t)poly * kN + t + 128 * m + kM] = __double2ll_rn(x[m].y * (1.0 / kM));
  }
}

int main() {
  // pass-1 twiddles on the host (same formula)
  {
    double2 h[16];
    h[0] = make_double2(0, 0);
    for (int s = 0; s < 4; s++)
      for (unsigned b = 0; b < (1u << s); b++) {
        unsigned r = 0;
        for (int k = 0; k < s; k++) r |= ((b >> k) & 1u) << (s - 1 - k);
        const long double ang = 3.14159265358979323846264338327950288L * (0.25L + r) / (long double)(1u << s);
        h[(1 << s) + b] = make_double2((double)cosl(ang), (double)sinl(ang));
      }
    cudaMemcpyToSymbol(c_tw1, h, sizeof(h));
  }
  const int n_cta = 148 * 2, n_poly = n_cta * 2;
  std::vector<int> ha((size_t)n_poly * kN), hb((size_t)n_poly * kN);
  srand(1);
  for (auto& v : ha) v = (rand() % 131072) - 65536;
  for (auto& v : hb) v = (rand() % 131072) - 65536;
  int *da, *db; long long *dout, *dcyc;
  cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4);
  cudaMalloc(&dout, ha.size() * 8); cudaMalloc(&dcyc, n_cta * 8);
  double2* ds; cudaMalloc(&ds, (size_t)n_poly * kM * 16);
  cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 256 * 16 + 2 * kPad * 16 + 16;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // correctness: one iteration, polynomial 0 and the last one against the schoolbook product
  probe<<<n_cta, 256, smem>>>(da, db, ds, dout, 1, dcyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  std::vector<long long> ho(ha.size());
  cudaMemcpy(ho.data(), dout, ho.size() * 8, cudaMemcpyDeviceToHost);
  for (int poly : {0, n_poly - 1}) {
    const int* a = &ha[(size_t)poly * kN];
    const int* b = &hb[(size_t)poly * kN];
    long long bad = 0;
    for (int k = 0; k < kN; k += 37) {
      long long acc = 0;
      for (int i = 0; i < kN; i++) {
        const int j = (k - i) & (kN - 1);
        const long long p = (long long)a[i] * b[j];
        acc += (i + j == k) ? p : -p;
      }
      if (acc != ho[(size_t)poly * kN + k]) bad++;
    }
    printf("polynomial %d: %lld mismatching coefficients (of %d checked)\n", poly, bad, (kN + 36) / 37);
  }
  for (int ctas : {148, 296}) {
    const int iters = 200;
    probe<<<ctas, 256, smem>>>(da, db, ds, dout, iters, dcyc);
    cudaDeviceSynchronize();
    probe<<<ctas, 256, smem>>>(da, db, ds, dout, iters, dcyc);
    cudaDeviceSynchronize();
    std::vector<long long> hc(ctas);
    cudaMemcpy(hc.data(), dcyc, ctas * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : hc) mx = v > mx ? v : mx;
    const double per_cta_iter = (double)mx / iters;           // 2 polynomials x (forward + inverse) per CTA
    const int per_sm = ctas / 148;
    printf("%d CTA/SM: %.0f cycles per iteration and CTA = %.0f cycles per transform and SM (4 transforms per CTA-iteration)\n",
           per_sm, per_cta_iter, per_cta_iter / (4.0 * per_sm));
  }
  return 0;
}
