// Round-2 design probe: the 2048-point shifted transform with TWO shared-memory exchanges instead of three.
// 16 complex points per thread, passes of 4 + 4 + 3 stages, 128 threads per polynomial, two polynomials per
// CTA of 256 threads.  The probe (a) checks the transform against a schoolbook
// negacyclic product, (b) times forward + pointwise product + inverse per polynomial with 1 and 2 CTAs per SM,
// to compare with the 8-points-per-thread transform of kernels.cuh (about 2.1 K cycles per transform and SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o t2_transform_probe t2_transform_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int kN = 4096, kM = 2048;
constexpr int kPad = 2176;  // 16 blocks of 136 slots: P(e) = e + (e >> 4)

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

struct Tw3 { double2 w8, w9a, w9b, w10[4]; };  // pass-3 twiddles of one thread (odd blocks = i * even blocks)

// forward transform of this group's polynomial: x[m] = z[t + 128 m] on entry, spectrum (t'', r) on exit
__device__ __forceinline__ void forward(double2 (&x)[16], double2* buf, const double2* tw2, const Tw3& t3, int t, int g) {
  fwd16<0>(x, [&](int sl, int bl) { return c_tw1[(1 << sl) + bl]; });
#pragma unroll
  for (int m = 0; m < 16; m++) buf[t + (t >> 4) + 136 * m] = x[m];
  gsync(g);
  const int a = t >> 3, c = t & 7;
#pragma unroll
  for (int m = 0; m < 16; m++) x[m] = buf[136 * a + 8 * m + c + (m >> 1)];
  const double2* ta = tw2 + 16 * a;
  fwd16<0>(x, [&](int sl, int bl) { return ta[(1 << sl) + bl]; });
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 16; m++) buf[136 * a + 8 * m + c + (m >> 1)] = x[m];
  __syncwarp();
  const int h = t & 7;
#pragma unroll
  for (int r = 0; r < 16; r++) x[r] = buf[136 * a + 17 * h + r];
  fwd16<1>(x, [&](int sl, int bl) {
    if (sl == 1) return (bl & 1) ? mul_i(t3.w8) : t3.w8;
    if (sl == 2) { const double2 w = (bl & 2) ? t3.w9b : t3.w9a; return (bl & 1) ? mul_i(w) : w; }
    const double2 w = t3.w10[bl >> 1];
    return (bl & 1) ? mul_i(w) : w;
  });
  gsync(g);  // the buffer may be overwritten by the next transform's block-level stores
}
__device__ __forceinline__ void inverse(double2 (&x)[16], double2* buf, const double2* tw2, const Tw3& t3, int t, int g) {
  const int a = t >> 3, c = t & 7, h = t & 7;
  inv16<1>(x, [&](int sl, int bl) {
    if (sl == 1) return (bl & 1) ? mul_i(t3.w8) : t3.w8;
    if (sl == 2) { const double2 w = (bl & 2) ? t3.w9b : t3.w9a; return (bl & 1) ? mul_i(w) : w; }
    const double2 w = t3.w10[bl >> 1];
    return (bl & 1) ? mul_i(w) : w;
  });
#pragma unroll
  for (int r = 0; r < 16; r++) buf[136 * a + 17 * h + r] = x[r];
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 16; m++) x[m] = buf[136 * a + 8 * m + c + (m >> 1)];
  const double2* ta = tw2 + 16 * a;
  inv16<0>(x, [&](int sl, int bl) { return ta[(1 << sl) + bl]; });
#pragma unroll
  for (int m = 0; m < 16; m++) buf[136 * a + 8 * m + c + (m >> 1)] = x[m];
  gsync(g);
#pragma unroll
  for (int m = 0; m < 16; m++) x[m] = buf[t + (t >> 4) + 136 * m];
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
  Tw3 t3;
  t3.w8 = zeta(8, 2u * t);
  t3.w9a = zeta(9, 4u * t);
  t3.w9b = zeta(9, 4u * t + 2);
#pragma unroll
  for (int k = 0; k < 4; k++) t3.w10[k] = zeta(10, 8u * t + 2 * k);
  __syncthreads();
  const int poly = blockIdx.x * 2 + g;
  const int* pa = A + (size_t)poly * kN;
  const int* pb = B + (size_t)poly * kN;
  double2 x[16];
  double2* sp = S + (size_t)poly * kM + t;  // spectrum of b at r * 128 + t: 512 B contiguous per warp access
#pragma unroll
  for (int m = 0; m < 16; m++) x[m] = make_double2((double)pb[t + 128 * m], (double)pb[t + 128 * m + kM]);
  forward(x, buf, tw2, t3, t, g);
#pragma unroll
  for (int r = 0; r < 16; r++) sp[128 * r] = x[r];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int m = 0; m < 16; m++) x[m] = make_double2((double)pa[t + 128 * m], (double)pa[t + 128 * m + kM]);
    forward(x, buf, tw2, t3, t, g);
#pragma unroll
    for (int r = 0; r < 16; r++) {
      const double2 u = x[r], v = __ldg(sp + 128 * r);
      x[r] = make_double2(u.x * v.x - u.y * v.y, u.x * v.y + u.y * v.x);
    }
    inverse(x, buf, tw2, t3, t, g);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
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
  const int smem = 256 * 16 + 2 * kPad * 16;
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
