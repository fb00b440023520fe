// Bounce integrals of computehI: the "BEGIN INTEGRAl CALCULATION" block of
// src/ModRamScb.f90:372-410 -- per RAM field line (i,j) the arc length, the equatorial-B fix-up,
// the mirror fields, GSL_Integration_hI and GSL_BounceAverage (src/ModRamGSL.f90:125-200 ->
// integrator_c / bounceaverage_c, src/RamGSL.c:449-602) and the I_cart / H_cart / HDens_cart /
// bZEq_Cart assignments.
//
// The reference hands the integrands
//     f_I = sqrt(Bm - B(theta)),  f_h = 1/sqrt(Bm - B(theta)),  f_D = n(theta)/sqrt(Bm - B(theta))
// (zero where B >= Bm; B and n are gsl_interp_linear tables over chiVal, src/RamGSL.c:295-322,
// :326-448) to gsl_integration_cquad with epsabs = epsrel = 1e-3.  B and n are piecewise LINEAR in
// theta, so each of the three integrals has a closed form per grid segment; the device sums those
// closed forms instead of running an adaptive rule: the result is the exact value of the integral
// cquad approximates (parity bar: 1e-3, cquad's own tolerance -- cquad lives in un-vendored GNU GSL,
// so the bar is checked against an independent adaptive quadrature of the same integrands, tests/).
// The mirror-point search, the out-of-domain / short-span / non-positive fall-backs and the
// pitch-angle chain yI(L) <- yI(L+1) follow integrator_c / bounceaverage_c statement by statement.
//
// One CTA per field line, one thread per pitch angle; B, n, chi of the line in shared memory.
// Algorithmic bytes per line: 5 nthe + 3 NPA + 1 doubles (x, y, z, B, n in; I, h, HDens, bZEq out).
#pragma once
#include <cuda_runtime.h>

struct HiArgs {
  int nthe, nR, nT, nPa, nThetaEquator;   // nThetaEquator 1-based like the reference
  double bnormal;
  const double *chi, *mu, *x, *y, *z, *b, *dens;   // chi(nthe) mu(nPa); x,y,z,b,dens (nthe,nR,nT)
  const int* outside;                               // outsideMGNP(nR,nT)
  double *Icart, *Hcart, *Dcart, *bzeq;             // (nR,nT,nPa) x3, (nR,nT)
};

struct HiSeg { double I, H, V; };

// closed forms over one grid segment of width h where u = Bm - B runs linearly from u0 to u1 and the
// averaged variable from v0 to v1; only the part with u > 0 contributes.  With a = sqrt(ua), b = sqrt(ub)
// on the contributing part of width hp:  int sqrt(u) = (2 hp / 3)(ua + a b + ub)/(a + b),
// int 1/sqrt(u) = 2 hp/(a + b),  int t/sqrt(u) dt = (2/3)(2a + b)/(a + b)^2  (no cancellation as u0 -> u1)
__device__ __forceinline__ HiSeg hi_segment(double h, double u0, double u1, double v0, double v1) {
  HiSeg r = {0.0, 0.0, 0.0};
  if (u0 <= 0.0 && u1 <= 0.0) return r;
  double hp = h, ua = u0, ub = u1, va = v0, vb = v1;
  if (u1 <= 0.0) {               // mirrors inside the segment: keep [0, t*]
    const double t = u0 / (u0 - u1);
    hp = h * t; ub = 0.0; vb = v0 + (v1 - v0) * t;
  } else if (u0 <= 0.0) {        // keep [t*, 1]
    const double t = u0 / (u0 - u1);
    hp = h * (1.0 - t); ua = 0.0; va = v0 + (v1 - v0) * t;
  }
  const double a = sqrt(ua), b = sqrt(ub), s = a + b;
  r.I = (2.0 * hp / 3.0) * ((ua + a * b + ub) / s);
  r.H = 2.0 * hp / s;
  r.V = hp * (va * (2.0 / s) + (vb - va) * ((2.0 / 3.0) * (2.0 * a + b) / (s * s)));
  return r;
}

// integral over [chi[k0], chi[k1]] (grid nodes) for mirror field bm
__device__ __forceinline__ HiSeg hi_integrate(const double* chi, const double* bf, const double* var, int k0, int k1, double bm) {
  HiSeg t = {0.0, 0.0, 0.0};
  for (int k = k0; k < k1; k++) {
    const HiSeg s = hi_segment(chi[k + 1] - chi[k], bm - bf[k], bm - bf[k + 1], var[k], var[k + 1]);
    t.I += s.I; t.H += s.H; t.V += s.V;
  }
  return t;
}

// mirror-point search of integrator_c / bounceaverage_c (src/RamGSL.c:457-476, :547-566) and the domain
// tests that follow (:484-487, :575-578).  kind: 0 = baseline (mirrors outside the domain), 1 = span of
// <= 4 nodes (copy L+1), 2 = integrate between nodes LH and RH
__device__ __forceinline__ int hi_mirror_span(const double* chi, const double* bf, int n, double m, int* LHo, int* RHo) {
  double a = 0.0, b = 0.0;
  int LH = 0, RH = 0;
  for (int i = 1; i < n - 1; i++)
    if (m <= bf[i - 1] && m >= bf[i]) { a = chi[i - 1]; LH = i - 1; break; }
  for (int i = n - 2; i > 0; i--)
    if (m >= bf[i - 1] && m <= bf[i]) { b = chi[i]; RH = i; break; }
  if (m >= bf[1] || a == 0.0) a = chi[0];
  if (m >= bf[n - 1] || b == 0.0) b = chi[n - 1];
  *LHo = LH; *RHo = RH;
  if (a <= chi[0] || b >= chi[n - 1]) return 0;
  if (RH - LH <= 4) return 1;
  return 2;
}

// dynamic shared memory: chi, bf, dens, seg (nthe each), mir, yI, yH, yD (nPa each) doubles, kind (2 nPa ints)
__global__ void k_hi_lines(HiArgs A) {
  extern __shared__ double hi_sm[];
  const int n = A.nthe, nPa = A.nPa;
  double* chi = hi_sm;
  double* bf = chi + n;
  double* dn = bf + n;
  double* seg = dn + n;
  double* mir = seg + n;
  double* yI = mir + nPa;
  double* yH = yI + nPa;
  double* yD = yH + nPa;
  int* kind = (int*)(yD + nPa);
  int* kindD = kind + nPa;
  __shared__ double s_len, s_r0;

  const int line = blockIdx.x;                 // i + nR*j
  const int tid = threadIdx.x, nth = blockDim.x;
  const size_t nl = (size_t)A.nR * A.nT;
  if (A.outside[line] != 0) {                  // ModRamScb.f90:380; h_Cart, I_Cart, bZEq_Cart were zeroed at :236-237
    for (int L = tid; L < nPa; L += nth) { A.Icart[line + nl * L] = 0.0; A.Hcart[line + nl * L] = 0.0; }
    if (tid == 0) A.bzeq[line] = 0.0;
    return;
  }
  const size_t o = (size_t)line * n;
  const int ke = A.nThetaEquator - 1;
  for (int k = tid; k < n; k += nth) {
    chi[k] = A.chi[k]; bf[k] = A.b[o + k]; dn[k] = A.dens[o + k];
    if (k > 0) {
      const double dx = A.x[o + k] - A.x[o + k - 1], dy = A.y[o + k] - A.y[o + k - 1], dz = A.z[o + k] - A.z[o + k - 1];
      seg[k] = sqrt(dx * dx + dy * dy + dz * dz);
    }
  }
  __syncthreads();
  if (tid == 0) {
    double len = 0.0, bmin = bf[0];
    for (int k = 1; k < n; k++) { len += seg[k]; bmin = fmin(bmin, bf[k]); }     // :382-386, serial like the reference
    s_len = len;
    s_r0 = sqrt(A.x[o + ke] * A.x[o + ke] + A.y[o + ke] * A.y[o + ke]);
    if (fabs(bf[ke] - bmin) > 1e-9) {                                              // :388-394
      if (2.0 * bmin - bf[ke] > 0.0) bf[ke] = 2.0 * bmin - bf[ke];
      else bf[ke] = bmin - 0.01;
    }
  }
  __syncthreads();
  for (int L = tid; L < nPa; L += nth)                                             // :396-397
    mir[L] = (L < nPa - 1) ? bf[ke] / (1.0 - A.mu[L] * A.mu[L]) : bf[n - 1];
  __syncthreads();

  // GSL_Integration_hI: every pitch angle's candidate in parallel, the L+1 -> L chain resolved afterwards
  for (int L = tid; L < nPa; L += nth) {
    if (L == nPa - 1) {
      const HiSeg t = hi_integrate(chi, bf, dn, 0, n - 1, mir[nPa - 1]);
      yI[L] = t.I; yH[L] = t.H; kind[L] = 2;
    } else if (L > 0) {
      int LH, RH;
      const int kd = hi_mirror_span(chi, bf, n, mir[L], &LH, &RH);
      kind[L] = kd;
      if (kd == 2) {
        const HiSeg t = hi_integrate(chi, bf, dn, LH, RH, mir[L]);
        yI[L] = t.I; yH[L] = t.H;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    for (int L = nPa - 2; L > 0; L--) {                                            // src/RamGSL.c:482-505
      if (kind[L] == 0) { mir[L] = mir[nPa - 1]; yI[L] = yI[nPa - 1]; yH[L] = yH[nPa - 1]; }
      else if (kind[L] == 1) { yI[L] = yI[L + 1]; yH[L] = yH[L + 1]; }
      else {
        if (yI[L] <= 0.0) yI[L] = yI[L + 1];
        if (yH[L] <= 0.0) yH[L] = yH[L + 1];
      }
    }
    yI[0] = 0.0;
    yH[0] = yH[1];
  }
  __syncthreads();

  // GSL_BounceAverage with the mirror fields integrator_c left behind (bM is INOUT in both wrappers)
  for (int L = tid; L < nPa; L += nth) {
    if (L == nPa - 1) {
      yD[L] = hi_integrate(chi, bf, dn, 0, n - 1, mir[nPa - 1]).V;
      kindD[L] = 2;
    } else if (L > 0) {
      int LH, RH;
      const int kd = hi_mirror_span(chi, bf, n, mir[L], &LH, &RH);
      kindD[L] = kd;
      if (kd == 2) yD[L] = hi_integrate(chi, bf, dn, LH, RH, mir[L]).V;
    }
  }
  __syncthreads();
  if (tid == 0) {
    for (int L = nPa - 2; L > 0; L--) {                                            // src/RamGSL.c:396-416
      if (kindD[L] == 0) { mir[L] = mir[nPa - 1]; yD[L] = yD[nPa - 1]; }
      else if (kindD[L] == 1) yD[L] = yD[L + 1];
      else if (yD[L] <= 0.0) yD[L] = yD[L + 1];
    }
    yD[0] = yD[1];
  }
  __syncthreads();

  const double PI_D = 3.141592653589793238462643383279502884197;
  const double cI = s_len / (PI_D * s_r0), cH = s_len / (PI_D * 2.0 * s_r0);      // ModRamScb.f90:402-405
  for (int L = tid; L < nPa; L += nth) {
    const double sq = sqrt(mir[L]);
    A.Icart[line + nl * L] = cI * yI[L] / sq;
    A.Hcart[line + nl * L] = cH * yH[L] * sq;
    A.Dcart[line + nl * L] = yD[L] / yH[L];
  }
  if (tid == 0) A.bzeq[line] = bf[ke] * A.bnormal;
}

// =============================================================================
// computehI after the integral block (src/ModRamScb.f90:413-637): four small kernels over the (nR, nT, NPA)
// arrays.  Serial dependences of the reference are kept inside one thread (radial chains per (MLT, pitch angle),
// the descending pitch-angle repair per line); everything else is pointwise.  Reference operation order,
// -fmad=false: bit-identical to the oracle (scbo_hi_tail).
// =============================================================================
struct HiTailArgs {
  int nR, nT, nPa, smooth;
  double DthI, bnes1;                        // bnes1 = 0.32/LZ(1)**3/1.e4 (host)
  const int *ScaleAt, *outside;              // ScaleAt(nT) 1-based radial index or 0; outsideMGNP(nR,nT)
  const double *Lz, *PA, *PAbn;
  double *I0, *H0, *D0, *bz0;                // as the integral block left them
  double *I1, *H1, *D1, *bz1, *hI, *iI;      // repaired arrays, h / I at PAbn
  double *I2, *H2, *D2, *hI2, *iI2;          // smoothed
  double *FNHS, *FNIS, *BOUNHS, *BOUNIS, *HDNS, *BNES, *dIdt, *dHdt, *dIbndt, *dBdt;   // (nR+1,nT,NPA) / (nR+1,nT)
  int* fail;
  double w[81];                              // gaussian_kernel(1.0): 9 x 9, column-major
};

// one column (all radial points of one MLT js and one pitch angle L) after the outer-boundary scaling (:417-470)
struct HiCol {
  const double* a; size_t base, st; int nR, ii; bool scale; double s; const int* out;
  double prev;
  __device__ void init(const HiTailArgs& A, const double* arr, int js, int L) {
    a = arr; nR = A.nR; st = 1; base = (size_t)A.nR * (js + (size_t)A.nT * L);
    ii = A.ScaleAt[js]; out = A.outside + (size_t)A.nR * js;
    scale = (ii != 0) && js >= 1 && L >= 1;
    s = 0.0; prev = 0.0;
    if (scale) {
      const double fr = (A.Lz[ii - 1] - A.Lz[ii - 2]) / (A.Lz[ii - 3] - A.Lz[ii - 2]);
      const double t = a[base + ii - 2] + fr * (a[base + ii - 3] - a[base + ii - 2]);
      s = (t <= 0) ? a[base + ii - 2] / a[base + ii - 1] : t / a[base + ii - 1];
    }
  }
  // call with i = 0 .. nR-1 in order
  __device__ double get(int i) {
    double v = a[base + i];
    if (scale && i >= ii - 1) v = (out[i] == 0) ? v * s : prev;
    prev = v;
    return v;
  }
};

// thread per (MLT j, pitch angle L): scaling, MLT continuity (column 1 <- column nT), near-90-degree corrections,
// negative repair (:413-508); out of place (column 1 reads column nT)
__global__ void k_hi_tail_cols(HiTailArgs A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nT * A.nPa) return;
  const int j = t % A.nT, L = t / A.nT;
  const int js = (j == 0) ? A.nT - 1 : j;
  const int Ls = L < 3 ? 3 : L;
  const double* src[3] = {A.I0, A.H0, A.D0};
  double* dst[3] = {A.I1, A.H1, A.D1};
  const double f3[3] = {0.50, 0.99, 0.999}, f2[3] = {0.20, 0.99, 0.999};
  const size_t o = (size_t)A.nR * (j + (size_t)A.nT * L);
  for (int c = 0; c < 3; c++) {
    HiCol col;
    col.init(A, src[c], js, Ls);
    double last = 0.0;
    for (int i = 0; i < A.nR; i++) {
      double v = col.get(i);
      if (L < 3) {
        v = f3[c] * v;                                   // L = 3 (1-based)
        if (L < 2) v = f2[c] * v;                        // L = 2
        if (L < 1) v = (c == 0) ? 0.0 : f2[c] * v;       // L = 1
      }
      if (i >= 1 && v < 0) v = last;                     // :496-504 (a no-op unless something is negative)
      dst[c][o + i] = v;
      last = v;
    }
  }
  if (L == 0) {
    const int ii = A.ScaleAt[js];
    const bool sc = (ii != 0) && js >= 1 && A.nPa >= 2;
    double prev = 0.0;
    for (int i = 0; i < A.nR; i++) {
      double v = A.bz0[i + (size_t)A.nR * js];
      if (sc && i >= ii - 1) v = prev;
      A.bz1[i + (size_t)A.nR * j] = v;
      prev = v;
    }
  }
}

// CTA per (i, j) line: "too large" repair descending in L (:509-517), then GSL_Interpolation_1D (Steffen) of h and I
// from PA(NPA:1:-1) onto PAbn(NPA-1:2:-1) (:519-532).  Dynamic shared memory: 8 NPA doubles.
__global__ void k_hi_tail_lines(HiTailArgs A) {
  extern __shared__ double hi_sm[];
  const int nPa = A.nPa;
  double* lI = hi_sm;            // the line, natural order
  double* lH = lI + nPa;
  double* lD = lH + nPa;
  double* xa = lD + nPa;         // filtered ascending abscissae and values
  double* fh = xa + nPa;
  double* fi = fh + nPa;
  double* yh = fi + nPa;
  double* yi = yh + nPa;
  __shared__ int s_n1;
  const int line = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const size_t nl = (size_t)A.nR * A.nT;
  for (int L = tid; L < nPa; L += nth) { lI[L] = A.I1[line + nl * L]; lH[L] = A.H1[line + nl * L]; lD[L] = A.D1[line + nl * L]; }
  __syncthreads();
  if (tid == 0) {
    for (int L = nPa - 2; L >= 0; L--) {
      if (lI[L] > lI[L + 1]) lI[L] = 0.99 * lI[L + 1];
      if (lH[L] > lH[L + 1]) lH[L] = 0.99 * lH[L + 1];
      if (lD[L] > lD[L + 1]) lD[L] = 0.999 * lD[L + 1];
    }
    int n1 = 1;                                          // src/ModRamGSL.f90:262-273
    xa[0] = A.PA[nPa - 1]; fh[0] = lH[nPa - 1]; fi[0] = lI[nPa - 1];
    for (int q = 1; q < nPa; q++) {
      const double xv = A.PA[nPa - 1 - q];
      if (xv > xa[n1 - 1]) { xa[n1] = xv; fh[n1] = lH[nPa - 1 - q]; fi[n1] = lI[nPa - 1 - q]; n1++; }
    }
    s_n1 = n1;
  }
  __syncthreads();
  const int n1 = s_n1;
  for (int L = tid; L < nPa; L += nth) { A.I1[line + nl * L] = lI[L]; A.H1[line + nl * L] = lH[L]; A.D1[line + nl * L] = lD[L]; }
  if (n1 < 3) { if (tid == 0) atomicAdd(A.fail, 1); return; }
  for (int i = tid; i < n1; i += nth) {                  // steffen_init
    for (int c = 0; c < 2; c++) {
      const double* fa = c ? fi : fh;
      double yp;
      if (i == 0) yp = (fa[1] - fa[0]) / (xa[1] - xa[0]);
      else if (i == n1 - 1) yp = (fa[n1 - 1] - fa[n1 - 2]) / (xa[n1 - 1] - xa[n1 - 2]);
      else {
        const double hi = xa[i + 1] - xa[i], him1 = xa[i] - xa[i - 1];
        const double si = (fa[i + 1] - fa[i]) / hi, sim1 = (fa[i] - fa[i - 1]) / him1;
        const double pi = (sim1 * hi + si * him1) / (him1 + hi);
        const double m1 = fabs(si) < 0.5 * fabs(pi) ? fabs(si) : 0.5 * fabs(pi);
        const double m2 = fabs(sim1) < m1 ? fabs(sim1) : m1;
        yp = (((sim1 < 0) ? -1.0 : 1.0) + ((si < 0) ? -1.0 : 1.0)) * m2;
      }
      (c ? yi : yh)[i] = yp;
    }
  }
  __syncthreads();
  bool bad = false;
  for (int q = tid; q < nPa - 2; q += nth) {
    const int Lt = nPa - 2 - q;                          // 0-based pitch-angle index of target q
    const double xb = A.PAbn[Lt];
    double r[2];
    for (int c = 0; c < 2; c++) {
      const double* fa = c ? fi : fh;
      const double* yp = c ? yi : yh;
      if (xb <= xa[0]) r[c] = fa[0] + (xb - xa[0]) / (xa[1] - xa[0]) * (fa[1] - fa[0]);
      else if (xb >= xa[n1 - 1]) r[c] = fa[n1 - 1] + (xb - xa[n1 - 1]) / (xa[n1 - 2] - xa[n1 - 1]) * (fa[n1 - 2] - fa[n1 - 1]);
      else if (xb == xb) {
        int ilo = 0, ihi = n1 - 1;
        while (ihi > ilo + 1) {
          const int m = (ihi + ilo) / 2;
          if (xa[m] > xb) ihi = m; else ilo = m;
        }
        const double hi = xa[ilo + 1] - xa[ilo], delx = xb - xa[ilo];
        const double si = (fa[ilo + 1] - fa[ilo]) / hi;
        const double a = (yp[ilo] + yp[ilo + 1] - 2 * si) / hi / hi;
        const double b = (3 * si - 2 * yp[ilo] - yp[ilo + 1]) / hi;
        r[c] = fa[ilo] + delx * (yp[ilo] + delx * (b + delx * a));
      } else { bad = true; r[c] = xb; }
    }
    A.hI[line + nl * Lt] = r[0];
    A.iI[line + nl * Lt] = r[1];
    if (Lt == nPa - 2) { A.hI[line + nl * (nPa - 1)] = r[0]; A.iI[line + nl * (nPa - 1)] = r[1]; }
    if (Lt == 1) { A.hI[line + nl * 0] = r[0]; A.iI[line + nl * 0] = r[1]; }
  }
  if (bad) atomicAdd(A.fail, 1);
}

// Gaussian smoothing of the five arrays, planes L = 2 .. NPA (:539-563): thread per (i, j, L), 9 x 9 window on the
// 3 x 3 reflected tiling of convolve (srcExternal/gaussian_filter.f90:98-187), products summed in array element order
__global__ void k_hi_smooth(HiTailArgs A) {
  const size_t nl = (size_t)A.nR * A.nT;
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= nl * A.nPa) return;
  const int i = (int)(t % A.nR), j = (int)((t / A.nR) % A.nT), L = (int)(t / nl);
  const double* src[5] = {A.H1, A.I1, A.hI, A.iI, A.D1};
  double* dst[5] = {A.H2, A.I2, A.hI2, A.iI2, A.D2};
  if (L == 0) {
    for (int c = 0; c < 5; c++) dst[c][t] = src[c][t];
    return;
  }
  double sum[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int dj = -4; dj <= 4; dj++) {
    int jj = j + dj;
    jj = jj < 0 ? -1 - jj : (jj >= A.nT ? 2 * A.nT - 1 - jj : jj);
    for (int di = -4; di <= 4; di++) {
      int ir = i + di;
      ir = ir < 0 ? -1 - ir : (ir >= A.nR ? 2 * A.nR - 1 - ir : ir);
      const double w = A.w[(di + 4) + 9 * (dj + 4)];
      const size_t o = ir + (size_t)A.nR * jj + nl * L;
      for (int c = 0; c < 5; c++) sum[c] += w * src[c][o];
    }
  }
  for (int c = 0; c < 5; c++) dst[c][t] = sum[c];
}

// thread per (MLT J, pitch angle L): the RAM variables, their time derivatives, the I = 1 row and the NaN repair
// (:566-637), serial in the radial index like the reference
__global__ void k_hi_fill(HiTailArgs A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nT * A.nPa) return;
  const int J = t % A.nT, L = t / A.nT, nR = A.nR;
  const double* H = A.smooth ? A.H2 : A.H1;
  const double* I = A.smooth ? A.I2 : A.I1;
  const double* D = A.smooth ? A.D2 : A.D1;
  const double* hI = A.smooth ? A.hI2 : A.hI;
  const double* iI = A.smooth ? A.iI2 : A.iI;
  const size_t oc = (size_t)nR * (J + (size_t)A.nT * L);              // column in (nR,nT,NPA)
  const size_t orr = (size_t)(nR + 1) * (J + (size_t)A.nT * L);        // column in (nR+1,nT,NPA)
  const bool still = fabs(A.DthI) <= 1e-9;
  for (int i = 1; i <= nR; i++) {                                      // 0-based row of the RAM arrays
    const double pI = A.FNIS[orr + i], pH = A.FNHS[orr + i], pbI = A.BOUNIS[orr + i];
    const double nH = H[oc + i - 1], nI = I[oc + i - 1], nbI = iI[oc + i - 1];
    A.FNHS[orr + i] = nH; A.FNIS[orr + i] = nI;
    A.HDNS[orr + i] = D[oc + i - 1];
    A.BOUNHS[orr + i] = hI[oc + i - 1]; A.BOUNIS[orr + i] = nbI;
    A.dIdt[orr + i] = still ? 0.0 : (nI - pI) / A.DthI;
    A.dHdt[orr + i] = still ? 0.0 : (nH - pH) / A.DthI;
    A.dIbndt[orr + i] = still ? 0.0 : (nbI - pbI) / A.DthI;
  }
  A.FNHS[orr] = A.FNHS[orr + 1]; A.FNIS[orr] = A.FNIS[orr + 1];
  A.BOUNHS[orr] = A.BOUNHS[orr + 1]; A.BOUNIS[orr] = A.BOUNIS[orr + 1];
  A.HDNS[orr] = A.HDNS[orr + 1];
  A.dIdt[orr] = 0.0; A.dHdt[orr] = 0.0; A.dIbndt[orr] = 0.0;
  for (int i = 1; i <= nR; i++) {
    if (isnan(A.FNIS[orr + i])) A.FNIS[orr + i] = A.FNIS[orr + i - 1];
    if (isnan(A.FNHS[orr + i])) A.FNHS[orr + i] = A.FNHS[orr + i - 1];
    if (isnan(A.BOUNIS[orr + i])) A.BOUNIS[orr + i] = A.BOUNIS[orr + i - 1];
    if (isnan(A.BOUNHS[orr + i])) A.BOUNHS[orr + i] = A.BOUNHS[orr + i - 1];
    if (isnan(A.HDNS[orr + i])) A.HDNS[orr + i] = A.HDNS[orr + i - 1];
    if (isnan(A.dIdt[orr + i])) A.dIdt[orr + i] = 0.0;
    if (isnan(A.dIbndt[orr + i])) A.dIbndt[orr + i] = 0.0;
  }
  if (L == 0) {
    const size_t o2 = (size_t)(nR + 1) * J;
    for (int i = 1; i <= nR; i++) {
      const double prev = A.BNES[o2 + i];
      const double v = A.bz1[i - 1 + (size_t)nR * J] / 1e9;
      A.BNES[o2 + i] = v;
      A.dBdt[o2 + i] = still ? 0.0 : (v - prev) / A.DthI;
    }
    A.BNES[o2] = A.bnes1;
    A.dBdt[o2] = 0.0;
  }
}

// =============================================================================
// computehI, "Convert SCB field lines to RAM field lines" (src/ModRamScb.f90:252-300): for every RAM equatorial point
// the winding-number test against the outer SCB ring, psiRAM by GSL_Interpolation_2D over the equatorial (x, y)
// scatter, then x, y, z, bf at every node k by GSL_Interpolation_2D over the (psi, alfa) scatter of that k.  The
// generic is Interpolation_2D_NN_point (src/ModRamGSL.f90:368-422): nine MINLOC passes over npsi (nzeta-1) squared
// distances, then NN_Interpolation_2D (:872-917).  This is the brute-force part of the routine (NR NT nthe queries x
// 4320 candidates x 9 passes in the reference).
//
// One CTA per (point set, slice of the queries): the candidate coordinates sit in shared memory in the reference's
// scatter order; a warp takes a query.  MINLOC + overwriting the pick with 999999.9, nine times, yields the nine
// smallest (distance, index) pairs in lexicographic order (first minimum wins ties): nn9_select below.  Weights and the 1 or 4
// weighted sums in the reference's order: bit-identical to the oracle.
// MODE 0: point set = equatorial plane (x, y) -> psiRAM, outsideSCB;  MODE 1: point set k = blockIdx.x, (psi, alfa) -> x, y, z, bf.
// =============================================================================
struct HiConvArgs {
  int nthe, npsi, nzeta, nR, nT, nThetaEquator;
  const double *x, *y, *z, *bf, *psi, *alfa;     // (nthe,npsi,nzeta+1)
  const double *qx, *qy, *alphaRAM;              // xo, yo (nR,nT); alphaRAM(nT)  (host tables: cos / sin of libm)
  double* psiRAM;                                // (nR,nT)
  int* outside;                                  // outsideSCB(nR,nT)
  double *xRAM, *yRAM, *zRAM, *bRAM;             // (nthe,nR,nT)
};

// The nine picks of Interpolation_2D_NN_point (src/ModRamGSL.f90:368-422: MINLOC, overwrite the pick with 999999.9, nine
// times) = the nine smallest (distance, index) pairs in lexicographic order.  A warp takes a query in two scans of the
// candidates in shared memory: the first finds every lane's own smallest distance; the ninth smallest of those 32 values
// bounds the ninth pick from above (nine distinct candidates are no farther), so the second scan runs the sorted insertion
// only for the handful of candidates within the bound instead of through the warm-up of every lane's list.  Nine shuffle
// elections then pop the picks from the lane heads.  The distances are the same expression in both scans (-fmad=false).
__device__ __forceinline__ void nn9_select(const double* __restrict__ cx, const double* __restrict__ cy, int M, double x2, double y2,
                                           int lane, int near[9]) {
  const double BIG = 1.7976931348623157e308;
  double m0 = BIG, m1 = BIG;
  int s = lane;
  for (; s + 32 < M; s += 64) {                                          // two independent chains
    const double dx0 = cx[s] - x2, dy0 = cy[s] - y2, dx1 = cx[s + 32] - x2, dy1 = cy[s + 32] - y2;
    const double d0 = dx0 * dx0 + dy0 * dy0, d1 = dx1 * dx1 + dy1 * dy1;
    m0 = d0 < m0 ? d0 : m0;
    m1 = d1 < m1 ? d1 : m1;
  }
  if (s < M) {
    const double dx0 = cx[s] - x2, dy0 = cy[s] - y2;
    const double d0 = dx0 * dx0 + dy0 * dy0;
    m0 = d0 < m0 ? d0 : m0;
  }
  double v = m1 < m0 ? m1 : m0, bound = BIG;
#pragma unroll 1
  for (int r = 0; r < 9; r++) {                                          // ninth smallest of the lane minima
    double b = v;
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, b, o);
      b = ob < b ? ob : b;
    }
    bound = b;
    int c = (v == b) ? lane : 32;
    for (int o = 16; o > 0; o >>= 1) {
      const int oc = __shfl_xor_sync(0xffffffffu, c, o);
      c = oc < c ? oc : c;
    }
    if (lane == c) v = BIG;
  }
  double ld[9];
  int li[9];
#pragma unroll
  for (int t = 0; t < 9; t++) { ld[t] = BIG; li[t] = 0x7fffffff; }
  for (s = lane; s < M; s += 32) {
    const double dx = cx[s] - x2, dy = cy[s] - y2;
    const double d = dx * dx + dy * dy;
    if (d <= bound && (d < ld[8] || (d == ld[8] && s < li[8]))) {
      ld[8] = d; li[8] = s;
#pragma unroll
      for (int t = 8; t > 0; t--) {
        if (ld[t] < ld[t - 1] || (ld[t] == ld[t - 1] && li[t] < li[t - 1])) {
          const double td = ld[t]; ld[t] = ld[t - 1]; ld[t - 1] = td;
          const int ti = li[t]; li[t] = li[t - 1]; li[t - 1] = ti;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 9; r++) {
    double bd = ld[0];
    int bi = li[0];
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    near[r] = bi;
    if (li[0] == bi) {                                                   // the owner pops its head
#pragma unroll
      for (int t = 0; t < 8; t++) { ld[t] = ld[t + 1]; li[t] = li[t + 1]; }
      ld[8] = BIG; li[8] = 0x7fffffff;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_hi_nn9(HiConvArgs A) {
  extern __shared__ double hi_sm[];
  const int m1 = A.nzeta - 1, M = A.npsi * m1;
  double* cx = hi_sm;
  double* cy = cx + M;
  const size_t sj = A.nthe, sk = (size_t)A.nthe * A.npsi;
  const int k = (MODE == 0) ? A.nThetaEquator - 1 : blockIdx.x;
  const double* px = (MODE == 0) ? A.x : A.psi;
  const double* py = (MODE == 0) ? A.y : A.alfa;
  for (int s = threadIdx.x; s < M; s += blockDim.x) {
    const size_t o = k + sj * (s / m1) + sk * (s % m1 + 1);
    cx[s] = px[o]; cy[s] = py[o];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nq = A.nR * A.nT;
  for (int q = blockIdx.y * nwarp + warp; q < nq; q += gridDim.y * nwarp) {
    const int j = q / A.nR;
    double x2, y2;
    if (MODE == 0) {
      x2 = A.qx[q]; y2 = A.qy[q];
      int wn = 0;                                                        // :262-280
      const size_t ring = (A.nThetaEquator - 1) + sj * (A.npsi - 2);
      for (int kz = lane; kz < A.nzeta; kz += 32) {
        const double yn = A.y[ring + sk * kz], yp = A.y[ring + sk * (kz + 1)], xn = A.x[ring + sk * kz], xp = A.x[ring + sk * (kz + 1)];
        if (yn <= y2) {
          if (yp > y2 && ((xp - xn) * (y2 - yn) - (yp - yn) * (x2 - xn)) > 0) wn++;
        } else {
          if (yp <= y2 && ((xp - xn) * (y2 - yn) - (yp - yn) * (x2 - xn)) < 0) wn--;
        }
      }
      for (int o = 16; o > 0; o >>= 1) wn += __shfl_xor_sync(0xffffffffu, wn, o);
      if (wn == 0) {
        if (lane == 0) { A.outside[q] = 1; A.psiRAM[q] = 0.0; }
        continue;
      }
      if (lane == 0) A.outside[q] = 0;
    } else {
      if (A.outside[q] != 0) {
        if (lane == 0) {
          const size_t o = k + (size_t)A.nthe * q;
          A.xRAM[o] = 0.0; A.yRAM[o] = 0.0; A.zRAM[o] = 0.0; A.bRAM[o] = 0.0;
        }
        continue;
      }
      x2 = A.psiRAM[q]; y2 = A.alphaRAM[j];
    }
    int near[9];
    nn9_select(cx, cy, M, x2, y2, lane, near);
    double w[9], wsum = 0.0;                                             // NN_Interpolation_2D
    for (int i = 0; i < 9; i++) {
      const double dx = cx[near[i]] - x2, dy = cy[near[i]] - y2;
      const double d = sqrt(dx * dx + dy * dy);
      if (fabs(d) <= 1e-9) {
        for (int c = 0; c < 9; c++) w[c] = 0.0;
        w[i] = 1.0; wsum = 1.0;
        break;
      }
      w[i] = 1 / (d * d);
      wsum = wsum + w[i];
    }
    const int nf = (MODE == 0) ? 1 : 4;
    if (lane < nf) {
      const double* f = (MODE == 0) ? A.psi : (lane == 0 ? A.x : lane == 1 ? A.y : lane == 2 ? A.z : A.bf);
      double v = 0.0;
      for (int i = 0; i < 9; i++) {
        const size_t o = k + sj * (near[i] / m1) + sk * (near[i] % m1 + 1);
        v = v + f[o] * w[i] / wsum;
      }
      if (MODE == 0) A.psiRAM[q] = v;
      else {
        double* dst = lane == 0 ? A.xRAM : lane == 1 ? A.yRAM : lane == 2 ? A.zRAM : A.bRAM;
        dst[k + (size_t)A.nthe * q] = v;
      }
    }
  }
}

// ScaleAt / outsideMGNP as the reference's NameBoundMag = 'SWMF' branch derives them (src/ModRamScb.f90:306-314): per MLT
// the first radial point outside the SCB domain, every outside line flagged.  Thread per MLT.
__global__ void k_hi_scaleat(int nR, int nT, const int* __restrict__ outsideSCB, int* __restrict__ outsideMGNP, int* __restrict__ ScaleAt,
                             int* /*unused*/) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nT) return;
  int sa = 0;
  for (int i = 0; i < nR; i++) {
    const int o = outsideSCB[i + (size_t)nR * j];
    if (o == 1 && sa == 0) sa = i + 1;
    outsideMGNP[i + (size_t)nR * j] = (o == 1) ? 1 : 0;
  }
  ScaleAt[j] = sa;
}
// densityMode "RAIRDEN" (src/ModRamScb.f90:362-371): 10**(polynomial of the geocentric distance), left-to-right like the Fortran
__global__ void k_hi_rairden(size_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                             double* __restrict__ dens) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double r = sqrt(x[t] * x[t] + y[t] * y[t] + z[t] * z[t]);
  const double r2 = r * r;
  dens[t] = pow(10.0, 13.326 - 3.6908 * r + 1.1362 * r2 - 0.16984 * (r2 * r) + 0.009553 * (r2 * r2));
}

// =============================================================================
// FLC_Radius (src/ModRamLoss.f90:176-336): curvature radius of the field lines and the zeta parameters of the field-line-
// curvature scattering model at the SCB equatorial points, then at the RAM equatorial points by the same 9-nearest-
// neighbour rule as above (GSL_Interpolation_2D -> Interpolation_2D_NN_point).  Only the equatorial slice of the
// reference's 3-D work arrays is ever read, so k_flc_curv evaluates the theta stencil i .. i+3 of that slice directly
// (same expressions, same operation order; compiled with -fmad=false).
// =============================================================================
struct FlcArgs {
  int nthe, npsi, nzeta, nR, nT, ie;             // ie: 0-based equatorial theta index
  double bnormal, REarth;
  const double *x, *y, *z;                       // (nthe,npsi,nzeta+1)
  const double *bx, *by, *bz;                    // (nthe,npsi,nzeta)
  double *xe, *ye, *rc, *z1, *z2;                // (npsi, nzeta-1) in the scatter order of the interpolation: s = j*(nzeta-1) + (k-1)
  const double *qx, *qy;                         // query points (nR,nT) (host: cos / sin of libm)
  double *rcEq, *z1Eq, *z2Eq;                    // (nR,nT)
};

struct FlcNode { double bb, vx, vy, vz; };
__device__ __forceinline__ FlcNode flc_node(const FlcArgs& A, size_t o) {
  FlcNode n;
  const double bx = A.bx[o], by = A.by[o], bz = A.bz[o];
  n.bb = sqrt(bx * bx + by * by + bz * bz);
  n.vx = bx / n.bb; n.vy = by / n.bb; n.vz = bz / n.bb;
  return n;
}
// ax, ay, az and ds at theta node i (1-based 2 .. nthe-1) of line (j, k): |a| -> r_curv
__device__ __forceinline__ void flc_accel(const FlcArgs& A, int i, size_t line, double* rcurv, double* ds_out, double* bb_i, double* bb_ip1) {
  const size_t o = (size_t)(i - 1) + line, o1 = o + 1;
  const FlcNode a = flc_node(A, o), b = flc_node(A, o1);
  const double dx = A.x[o1] - A.x[o], dy = A.y[o1] - A.y[o], dz = A.z[o1] - A.z[o];
  const double dvx = b.vx - a.vx, dvy = b.vy - a.vy, dvz = b.vz - a.vz;
  double ax, ay, az;
  if (dx == 0.0) {
    ax = a.vy * (dvx / dy) + a.vz * (dvx / dz);
    ay = a.vy * (dvy / dy) + a.vz * (dvy / dz);
    az = a.vy * (dvz / dy) + a.vz * (dvz / dz);
  } else if (dy == 0.0) {
    ax = a.vx * (dvx / dx) + a.vz * (dvx / dz);
    ay = a.vx * (dvy / dx) + a.vz * (dvy / dz);
    az = a.vx * (dvz / dx) + a.vz * (dvz / dz);
  } else {
    ax = a.vx * dvx / dx + a.vy * dvx / dy + a.vz * dvx / dz;
    ay = a.vx * dvy / dx + a.vy * dvy / dy + a.vz * dvy / dz;
    az = a.vx * dvz / dx + a.vy * dvz / dy + a.vz * dvz / dz;
  }
  *rcurv = 1. / sqrt(ax * ax + ay * ay + az * az) * A.REarth;
  *ds_out = sqrt(dx * dx + dy * dy + dz * dz) * A.REarth;
  *bb_i = a.bb; *bb_ip1 = b.bb;
}
// r_curv(i) with the end rules r_curv(1) = r_curv(2), r_curv(nthe) = r_curv(nthe-1)   (:276-278)
__device__ __forceinline__ double flc_rcurv(const FlcArgs& A, int i, size_t line) {
  const int ii = (i < 2) ? 2 : ((i > A.nthe - 1) ? A.nthe - 1 : i);
  double rc, ds, b0, b1;
  flc_accel(A, ii, line, &rc, &ds, &b0, &b1);
  return rc;
}
// thread per equatorial point (j, k), k = 2..nzeta (Fortran)
__global__ void k_flc_curv(FlcArgs A) {
  const int m1 = A.nzeta - 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.npsi * m1) return;
  const int j = t / m1, kk = t - j * m1;                   // kk = k - 2
  const size_t line = (size_t)A.nthe * (j + (size_t)A.npsi * (kk + 1));
  const int i = A.ie + 1;                                  // 1-based theta index of the equator; needs 2 <= i <= nthe-3
  double rc[3], ds[2], b0[2], b1[2];
  for (int q = 0; q < 2; ++q) flc_accel(A, i + q, line, &rc[q], &ds[q], &b0[q], &b1[q]);
  rc[0] = flc_rcurv(A, i, line); rc[1] = flc_rcurv(A, i + 1, line); rc[2] = flc_rcurv(A, i + 2, line);
  // dBdS, dRcdS at i and i+1 (:280-283), second differences at i (:288-291)
  const double dBdS0 = A.bnormal * (b1[0] - b0[0]) * 1.0e-9 / ds[0], dBdS1 = A.bnormal * (b1[1] - b0[1]) * 1.0e-9 / ds[1];
  const double dRdS0 = (rc[1] - rc[0]) / ds[0], dRdS1 = (rc[2] - rc[1]) / ds[1];
  const double d2B = (dBdS1 - dBdS0) / ds[0], d2R = (dRdS1 - dRdS0) / ds[0];
  const size_t o = (size_t)(i - 1) + line;
  A.xe[t] = A.x[o]; A.ye[t] = A.y[o];
  A.rc[t] = rc[0];
  A.z1[t] = rc[0] * d2R;
  A.z2[t] = rc[0] * rc[0] / (A.bnormal * 1.0e-9 * b0[0]) * d2B;
}

// 9-nearest-neighbour interpolation of rc, z1, z2 to the RAM points: a warp per query, candidates in shared memory
// (selection = the single-scan form of k_hi_nn9: nine smallest (distance, index) pairs per lane, nine shuffle elections)
__global__ void __launch_bounds__(256) k_flc_nn9(FlcArgs A) {
  extern __shared__ double hi_sm[];
  const int m1 = A.nzeta - 1, M = A.npsi * m1;
  double* cx = hi_sm;
  double* cy = cx + M;
  for (int s = threadIdx.x; s < M; s += blockDim.x) { cx[s] = A.xe[s]; cy[s] = A.ye[s]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nq = (A.nR - 1) * (A.nT - 1);
  for (int q = blockIdx.x * nwarp + warp; q < nq; q += gridDim.x * nwarp) {
    const int i = 1 + q % (A.nR - 1), j = q / (A.nR - 1);                // 0-based (Fortran i = 2..nR, j = 1..nT-1)
    const size_t oq = (size_t)i + (size_t)A.nR * j;
    const double x2 = A.qx[oq], y2 = A.qy[oq];
    int near[9];
    nn9_select(cx, cy, M, x2, y2, lane, near);
    double w[9], wsum = 0.0;                                             // NN_Interpolation_2D (src/ModRamGSL.f90:872-917)
    for (int c = 0; c < 9; c++) {
      const double dx = cx[near[c]] - x2, dy = cy[near[c]] - y2;
      const double d = sqrt(dx * dx + dy * dy);
      if (fabs(d) <= 1e-9) {
        for (int e = 0; e < 9; e++) w[e] = 0.0;
        w[c] = 1.0; wsum = 1.0;
        break;
      }
      w[c] = 1 / (d * d);
      wsum = wsum + w[c];
    }
    if (lane < 3) {
      const double* f = lane == 0 ? A.rc : (lane == 1 ? A.z1 : A.z2);
      double v = 0.0;
      // the oracle's candidate order is i-outer / j-inner over (npsi, nzeta-1) stored (npsi fastest): s = j*m1 + kk here
      for (int c = 0; c < 9; c++) v = v + f[near[c]] * w[c] / wsum;
      (lane == 0 ? A.rcEq : (lane == 1 ? A.z1Eq : A.z2Eq))[oq] = v;
    }
  }
}
// MLT = 24 is MLT = 0, the innermost circle repeats the second one (:311-316); thread per (i, j)
__global__ void k_flc_edges(FlcArgs A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nR * A.nT) return;
  const int i = t % A.nR, j = t / A.nR;
  if (i != 0 && j != A.nT - 1) return;
  const int js = (j == A.nT - 1) ? 0 : j, is = (i == 0) ? 1 : i;
  const size_t src = (size_t)is + (size_t)A.nR * js;
  A.rcEq[t] = A.rcEq[src]; A.z1Eq[t] = A.z1Eq[src]; A.z2Eq[t] = A.z2Eq[src];
}
