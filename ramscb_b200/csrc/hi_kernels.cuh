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
