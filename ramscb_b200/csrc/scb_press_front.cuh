// `pressure`, 2-D front end on the device (src/ModScbRun.f90:838-1086, anisotropic branch with RAM pressures; SURVEY
// 8(f)-2): with it an SCB outer iteration no longer calls back to the host.
//   k_scb_press_raw  once per scb_run: RAM pressures of the species%SCB species summed on the RAM grid, extended
//                    radially (PressMode), smoothed (SavGol7 x SavGolIters, Gaussian 9 x 9)
//   k_scb_press_eq   per `pressure`: equatorial foot points -> (r^2, azimuth), gsl bilinear rule, extap inside 2 RE, floor
//   k_scb_press_wrap periodic columns + normalisation
// Reference operation order throughout (summations sequential like DOT_PRODUCT / SUM in array element order); the
// transcendental pieces that are pure functions of the grids (SKD / ROE profile factors, Gaussian weights) come from
// the host as small tables.
#pragma once
#include <cuda_runtime.h>

struct PressRawArgs {
  int nS, NR, NT, nX, nXRaw, nAz, mode, iSm2, iters;
  const double *PPerT, *PParT;   // (nS,NR,NT)
  const int* scb;                // [nS]
  const double* radExt;          // [nX]
  const double* fac;             // [nX] profile of the radial extension (SKD: 89 exp(-0.59 r) + 8.9 r^-1.53; ROE: pRoeRad)
  const double* w;               // [81] Gaussian weights
  double *per, *par;             // out (nX,nAz)
  double *t0, *t1, *t2;          // scratch, 2 * nX * nAz each (per | par)
};

__device__ __forceinline__ double sg_b(int r, int c) {            // BSav(r,c), RESHAPE by columns (src/ModScbFunctions.f90:157-163)
  const double vals[49] = {32, 5, 1, -2, -2, -1, 5, 15, 4, 3, 3, 1, 0, -3, 3, 3, 4, 6, 3, 1, -6, -4, 2, 4, 7, 4, 2, -4,
                           -6, 1, 3, 6, 4, 3, 3, -3, 0, 1, 3, 3, 4, 15, 5, -1, -2, -2, 1, 5, 32};
  return vals[(c - 1) * 7 + (r - 1)];
}

// one CTA; `n` = nX * nAz points of each of the two arrays (a = 0 per, 1 par)
__global__ void __launch_bounds__(256) k_scb_press_raw(PressRawArgs A) {
  __shared__ double smin[2];
  const int nX = A.nX, nAz = A.nAz, n = nX * nAz, tid = threadIdx.x, T = blockDim.x;
  double* cur = A.t0;
#define PIX(buf, a, j, k) (buf)[(size_t)(a) * n + (size_t)((j)-1) + (size_t)nX * ((k)-1)]
  // RAM pressures summed over the SCB species, keV/cm^3 -> nPa (:858-885)
  for (int q = tid; q < 2 * n; q += T) cur[q] = 0.0;
  __syncthreads();
  for (int q = tid; q < A.nXRaw * nAz; q += T) {
    const int j1 = 1 + q % A.nXRaw, k1 = 1 + q / A.nXRaw;
    double sp = 0.0, sa = 0.0;
    for (int iS = 0; iS < A.nS; ++iS)
      if (A.scb[iS]) {
        sp = sp + A.PPerT[(size_t)iS + (size_t)A.nS * (j1 + (size_t)A.NR * (k1 - 1))];
        sa = sa + A.PParT[(size_t)iS + (size_t)A.nS * (j1 + (size_t)A.NR * (k1 - 1))];
      }
    PIX(cur, 0, j1, k1) = 0.16 * sp;
    PIX(cur, 1, j1, k1) = 0.16 * sa;
  }
  __syncthreads();
  // radial extension (:884-938); EXT is a recurrence in j1: one thread per (array, azimuth) column
  for (int q = tid; q < 2 * nAz; q += T) {
    const int a = q / nAz, k1 = 1 + q % nAz, nXRaw = A.nXRaw;
    const double* R = A.radExt - 1;
    if (A.mode == 0) {
      for (int j1 = nXRaw - 1; j1 <= nX; ++j1) PIX(cur, a, j1, k1) = PIX(cur, a, nXRaw - 2, k1) * A.fac[j1 - 1] / A.fac[nXRaw - 3];
    } else if (A.mode == 1) {
      for (int j1 = nXRaw + 1; j1 <= nX; ++j1) PIX(cur, a, j1, k1) = PIX(cur, a, nXRaw, k1) * A.fac[j1 - 1] / A.fac[nXRaw - 1];
    } else if (A.mode == 2) {
      for (int j1 = nXRaw + 1; j1 <= nX; ++j1)
        PIX(cur, a, j1, k1) = PIX(cur, a, j1 - 1, k1) + (R[j1] - R[j1 - 1]) / (R[j1 - 2] - R[j1 - 1]) * (PIX(cur, a, j1 - 2, k1) - PIX(cur, a, j1 - 1, k1));
    } else {
      for (int j1 = nXRaw + 1; j1 <= nX - 1; ++j1) PIX(cur, a, j1, k1) = PIX(cur, a, nXRaw, k1);
      PIX(cur, a, nX, k1) = 0.0;
    }
  }
  __syncthreads();
  if (A.iSm2 == 1 || A.iSm2 == 4) {                         // SavGol7 (src/ModScbFunctions.f90:137-230)
    if (tid < 2) {                                          // MINVAL(pres) of the input
      double mn = cur[(size_t)tid * n];
      for (int q = 1; q < n; ++q) { const double v = cur[(size_t)tid * n + q]; mn = v < mn ? v : mn; }
      smin[tid] = mn;
    }
    double *p0 = cur, *p1 = A.t1, *p2 = A.t2;
    for (int it = 0; it < A.iters; ++it) {
      for (int q = tid; q < 2 * n; q += T) {                // first pass, along j
        const int a = q / n, r = q - a * n, j = 1 + r % nX, k = 1 + r / nX;
        double v;
        if (j > 3 && j < nX - 2) {
          v = 0.0;
          for (int m = 1; m <= 7; ++m) v += (sg_b(4, m) / 21.) * PIX(p0, a, j - 4 + m, k);
        } else if (j <= 3) v = PIX(p0, a, j, k);
        else {
          const int row = (j == nX - 2) ? 5 : ((j == nX - 1) ? 6 : 7);
          const double den = (row == 7) ? 42. : 14.;
          v = 0.0;
          for (int m = 1; m <= 7; ++m) v += (sg_b(row, m) / den) * PIX(p0, a, nX - 7 + m, k);
        }
        PIX(p1, a, j, k) = v;
      }
      __syncthreads();
      for (int q = tid; q < 2 * n; q += T) {                // second pass, along k (periodic, J = 1 and J = NT coincide)
        const int a = q / n, r = q - a * n, j = 1 + r % nX, k = 1 + r / nX;
        int idx[7];
        if (k > 3 && k < nAz - 2) { for (int m = 0; m < 7; ++m) idx[m] = k - 3 + m; }
        else if (k == 1 || k == nAz) { idx[0] = nAz - 3; idx[1] = nAz - 2; idx[2] = nAz - 1; idx[3] = 1; idx[4] = 2; idx[5] = 3; idx[6] = 4; }
        else if (k == 2) { idx[0] = nAz - 2; idx[1] = nAz - 1; idx[2] = 1; idx[3] = 2; idx[4] = 3; idx[5] = 4; idx[6] = 5; }
        else if (k == nAz - 1) { idx[0] = nAz - 4; idx[1] = nAz - 3; idx[2] = nAz - 2; idx[3] = nAz - 1; idx[4] = 1; idx[5] = 2; idx[6] = 3; }
        else if (k == 3) { idx[0] = nAz - 1; idx[1] = 1; idx[2] = 2; idx[3] = 3; idx[4] = 4; idx[5] = 5; idx[6] = 6; }
        else { idx[0] = nAz - 5; idx[1] = nAz - 4; idx[2] = nAz - 3; idx[3] = nAz - 2; idx[4] = nAz - 1; idx[5] = 1; idx[6] = 2; }
        double v = 0.0;
        for (int m = 1; m <= 7; ++m) v += (sg_b(4, m) / 21.) * PIX(p1, a, j, idx[m - 1]);
        PIX(p2, a, j, k) = v;
      }
      __syncthreads();
      double* sw = p0; p0 = p2; p2 = sw;                    // pres0 = pres2
    }
    for (int q = tid; q < 2 * n; q += T) { const double v = p0[q]; p0[q] = (v < 0) ? smin[q / n] : v; }
    __syncthreads();
    cur = p0;
  }
  double* dst = (cur == A.t1) ? A.t2 : A.t1;
  if (A.iSm2 == 3 || A.iSm2 == 4) {                         // gaussian_kernel(1.0) + convolve (srcExternal/gaussian_filter.f90)
    for (int q = tid; q < 2 * n; q += T) {
      const int a = q / n, r = q - a * n, i = r % nX, j = r / nX;
      double sum = 0.0;
      for (int dj = -4; dj <= 4; ++dj) {
        int jj = j + dj;
        jj = jj < 0 ? -1 - jj : (jj >= nAz ? 2 * nAz - 1 - jj : jj);
        for (int di = -4; di <= 4; ++di) {
          int ir = i + di;
          ir = ir < 0 ? -1 - ir : (ir >= nX ? 2 * nX - 1 - ir : ir);
          sum += A.w[(di + 4) + 9 * (dj + 4)] * cur[(size_t)a * n + ir + (size_t)nX * jj];
        }
      }
      dst[q] = sum;
    }
    __syncthreads();
    cur = dst;
  }
  for (int q = tid; q < n; q += T) { A.per[q] = cur[q]; A.par[q] = cur[n + q]; }
#undef PIX
}

__device__ __forceinline__ int pf_bsearch(const double* xa, double x, int n) {
  int ilo = 0, ihi = n - 1;
  while (ihi > ilo + 1) {
    const int i = (ihi + ilo) / 2;
    if (xa[i] > x) ihi = i; else ilo = i;
  }
  return ilo;
}
__device__ __forceinline__ double pf_bilinear(int n1, int m1, const double* xa, const double* ya, const double* za, double x, double y) {
  const int xi = pf_bsearch(xa, x, n1), yi = pf_bsearch(ya, y, m1);
  const double xmin = xa[xi], xmax = xa[xi + 1], ymin = ya[yi], ymax = ya[yi + 1];
  const double zminmin = za[(size_t)yi * n1 + xi], zminmax = za[(size_t)(yi + 1) * n1 + xi];
  const double zmaxmin = za[(size_t)yi * n1 + xi + 1], zmaxmax = za[(size_t)(yi + 1) * n1 + xi + 1];
  const double dx = xmax - xmin, dy = ymax - ymin;
  const double t = (x - xmin) / dx, u = (y - ymin) / dy;
  return (1. - t) * (1. - u) * zminmin + t * (1. - u) * zmaxmin + (1. - t) * u * zminmax + t * u * zmaxmax;
}
__device__ __forceinline__ void pf_extap(double x1, double x2, double x3, double& x4) {     // src/ModScbFunctions.f90:57-76
  x4 = 3. * x3 - 3. * x2 + x1;
  const double ddx1 = x3 - x2, ddx2 = x2 - x1;
  double ddx = x4 - x3;
  const double pm = ddx * ddx1;
  if (pm > 0.) return;
  if (fabs(ddx2) <= 1e-9) { x4 = 2. * x3 - x2; return; }
  ddx = (ddx1 * ddx1) / ddx2;
  x4 = x3 + ddx;
}

// thread per zeta column k (0-based 0..nzeta): interpolation for k = 1..nzeta-1, the extap repair, the floor.
// eq = xEq | yEq (npsi, nzeta+1) from k_scb_gather_eq; out = pperEq | pparEq (npsi, nzeta+1), un-normalised
__global__ void k_scb_press_eq(int npsi, int nzeta, const double* __restrict__ eq, int nX, int nAz, const double* __restrict__ rad2,
                               const double* __restrict__ azim, const double* __restrict__ per, const double* __restrict__ par,
                               double pnormal, double* __restrict__ out) {
  const double PI_D = 3.141592653589793238462643383279502884197;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nzeta) return;
  const size_t n2 = (size_t)npsi * (nzeta + 1);
  double* pe = out + (size_t)npsi * k;
  double* pa = out + n2 + (size_t)npsi * k;
  const bool inner = (k >= 1 && k <= nzeta - 1);
  for (int j = 0; j < npsi; ++j) {
    double ve = 0.0, va = 0.0;
    if (inner) {
      const double xe = eq[(size_t)npsi * k + j], ye = eq[n2 + (size_t)npsi * k + j];
      const double radius = sqrt(xe * xe + ye * ye);
      double angle = asin(ye / radius) + PI_D;
      if (xe <= 0 && ye >= 0) angle = 2.0 * PI_D - asin(ye / radius);
      if (xe <= 0 && ye <= 0) angle = -asin(ye / radius);
      ve = pf_bilinear(nX, nAz, rad2, azim, per, radius * radius, angle);
      va = pf_bilinear(nX, nAz, rad2, azim, par, radius * radius, angle);
    }
    pe[j] = ve;
    pa[j] = va;
  }
  if (k <= nzeta - 1) {                                      // Fortran k = 1..nzeta; radGrid(:,1) is 0 (< 2)
    for (int j = 9; j >= 0; --j) {
      double radius = 0.0;
      if (inner) {
        const double xe = eq[(size_t)npsi * k + j], ye = eq[n2 + (size_t)npsi * k + j];
        radius = sqrt(xe * xe + ye * ye);
      }
      if (radius < 2.0) {
        pf_extap(pe[j + 3], pe[j + 2], pe[j + 1], pe[j]);
        pf_extap(pa[j + 3], pa[j + 2], pa[j + 1], pa[j]);
      }
    }
  }
  for (int j = 0; j < npsi; ++j) {
    if (pe[j] <= 0.0) pe[j] = 1e-1 / pnormal;
    if (pa[j] <= 0.0) pa[j] = 1e-1 / pnormal;
  }
}
// periodic columns (:1079-1082), then the normalisation (:1085-1086)
__global__ void k_scb_press_wrap(int npsi, int nzeta, double pnormal, double* __restrict__ out) {
  const size_t n2 = (size_t)npsi * (nzeta + 1);
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 2 * n2) return;
  const size_t a = q / n2, r = q - a * n2;
  const int k = (int)(r / npsi), j = (int)(r - (size_t)k * npsi);
  const int ks = (k == nzeta) ? 1 : ((k == 0) ? nzeta - 1 : k);
  out[2 * n2 + q] = out[a * n2 + (size_t)npsi * ks + j] / pnormal;        // second half of the buffer: the wrapped, normalised pair
}
