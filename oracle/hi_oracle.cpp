// TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by ramscb_b200/).
//
// Restatement of the integral block of computehI, /root/reference/src/ModRamScb.f90:372-410, with
// GSL_Integration_hI / GSL_BounceAverage (src/ModRamGSL.f90:125-200) and their C drivers
// integrator_c (src/RamGSL.c:535-602) and bounceaverage_c (:479-533), in the reference's own
// structure: mirror points for every pitch angle first, then the serial loop L = nPa-2 .. 1.
//
// PARITY UNPINNED at the bit level: the reference integrates f_I = sqrt(Bm-B), f_h = 1/sqrt(Bm-B),
// f_D = n/sqrt(Bm-B) (src/RamGSL.c:326-448; B, n by gsl_interp_linear, :295-322) with
// gsl_integration_cquad(epsabs = epsrel = 1e-3) from GNU GSL (2.5/2.6, NOT under /root/reference).
// This oracle integrates the same piecewise-linear integrands in closed form per grid segment, i.e.
// it returns the limit cquad converges to; tests/test_cpu.py::test_hi_oracle_vs_adaptive_quadrature
// checks it against an independent adaptive quadrature (QUADPACK) of the literal integrands at
// cquad's tolerance, and test_hi_integrals_reproduce_reference_dipole_functions pins H_cart / I_cart on dipole
// lines to the reference's own closed forms funt / funi (src/ModRamFunctions.f90:90-143) to 1e-3.  Layouts are the reference's (Fortran order, theta fastest).
#include <cmath>
#include <vector>

namespace {

struct Seg3 { double I, H, V; };

// int over one segment of width h; u = Bm - B linear u0 -> u1, v linear v0 -> v1; integrands are 0 where u <= 0
Seg3 segment(double h, double u0, double u1, double v0, double v1) {
  Seg3 r = {0.0, 0.0, 0.0};
  if (u0 <= 0.0 && u1 <= 0.0) return r;
  double hp = h, ua = u0, ub = u1, va = v0, vb = v1;
  if (u1 <= 0.0) {
    double t = u0 / (u0 - u1);
    hp = h * t; ub = 0.0; vb = v0 + (v1 - v0) * t;
  } else if (u0 <= 0.0) {
    double t = u0 / (u0 - u1);
    hp = h * (1.0 - t); ua = 0.0; va = v0 + (v1 - v0) * t;
  }
  double a = std::sqrt(ua), b = std::sqrt(ub), s = a + b;
  r.I = (2.0 * hp / 3.0) * ((ua + a * b + ub) / s);
  r.H = 2.0 * hp / s;
  r.V = hp * (va * (2.0 / s) + (vb - va) * ((2.0 / 3.0) * (2.0 * a + b) / (s * s)));
  return r;
}

// the three integrals between grid nodes k0 and k1 (a = cVal[k0], b = cVal[k1])
Seg3 integrate(int k0, int k1, const double* cVal, const double* bf, const double* var, double mirror) {
  Seg3 t = {0.0, 0.0, 0.0};
  for (int k = k0; k < k1; k++) {
    Seg3 s = segment(cVal[k + 1] - cVal[k], mirror - bf[k], mirror - bf[k + 1], var[k], var[k + 1]);
    t.I += s.I; t.H += s.H; t.V += s.V;
  }
  return t;
}

void mirror_points(int nT, int nPa, const double* mirror, const double* cVal, const double* bf, std::vector<double>& a,
                   std::vector<double>& b, std::vector<int>& LH, std::vector<int>& RH) {
  for (int L = 1; L < nPa - 1; L++) {
    a[L] = 0; b[L] = 0; LH[L] = 0; RH[L] = 0;
    for (int i = 1; i < nT - 1; i++)
      if (mirror[L] <= bf[i - 1] && mirror[L] >= bf[i]) { a[L] = cVal[i - 1]; LH[L] = i - 1; break; }
    for (int i = nT - 2; i > 0; i--)
      if (mirror[L] >= bf[i - 1] && mirror[L] <= bf[i]) { b[L] = cVal[i]; RH[L] = i; break; }
  }
}

// integrator_c, src/RamGSL.c:535-602
void integrator(int nT, int nPa, double* mirror, const double* cVal, const double* bf, const double* var, double* yI, double* yH) {
  std::vector<double> a(nPa), b(nPa);
  std::vector<int> LH(nPa), RH(nPa);
  mirror_points(nT, nPa, mirror, cVal, bf, a, b, LH, RH);
  Seg3 base = integrate(0, nT - 1, cVal, bf, var, mirror[nPa - 1]);
  yI[nPa - 1] = base.I; yH[nPa - 1] = base.H;
  for (int L = nPa - 2; L > 0; L--) {
    if (mirror[L] >= bf[1] || a[L] == 0) a[L] = cVal[0];
    if (mirror[L] >= bf[nT - 1] || b[L] == 0) b[L] = cVal[nT - 1];
    if (a[L] <= cVal[0] || b[L] >= cVal[nT - 1]) {
      mirror[L] = mirror[nPa - 1];
      yI[L] = yI[nPa - 1];
      yH[L] = yH[nPa - 1];
    } else {
      if ((RH[L] - LH[L]) <= 4) { yI[L] = yI[L + 1]; yH[L] = yH[L + 1]; continue; }
      Seg3 t = integrate(LH[L], RH[L], cVal, bf, var, mirror[L]);
      yI[L] = t.I; yH[L] = t.H;
      if (yI[L] <= 0) yI[L] = yI[L + 1];
      if (yH[L] <= 0) yH[L] = yH[L + 1];
    }
  }
  yI[0] = 0;
  yH[0] = yH[1];
}

// bounceaverage_c, src/RamGSL.c:479-533
void bounceaverage(int nT, int nPa, double* mirror, const double* cVal, const double* bf, const double* var, double* yV) {
  std::vector<double> a(nPa), b(nPa);
  std::vector<int> LH(nPa), RH(nPa);
  mirror_points(nT, nPa, mirror, cVal, bf, a, b, LH, RH);
  yV[nPa - 1] = integrate(0, nT - 1, cVal, bf, var, mirror[nPa - 1]).V;
  for (int L = nPa - 2; L > 0; L--) {
    if (mirror[L] >= bf[1] || a[L] == 0) a[L] = cVal[0];
    if (mirror[L] >= bf[nT - 1] || b[L] == 0) b[L] = cVal[nT - 1];
    if (a[L] <= cVal[0] || b[L] >= cVal[nT - 1]) {
      mirror[L] = mirror[nPa - 1];
      yV[L] = yV[nPa - 1];
    } else {
      if ((RH[L] - LH[L]) <= 4) { yV[L] = yV[L + 1]; continue; }
      yV[L] = integrate(LH[L], RH[L], cVal, bf, var, mirror[L]).V;
      if (yV[L] <= 0) yV[L] = yV[L + 1];
    }
  }
  yV[0] = yV[1];
}

}  // namespace

extern "C" {

// one line: the C drivers alone (used by the quadrature cross-check); mirror is INOUT like the reference's bM
void hio_line(int nT, int nPa, double* mirror, const double* cVal, const double* bf, const double* var, double* yI, double* yH,
              double* yV) {
  integrator(nT, nPa, mirror, cVal, bf, var, yI, yH);
  bounceaverage(nT, nPa, mirror, cVal, bf, var, yV);
}

// the loop nest of src/ModRamScb.f90:378-410 (0-based i, j here; Fortran-order arrays)
void hio_integrals(int nthe, int nR, int nT, int nPa, int nThetaEquator, double bnormal, const double* chiVal, const double* mu,
                   const double* xRAM, const double* yRAM, const double* zRAM, const double* bRAM_in, const double* density,
                   const int* outsideMGNP, double* I_cart, double* H_cart, double* HDens_cart, double* bZEq_cart,
                   double* bfMirror_out) {
  const double pi_d = 3.141592653589793238462643383279502884197;
  const size_t nl = (size_t)nR * nT;
  const int ke = nThetaEquator - 1;
  std::vector<double> bR(nthe), mir(nPa), yI(nPa), yH(nPa), yD(nPa);
  for (int i = 0; i < nR; i++)
    for (int j = 0; j < nT; j++) {
      const size_t line = i + (size_t)nR * j, o = line * nthe;
      if (outsideMGNP[line] != 0) {
        for (int L = 0; L < nPa; L++) { I_cart[line + nl * L] = 0.0; H_cart[line + nl * L] = 0.0; }   // :236-237
        bZEq_cart[line] = 0.0;
        continue;
      }
      double length = 0.0;
      for (int k = 1; k < nthe; k++) {
        double dx = xRAM[o + k] - xRAM[o + k - 1], dy = yRAM[o + k] - yRAM[o + k - 1], dz = zRAM[o + k] - zRAM[o + k - 1];
        length = length + std::sqrt(dx * dx + dy * dy + dz * dz);
      }
      double r0 = std::sqrt(xRAM[o + ke] * xRAM[o + ke] + yRAM[o + ke] * yRAM[o + ke]);
      double bmin = bRAM_in[o];
      for (int k = 0; k < nthe; k++) { bR[k] = bRAM_in[o + k]; if (bR[k] < bmin) bmin = bR[k]; }
      if (std::fabs(bR[ke] - bmin) > 1e-9) {
        if (2.0 * bmin - bR[ke] > 0.0) bR[ke] = 2.0 * bmin - bR[ke];
        else bR[ke] = bmin - 0.01;
      }
      for (int L = 0; L < nPa - 1; L++) mir[L] = bR[ke] / (1.0 - mu[L] * mu[L]);
      mir[nPa - 1] = bR[nthe - 1];
      integrator(nthe, nPa, mir.data(), chiVal, bR.data(), density + o, yI.data(), yH.data());
      bounceaverage(nthe, nPa, mir.data(), chiVal, bR.data(), density + o, yD.data());
      for (int L = 0; L < nPa; L++) {
        I_cart[line + nl * L] = (length / (pi_d * r0)) * yI[L] / std::sqrt(mir[L]);
        H_cart[line + nl * L] = (length / (pi_d * 2 * r0)) * yH[L] * std::sqrt(mir[L]);
        HDens_cart[line + nl * L] = yD[L] / yH[L];
        if (bfMirror_out) bfMirror_out[line + nl * L] = mir[L];
      }
      bZEq_cart[line] = bR[ke] * bnormal;
    }
}

}  // extern "C"

// ---- computehI, "Convert SCB field lines to RAM field lines" (src/ModRamScb.f90:252-300) -------------------------
// winding-number test against the ring (nThetaEquator, npsi-1, :), psiRAM by GSL_Interpolation_2D on the scattered
// equatorial points, then x, y, z, bf of every node k of the line by GSL_Interpolation_2D in (psi, alfa) space.  The
// generic resolves to Interpolation_2D_NN_point (src/ModRamGSL.f90:368-422): the 9 nearest scattered points by nine
// MINLOC passes, then NN_Interpolation_2D (:872-917), inverse-distance-squared weights.  Pure Fortran in the
// reference (no GSL call), restated literally.  Arrays (nthe,npsi,nzeta+1), Fortran order.
namespace {
double nn9(int n1, int m1, const double* x1, const double* y1, size_t s1, size_t s2, const double* const* f1, int nf, double x2,
           double y2, std::vector<double>& distance, double* f2) {
  // x1(i,j) = x1[i*s1 + j*s2]; scatter order of :391-399: i outer, j inner
  const int nTotal = n1 * m1;
  int it = 0;
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < m1; j++, it++) {
      const double dx = x1[i * s1 + j * s2] - x2, dy = y1[i * s1 + j * s2] - y2;
      distance[it] = dx * dx + dy * dy;
    }
  size_t near[9];
  for (int k = 0; k < 9; k++) {
    int best = 0;
    for (int q = 1; q < nTotal; q++)
      if (distance[q] < distance[best]) best = q;      // MINLOC: first minimum
    near[k] = (size_t)(best / m1) * s1 + (size_t)(best % m1) * s2;
    distance[best] = 999999.9;
  }
  double w[9], wsum = 0.0;                              // NN_Interpolation_2D
  for (int i = 0; i < 9; i++) {
    const double dx = x1[near[i]] - x2, dy = y1[near[i]] - y2;
    const double d = std::sqrt(dx * dx + dy * dy);
    if (std::fabs(d) <= 1e-9) {
      for (int q = 0; q < 9; q++) w[q] = 0.0;
      w[i] = 1.0; wsum = 1.0;
      break;
    }
    w[i] = 1 / (d * d);
    wsum = wsum + w[i];
  }
  for (int c = 0; c < nf; c++) {
    double v = 0.0;
    for (int i = 0; i < 9; i++) v = v + f1[c][near[i]] * w[i] / wsum;
    f2[c] = v;
  }
  return wsum;
}
}  // namespace

extern "C" void hio_convert_lines(int nthe, int npsi, int nzeta, int nR, int nT, int nThetaEquator, const double* x, const double* y,
                                  const double* z, const double* bf, const double* psi, const double* alfa, const double* Lz,
                                  const double* MLT, double* xRAM, double* yRAM, double* zRAM, double* bRAM, int* outsideSCB,
                                  double* psiRAM_out) {
  const double pi_d = 3.141592653589793238462643383279502884197, twopi_d = 2.0 * pi_d;
  const size_t sj = nthe, sk = (size_t)nthe * npsi;
  const int ke = nThetaEquator - 1;
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int i = 0; i < nR; i++)
    for (int j = 0; j < nT; j++) {
      std::vector<double> distance((size_t)npsi * (nzeta - 1));
      const size_t line = i + (size_t)nR * j;
      for (int k = 0; k < nthe; k++) { xRAM[k + nthe * line] = 0; yRAM[k + nthe * line] = 0; zRAM[k + nthe * line] = 0; bRAM[k + nthe * line] = 0; }
      const double xo = Lz[i + 1] * std::cos(MLT[j] * 2.0 * pi_d / 24.0 - pi_d);
      const double yo = Lz[i + 1] * std::sin(MLT[j] * 2.0 * pi_d / 24.0 - pi_d);
      int wn = 0;
      for (int k = 0; k < nzeta; k++) {                                  // ring (nThetaEquator, npsi-1, k), k = 1..nzeta
        const size_t o = ke + sj * (npsi - 2) + sk * k;
        const double yn = y[o], yp = y[o + sk], xn = x[o], xp = x[o + sk];
        if (yn <= yo) {
          if (yp > yo)
            if (((xp - xn) * (yo - yn) - (yp - yn) * (xo - xn)) > 0) wn = wn + 1;
        } else {
          if (yp <= yo)
            if (((xp - xn) * (yo - yn) - (yp - yn) * (xo - xn)) < 0) wn = wn - 1;
        }
      }
      outsideSCB[line] = 0;
      if (psiRAM_out) psiRAM_out[line] = 0.0;
      if (std::abs(wn) > 0) {
        double alphaRAM = MLT[j] * pi_d / 12.0 + pi_d;
        if (alphaRAM > twopi_d) alphaRAM = alphaRAM - twopi_d;
        double psiRAM;
        const double* fp[1] = {psi + ke + sk};                         // (nThetaEquator, :, 2:nzeta)
        nn9(npsi, nzeta - 1, x + ke + sk, y + ke + sk, sj, sk, fp, 1, xo, yo, distance, &psiRAM);
        if (psiRAM_out) psiRAM_out[line] = psiRAM;
        for (int k = 0; k < nthe; k++) {
          const double* f4[4] = {x + k + sk, y + k + sk, z + k + sk, bf + k + sk};
          double r[4];
          nn9(npsi, nzeta - 1, psi + k + sk, alfa + k + sk, sj, sk, f4, 4, psiRAM, alphaRAM, distance, r);
          xRAM[k + nthe * line] = r[0]; yRAM[k + nthe * line] = r[1]; zRAM[k + nthe * line] = r[2]; bRAM[k + nthe * line] = r[3];
        }
      } else {
        outsideSCB[line] = 1;
      }
    }
}


// ---- FLC_Radius (src/ModRamLoss.f90:176-336): field-line curvature radius and the zeta parameters of the FLC scattering
// model on the SCB grid, then their values at the RAM equatorial points by GSL_Interpolation_2D = the 9-nearest-neighbour
// inverse-distance rule above (Interpolation_2D_NN_point).  x, y, z (nthe,npsi,nzeta+1); bx, by, bz (nthe,npsi,nzeta) of
// computeBandJacob; radRaw(1:nR) and azimRaw(1:nT) as the reference holds them; outputs (nR,nT).  Only the equatorial
// slice of the 3-D fields is read at the end; the slots the reference leaves unassigned (theta ends of the second
// differences) are never touched by it and are left at 0 here.  The caller keeps the "every Dt_bc" gate (:207).
extern "C" void hio_flc_radius(int nthe, int npsi, int nzeta, int nR, int nT, int nThetaEquator, double bnormal, double REarth,
                               const double* x, const double* y, const double* z, const double* bx, const double* by, const double* bz,
                               const double* radRaw, const double* azimRaw, double* r_curvEq, double* zeta1Eq, double* zeta2Eq) {
  const double PI = 3.1415926535897932384626433832795;      // ModRamConst PI = cPi
  const size_t sj = nthe, sk3 = (size_t)nthe * npsi;          // strides of both (…,nzeta) and (…,nzeta+1) arrays
  const size_t n3 = sk3 * nzeta;
  auto X3 = [&](const double* a, int i, int j, int k) { return a[(size_t)(i - 1) + sj * (j - 1) + sk3 * (k - 1)]; };
  std::vector<double> bb(n3), vbx(n3), vby(n3), vbz(n3), ax(n3, 0.0), ay(n3, 0.0), az(n3, 0.0), ds(n3, 0.0), rc(n3, 0.0), dBdS(n3, 0.0),
      dRcdS(n3, 0.0), d2B(n3, 0.0), d2R(n3, 0.0);
  auto W3 = [&](std::vector<double>& a, int i, int j, int k) -> double& { return a[(size_t)(i - 1) + sj * (j - 1) + sk3 * (k - 1)]; };
  for (size_t q = 0; q < n3; ++q) {
    bb[q] = std::sqrt(bx[q] * bx[q] + by[q] * by[q] + bz[q] * bz[q]);
    vbx[q] = bx[q] / bb[q]; vby[q] = by[q] / bb[q]; vbz[q] = bz[q] / bb[q];
  }
  for (int i = 2; i <= nthe - 1; ++i)
    for (int j = 1; j <= npsi; ++j)
      for (int k = 1; k <= nzeta; ++k) {
        const double dx = X3(x, i + 1, j, k) - X3(x, i, j, k), dy = X3(y, i + 1, j, k) - X3(y, i, j, k), dz = X3(z, i + 1, j, k) - X3(z, i, j, k);
        const double dvx = W3(vbx, i + 1, j, k) - W3(vbx, i, j, k), dvy = W3(vby, i + 1, j, k) - W3(vby, i, j, k),
                     dvz = W3(vbz, i + 1, j, k) - W3(vbz, i, j, k);
        const double bxv = W3(vbx, i, j, k), byv = W3(vby, i, j, k), bzv = W3(vbz, i, j, k);
        if (dx == 0.0) {
          W3(ax, i, j, k) = byv * (dvx / dy) + bzv * (dvx / dz);
          W3(ay, i, j, k) = byv * (dvy / dy) + bzv * (dvy / dz);
          W3(az, i, j, k) = byv * (dvz / dy) + bzv * (dvz / dz);
        } else if (dy == 0.0) {
          W3(ax, i, j, k) = bxv * (dvx / dx) + bzv * (dvx / dz);
          W3(ay, i, j, k) = bxv * (dvy / dx) + bzv * (dvy / dz);
          W3(az, i, j, k) = bxv * (dvz / dx) + bzv * (dvz / dz);
        } else {
          W3(ax, i, j, k) = bxv * dvx / dx + byv * dvx / dy + bzv * dvx / dz;
          W3(ay, i, j, k) = bxv * dvy / dx + byv * dvy / dy + bzv * dvy / dz;
          W3(az, i, j, k) = bxv * dvz / dx + byv * dvz / dy + bzv * dvz / dz;
        }
        W3(ds, i, j, k) = std::sqrt(dx * dx + dy * dy + dz * dz) * REarth;
      }
  for (size_t q = 0; q < n3; ++q) rc[q] = 1. / std::sqrt(ax[q] * ax[q] + ay[q] * ay[q] + az[q] * az[q]) * REarth;
  for (int j = 1; j <= npsi; ++j)
    for (int k = 1; k <= nzeta; ++k) { W3(rc, 1, j, k) = W3(rc, 2, j, k); W3(rc, nthe, j, k) = W3(rc, nthe - 1, j, k); }
  for (int i = 2; i <= nthe - 1; ++i)
    for (int j = 1; j <= npsi; ++j)
      for (int k = 1; k <= nzeta; ++k) {
        W3(dBdS, i, j, k) = bnormal * (W3(bb, i + 1, j, k) - W3(bb, i, j, k)) * 1.0e-9 / W3(ds, i, j, k);
        W3(dRcdS, i, j, k) = (W3(rc, i + 1, j, k) - W3(rc, i, j, k)) / W3(ds, i, j, k);
      }
  for (int j = 1; j <= npsi; ++j)
    for (int k = 1; k <= nzeta; ++k) { W3(dBdS, 1, j, k) = W3(dBdS, 2, j, k); W3(dRcdS, 1, j, k) = W3(dRcdS, 2, j, k); }
  for (int i = 2; i <= nthe - 2; ++i)
    for (int j = 1; j <= npsi; ++j)
      for (int k = 1; k <= nzeta; ++k) {
        W3(d2B, i, j, k) = (W3(dBdS, i + 1, j, k) - W3(dBdS, i, j, k)) / W3(ds, i, j, k);
        W3(d2R, i, j, k) = (W3(dRcdS, i + 1, j, k) - W3(dRcdS, i, j, k)) / W3(ds, i, j, k);
      }
  // equatorial slices (1:npsi, 2:nzeta) as contiguous (npsi, nzeta-1) arrays
  const int ie = nThetaEquator, m1 = nzeta - 1;
  std::vector<double> xe((size_t)npsi * m1), ye(xe.size()), f_rc(xe.size()), f_z1(xe.size()), f_z2(xe.size()), distance(xe.size());
  for (int j = 1; j <= npsi; ++j)
    for (int k = 2; k <= nzeta; ++k) {
      const size_t o = (size_t)(j - 1) + (size_t)npsi * (k - 2);
      xe[o] = X3(x, ie, j, k); ye[o] = X3(y, ie, j, k);
      const double r = W3(rc, ie, j, k);
      f_rc[o] = r;
      f_z1[o] = r * W3(d2R, ie, j, k);
      f_z2[o] = r * r / (bnormal * 1.0e-9 * W3(bb, ie, j, k)) * W3(d2B, ie, j, k);
    }
  const double* fp[3] = {f_rc.data(), f_z1.data(), f_z2.data()};
  for (int i = 2; i <= nR; ++i)
    for (int j = 1; j <= nT - 1; ++j) {
      const double c1 = radRaw[i - 1] * std::cos(azimRaw[j - 1] * 2 * PI / 24. - PI);
      const double c2 = radRaw[i - 1] * std::sin(azimRaw[j - 1] * 2 * PI / 24. - PI);
      double r[3];
      nn9(npsi, m1, xe.data(), ye.data(), 1, (size_t)npsi, fp, 3, c1, c2, distance, r);
      const size_t o = (size_t)(i - 1) + (size_t)nR * (j - 1);
      r_curvEq[o] = r[0]; zeta1Eq[o] = r[1]; zeta2Eq[o] = r[2];
    }
  for (int i = 1; i <= nR; ++i) {                    // MLT = 24 is MLT = 0
    const size_t a = (size_t)(i - 1) + (size_t)nR * (nT - 1), b = (size_t)(i - 1);
    r_curvEq[a] = r_curvEq[b]; zeta1Eq[a] = zeta1Eq[b]; zeta2Eq[a] = zeta2Eq[b];
  }
  for (int j = 1; j <= nT; ++j) {                    // the innermost circle
    const size_t a = (size_t)nR * (j - 1);
    r_curvEq[a] = r_curvEq[a + 1]; zeta1Eq[a] = zeta1Eq[a + 1]; zeta2Eq[a] = zeta2Eq[a + 1];
  }
}
