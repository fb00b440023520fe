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
// cquad's tolerance.  Layouts are the reference's (Fortran order, theta fastest).
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
