// =============================================================================
// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU oracle for the SCB hot path: a restatement in plain C++ (FP64, no
// fast-math, no FMA contraction, reference loop order and operation order,
// Fortran column-major (theta,psi,zeta) layout) of
//
//   computeBandJacob / Compute_convergence   src/ModScbCompute.f90:412-754
//   metrica / metric / newk / newj           src/ModScbEquation.f90:18-665
//   iterateAlpha / iteratePsi (+extap)       src/ModScbEuler.f90:160-299,469-612
//                                            src/ModScbFunctions.f90:57-76
//   GSL_Derivs (3-D driver)                  src/ModRamGSL.f90:794-869,
//                                            src/RamGSL.c:228-291
//   mapAlpha / mapPsi / mapTheta             src/ModScbEuler.f90:15-147,403-457
//   pressure, anisotropic mapping tail       src/ModScbRun.f90:1087-1175
//   GSL_Interpolation_1D (Steffen)           src/ModRamGSL.f90:240-311, src/RamGSL.c:111-174
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may load this.
//
// PARITY PINNING: "parity unpinned".  The reference cannot be built here (no
// Fortran compiler), and the Steffen spline itself lives in GNU GSL (system
// package, libgsl 2.5/2.6 per the reference's Dockerfile / schemeSetup.sh; not
// vendored): steffen_derivs() below restates GSL's interpolation/steffen.c
// (Steffen 1990, A&A 239, 443) from its published algorithm.  The reference's
// only tests touching this path are end-to-end (test3/test4 pressure.ref,
// hI.ref) and need missing input blobs.  Checks that do exist: manufactured
// solutions for the SOR, exactness of the derivative on linear data and
// monotonicity preservation, dipole force balance (J x B ~ 0 for p = 0), and a
// second, independent numpy restatement of computeBandJacob, metrica, metric,
// newk and the lexicographic SOR sweep (tests/independent_scb.py, on the numpy
// Steffen of ramscb_b200/scb_synthetic.py) that agrees with this file bit for bit.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Scb {
  int nthe, npsi, nzeta;
  std::map<std::string, double*> d;
  std::map<std::string, double> s;
  std::map<std::string, int> iv;
  double* D(const char* n) {
    auto it = d.find(n);
    if (it == d.end() || !it->second) { std::fprintf(stderr, "scb oracle: array %s not set\n", n); std::abort(); }
    return it->second;
  }
  double S(const char* n) {
    auto it = s.find(n);
    if (it == s.end()) { std::fprintf(stderr, "scb oracle: scalar %s not set\n", n); std::abort(); }
    return it->second;
  }
  int I(const char* n) {
    auto it = iv.find(n);
    if (it == iv.end()) { std::fprintf(stderr, "scb oracle: int %s not set\n", n); std::abort(); }
    return it->second;
  }
};

// 1-based column-major (nthe, npsi, *) accessor
#define X3(a, i, j, k) (a)[(size_t)((i)-1) + (size_t)nthe * ((size_t)((j)-1) + (size_t)npsi * (size_t)((k)-1))]
#define DIMS const int nthe = o->nthe, npsi = o->npsi, nzeta = o->nzeta; (void)nthe; (void)npsi; (void)nzeta;

inline double sq(double x) { return x * x; }
const double PI_D = 3.141592653589793238462643383279502884197;

// ---- GSL steffen spline, derivative at the nodes -----------------------------
// gsl interpolation/steffen.c (steffen_init + steffen_eval_deriv) as driven by
// interpolation_derivs_c (src/RamGSL.c:255-277): node i < n-1 lies in interval
// i with delx = 0 (=> y'_i); the last node is evaluated through the cubic of
// interval n-2 with delx = h; exact zeros are nudged to 1e-31 (:276).
inline double steffen_copysign(double x, double y) {
  if ((x < 0 && y > 0) || (x > 0 && y < 0)) return -x;
  return x;
}
void steffen_derivs(int n, const double* xa, const double* ya, double* dx, std::vector<double>& yp) {
  yp.resize(n);
  const double h0 = xa[1] - xa[0];
  const double s0 = (ya[1] - ya[0]) / h0;
  yp[0] = s0;
  for (int i = 1; i < n - 1; ++i) {
    const double hi = xa[i + 1] - xa[i];
    const double him1 = xa[i] - xa[i - 1];
    const double si = (ya[i + 1] - ya[i]) / hi;
    const double sim1 = (ya[i] - ya[i - 1]) / him1;
    const double pi = (sim1 * hi + si * him1) / (him1 + hi);
    const double m1 = std::fabs(si) < 0.5 * std::fabs(pi) ? std::fabs(si) : 0.5 * std::fabs(pi);
    const double m2 = std::fabs(sim1) < m1 ? std::fabs(sim1) : m1;
    yp[i] = (steffen_copysign(1.0, sim1) + steffen_copysign(1.0, si)) * m2;
  }
  yp[n - 1] = (ya[n - 1] - ya[n - 2]) / (xa[n - 1] - xa[n - 2]);
  for (int i = 0; i < n - 1; ++i) dx[i] = yp[i];  // c + 0*(...) with delx = 0
  {
    const int i = n - 2;
    const double hi = xa[i + 1] - xa[i];
    const double si = (ya[i + 1] - ya[i]) / hi;
    const double a = (yp[i] + yp[i + 1] - 2 * si) / hi / hi;
    const double b = (3 * si - 2 * yp[i] - yp[i + 1]) / hi;
    const double c = yp[i];
    const double delx = xa[n - 1] - xa[i];
    dx[n - 1] = c + delx * (2.0 * b + delx * 3.0 * a);
  }
  for (int i = 0; i < n; ++i)
    if (dx[i] == 0.0) dx[i] = dx[i] + 1e-31;
}

// Interpolation_3D_Derivs, src/ModRamGSL.f90:794-869; f, dT, dR, dZ are
// (nthe,npsi,nzeta) views whose k-stride is that of the (possibly larger) parent
void derivs3d(Scb* o, const double* f, double* dT, double* dR, double* dZ) {
  DIMS
  const double *tv = o->D("thetaVal"), *rv = o->D("rhoVal"), *zv = o->D("zetaVal");
#pragma omp parallel
  {
    std::vector<double> ya(std::max(nthe, std::max(npsi, nzeta))), dd(ya.size()), yp;
#pragma omp for collapse(2)
    for (int i = 1; i <= nthe; ++i)
      for (int j = 1; j <= npsi; ++j) {
        for (int k = 1; k <= nzeta; ++k) ya[k - 1] = X3(f, i, j, k);
        steffen_derivs(nzeta, zv, ya.data(), dd.data(), yp);
        for (int k = 1; k <= nzeta; ++k) X3(dZ, i, j, k) = dd[k - 1];
      }
#pragma omp for collapse(2)
    for (int i = 1; i <= nthe; ++i)
      for (int k = 1; k <= nzeta; ++k) {
        for (int j = 1; j <= npsi; ++j) ya[j - 1] = X3(f, i, j, k);
        steffen_derivs(npsi, rv, ya.data(), dd.data(), yp);
        for (int j = 1; j <= npsi; ++j) X3(dR, i, j, k) = dd[j - 1];
      }
#pragma omp for collapse(2)
    for (int j = 1; j <= npsi; ++j)
      for (int k = 1; k <= nzeta; ++k) {
        for (int i = 1; i <= nthe; ++i) ya[i - 1] = X3(f, i, j, k);
        steffen_derivs(nthe, tv, ya.data(), dd.data(), yp);
        for (int i = 1; i <= nthe; ++i) X3(dT, i, j, k) = dd[i - 1];
      }
  }
}

// ---- computeBandJacob, src/ModScbCompute.f90:412-496 --------------------------------
int computeBandJacob(Scb* o) {
  DIMS
  const double *x = o->D("x"), *y = o->D("y"), *z = o->D("z"), *f = o->D("f"), *fzet = o->D("fzet");
  double *dXT = o->D("derivXTheta"), *dXR = o->D("derivXRho"), *dXZ = o->D("derivXZeta");
  double *dYT = o->D("derivYTheta"), *dYR = o->D("derivYRho"), *dYZ = o->D("derivYZeta");
  double *dZT = o->D("derivZTheta"), *dZR = o->D("derivZRho"), *dZZ = o->D("derivZZeta");
  double *jac = o->D("jacobian");
  double *gRX = o->D("gradRhoX"), *gRY = o->D("gradRhoY"), *gRZ = o->D("gradRhoZ");
  double *gZX = o->D("gradZetaX"), *gZY = o->D("gradZetaY"), *gZZ = o->D("gradZetaZ");
  double *gTX = o->D("gradThetaX"), *gTY = o->D("gradThetaY"), *gTZ = o->D("gradThetaZ");
  double *GRS = o->D("GradRhoSq"), *GTS = o->D("GradThetaSq"), *GZS = o->D("GradZetaSq");
  double *GRGT = o->D("GradRhoGradTheta"), *GRGZ = o->D("GradRhoGradZeta"), *GTGZ = o->D("GradThetaGradZeta");
  double *Bx = o->D("Bx"), *By = o->D("By"), *Bz = o->D("Bz"), *bsq = o->D("bsq"), *bf = o->D("bf");
  derivs3d(o, x, dXT, dXR, dXZ);
  derivs3d(o, y, dYT, dYR, dYZ);
  derivs3d(o, z, dZT, dZR, dZZ);
  const size_t n = (size_t)nthe * npsi * nzeta;
  for (size_t q = 0; q < n; ++q) {
    jac[q] = dXR[q] * (dYZ[q] * dZT[q] - dYT[q] * dZZ[q]) + dXZ[q] * (dYT[q] * dZR[q] - dYR[q] * dZT[q]) +
             dXT[q] * (dYR[q] * dZZ[q] - dYZ[q] * dZR[q]);
    gRX[q] = (dYZ[q] * dZT[q] - dYT[q] * dZZ[q]) / jac[q];
    gRY[q] = (dZZ[q] * dXT[q] - dZT[q] * dXZ[q]) / jac[q];
    gRZ[q] = (dXZ[q] * dYT[q] - dXT[q] * dYZ[q]) / jac[q];
    gZX[q] = (dYT[q] * dZR[q] - dYR[q] * dZT[q]) / jac[q];
    gZY[q] = (dZT[q] * dXR[q] - dZR[q] * dXT[q]) / jac[q];
    gZZ[q] = (dXT[q] * dYR[q] - dXR[q] * dYT[q]) / jac[q];
    gTX[q] = (dYR[q] * dZZ[q] - dYZ[q] * dZR[q]) / jac[q];
    gTY[q] = (dZR[q] * dXZ[q] - dZZ[q] * dXR[q]) / jac[q];
    gTZ[q] = (dXR[q] * dYZ[q] - dXZ[q] * dYR[q]) / jac[q];
    GRS[q] = gRX[q] * gRX[q] + gRY[q] * gRY[q] + gRZ[q] * gRZ[q];
    GRGZ[q] = gRX[q] * gZX[q] + gRY[q] * gZY[q] + gRZ[q] * gZZ[q];
    GRGT[q] = gRX[q] * gTX[q] + gRY[q] * gTY[q] + gRZ[q] * gTZ[q];
    GTS[q] = gTX[q] * gTX[q] + gTY[q] * gTY[q] + gTZ[q] * gTZ[q];
    GTGZ[q] = gTX[q] * gZX[q] + gTY[q] * gZY[q] + gTZ[q] * gZZ[q];
    GZS[q] = gZX[q] * gZX[q] + gZY[q] * gZY[q] + gZZ[q] * gZZ[q];
  }
  int fail = 0;
  for (int k = 2; k <= nzeta && !fail; ++k)
    for (int j = 1; j <= npsi && !fail; ++j)
      for (int i = 1; i <= nthe; ++i) {
        X3(Bx, i, j, k) = (f[j - 1] * fzet[k - 1] * X3(dXT, i, j, k) / X3(jac, i, j, k));
        X3(By, i, j, k) = (f[j - 1] * fzet[k - 1] * X3(dYT, i, j, k) / X3(jac, i, j, k));
        X3(Bz, i, j, k) = (f[j - 1] * fzet[k - 1] * X3(dZT, i, j, k) / X3(jac, i, j, k));
        X3(bsq, i, j, k) = (X3(GRS, i, j, k) * X3(GZS, i, j, k) - sq(X3(GRGZ, i, j, k))) * sq(f[j - 1] * fzet[k - 1]);
        X3(bf, i, j, k) = std::sqrt(X3(bsq, i, j, k));
        if (std::isnan(X3(bf, i, j, k))) { fail = 1; break; }
      }
  if (fail) return 1;
  for (int j = 1; j <= npsi; ++j)
    for (int i = 1; i <= nthe; ++i) {
      X3(Bx, i, j, 1) = X3(Bx, i, j, nzeta);
      X3(By, i, j, 1) = X3(By, i, j, nzeta);
      X3(Bz, i, j, 1) = X3(Bz, i, j, nzeta);
      X3(bf, i, j, 1) = X3(bf, i, j, nzeta);
      X3(bsq, i, j, 1) = X3(bsq, i, j, nzeta);
    }
  return 0;
}

// ---- one stencil position of metrica / metric: Jacobian, contravariant gradients,
// their squares and dot products (src/ModScbEquation.f90:154-250 / :404-510)
struct Pos {
  double aj, grs, gps, gts, grgp, gpgt, gtgr;
};
inline Pos geom(double xt, double yt, double zt, double xp, double yp, double zp, double xr, double yr, double zr) {
  Pos g;
  g.aj = xr * (yp * zt - yt * zp) + xp * (yt * zr - yr * zt) + xt * (yr * zp - yp * zr);
  const double grx = (yp * zt - yt * zp) / g.aj, gry = (zp * xt - zt * xp) / g.aj, grz = (xp * yt - xt * yp) / g.aj;
  const double gpx = (yt * zr - yr * zt) / g.aj, gpy = (zt * xr - zr * xt) / g.aj, gpz = (xt * yr - xr * yt) / g.aj;
  const double gtx = (yr * zp - yp * zr) / g.aj, gty = (zr * xp - zp * xr) / g.aj, gtz = (xr * yp - xp * yr) / g.aj;
  g.grs = (grx * grx + gry * gry + grz * grz);
  g.gps = (gpx * gpx + gpy * gpy + gpz * gpz);
  g.gts = (gtx * gtx + gty * gty + gtz * gtz);
  g.grgp = (gpx * grx + gpy * gry + gpz * grz);
  g.gpgt = (gpx * gtx + gpy * gty + gpz * gtz);
  g.gtgr = (gtx * grx + gty * gry + gtz * grz);
  return g;
}

struct Spacing {
  double rdr, rdt, rdp, rdrsq, rdtsq, rdpsq, rdr2, rdt2, rdp2, rdr4, rdt4, rdp4, rdpdt4, rdtdr4, dr, dt, dpPrime;
};
Spacing spacing(Scb* o) {  // src/ModScbInit.f90:131-148
  DIMS
  Spacing s;
  s.dr = 1.0 / (double)(npsi - 1);
  s.dt = PI_D / (double)(nthe - 1);
  s.dpPrime = 2 * PI_D / (double)(nzeta - 1);
  s.rdr = 1.0 / s.dr; s.rdt = 1.0 / s.dt; s.rdp = 1.0 / s.dpPrime;
  s.rdrsq = s.rdr * s.rdr; s.rdtsq = s.rdt * s.rdt; s.rdpsq = s.rdp * s.rdp;
  s.rdr2 = 0.5 * s.rdr; s.rdt2 = 0.5 * s.rdt; s.rdp2 = 0.5 * s.rdp;
  s.rdr4 = 0.25 * s.rdr; s.rdt4 = 0.25 * s.rdt; s.rdp4 = 0.25 * s.rdp;
  s.rdpdt4 = 0.25 * s.rdp * s.rdt;
  s.rdtdr4 = 0.25 * s.rdt * s.rdr;
  return s;
}

void zero_vecs(Scb* o) {
  DIMS
  const size_t n = (size_t)nthe * npsi * nzeta;
  for (const char* nm : {"vecd", "vec1", "vec2", "vec3", "vec4", "vec6", "vec7", "vec8", "vec9"}) std::memset(o->D(nm), 0, n * sizeof(double));
}

// ---- metrica, src/ModScbEquation.f90:18-280 (alpha equation, (theta,zeta) stencil) ----
void metrica(Scb* o) {
  DIMS
  const Spacing s = spacing(o);
  const double *x = o->D("x"), *y = o->D("y"), *z = o->D("z");
  double *vecd = o->D("vecd"), *vec1 = o->D("vec1"), *vec2 = o->D("vec2"), *vec3 = o->D("vec3"), *vec4 = o->D("vec4"),
         *vec6 = o->D("vec6"), *vec7 = o->D("vec7"), *vec8 = o->D("vec8"), *vec9 = o->D("vec9");
  zero_vecs(o);
#define DT_A(a) ((X3(a, i + 1, j, k) - X3(a, i, j, k)) * s.rdt)
#define DT_B(a) ((X3(a, i + 1, j, k + 1) + X3(a, i + 1, j, k) - X3(a, i - 1, j, k + 1) - X3(a, i - 1, j, k)) * s.rdt4)
#define DT_C(a) ((X3(a, i, j, k) - X3(a, i - 1, j, k)) * s.rdt)
#define DT_D(a) ((X3(a, i + 1, j, k) + X3(a, i + 1, j, k - 1) - X3(a, i - 1, j, k) - X3(a, i - 1, j, k - 1)) * s.rdt4)
#define DT_E(a) ((X3(a, i + 1, j, k) - X3(a, i - 1, j, k)) * s.rdt2)
#define DP_A(a) ((X3(a, i + 1, j, k + 1) + X3(a, i, j, k + 1) - X3(a, i + 1, j, k - 1) - X3(a, i, j, k - 1)) * s.rdp4)
#define DP_B(a) ((X3(a, i, j, k + 1) - X3(a, i, j, k)) * s.rdp)
#define DP_C(a) ((X3(a, i, j, k + 1) + X3(a, i - 1, j, k + 1) - X3(a, i, j, k - 1) - X3(a, i - 1, j, k - 1)) * s.rdp4)
#define DP_D(a) ((X3(a, i, j, k) - X3(a, i, j, k - 1)) * s.rdp)
#define DP_E(a) ((X3(a, i, j, k + 1) - X3(a, i, j, k - 1)) * s.rdp2)
#define DR_A(a) ((X3(a, i + 1, j + 1, k) + X3(a, i, j + 1, k) - X3(a, i + 1, j - 1, k) - X3(a, i, j - 1, k)) * s.rdr4)
#define DR_B(a) ((X3(a, i, j + 1, k + 1) + X3(a, i, j + 1, k) - X3(a, i, j - 1, k + 1) - X3(a, i, j - 1, k)) * s.rdr4)
#define DR_C(a) ((X3(a, i, j + 1, k) + X3(a, i - 1, j + 1, k) - X3(a, i, j - 1, k) - X3(a, i - 1, j - 1, k)) * s.rdr4)
#define DR_D(a) ((X3(a, i, j + 1, k - 1) + X3(a, i, j + 1, k) - X3(a, i, j - 1, k - 1) - X3(a, i, j - 1, k)) * s.rdr4)
#define DR_E(a) ((X3(a, i, j + 1, k) - X3(a, i, j - 1, k)) * s.rdr2)
#pragma omp parallel for collapse(2)
  for (int j = 2; j <= npsi - 1; ++j)
    for (int k = 2; k <= nzeta; ++k)
      for (int i = 2; i <= nthe - 1; ++i) {
        const Pos a = geom(DT_A(x), DT_A(y), DT_A(z), DP_A(x), DP_A(y), DP_A(z), DR_A(x), DR_A(y), DR_A(z));
        const Pos b = geom(DT_B(x), DT_B(y), DT_B(z), DP_B(x), DP_B(y), DP_B(z), DR_B(x), DR_B(y), DR_B(z));
        const Pos c = geom(DT_C(x), DT_C(y), DT_C(z), DP_C(x), DP_C(y), DP_C(z), DR_C(x), DR_C(y), DR_C(z));
        const Pos d = geom(DT_D(x), DT_D(y), DT_D(z), DP_D(x), DP_D(y), DP_D(z), DR_D(x), DR_D(y), DR_D(z));
        const double v1a = (a.grs * a.gts - sq(a.gtgr)) * a.aj * s.rdtsq;
        const double v1c = (c.grs * c.gts - sq(c.gtgr)) * c.aj * s.rdtsq;
        const double v2a = (a.grs * a.gpgt - a.grgp * a.gtgr) * a.aj * s.rdpdt4;
        const double v2b = (b.grs * b.gpgt - b.grgp * b.gtgr) * b.aj * s.rdpdt4;
        const double v2c = (c.grs * c.gpgt - c.grgp * c.gtgr) * c.aj * s.rdpdt4;
        const double v2d = (d.grs * d.gpgt - d.grgp * d.gtgr) * d.aj * s.rdpdt4;
        const double v3b = (b.grs * b.gps - sq(b.grgp)) * b.aj * s.rdpsq;
        const double v3d = (d.grs * d.gps - sq(d.grgp)) * d.aj * s.rdpsq;
        X3(vecd, i, j, k) = (v1a + v1c) + (v3b + v3d);
        X3(vec1, i, j, k) = (v2c + v2d);
        X3(vec2, i, j, k) = (v2c - v2a) + v3d;
        X3(vec3, i, j, k) = -(v2a + v2d);
        X3(vec4, i, j, k) = v1c + (v2d - v2b);
        X3(vec6, i, j, k) = v1a + (v2b - v2d);
        X3(vec7, i, j, k) = -(v2c + v2b);
        X3(vec8, i, j, k) = v3b + (v2a - v2c);
        X3(vec9, i, j, k) = (v2a + v2b);
      }
}

// ---- metric, src/ModScbEquation.f90:283-540 (psi equation, (theta,rho) stencil) -------
void metric(Scb* o) {
  DIMS
  const Spacing s = spacing(o);
  const double *x = o->D("x"), *y = o->D("y"), *z = o->D("z");
  double *vecd = o->D("vecd"), *vec1 = o->D("vec1"), *vec2 = o->D("vec2"), *vec3 = o->D("vec3"), *vec4 = o->D("vec4"),
         *vec6 = o->D("vec6"), *vec7 = o->D("vec7"), *vec8 = o->D("vec8"), *vec9 = o->D("vec9");
  zero_vecs(o);
#define MT_B(a) ((X3(a, i + 1, j + 1, k) + X3(a, i + 1, j, k) - X3(a, i - 1, j + 1, k) - X3(a, i - 1, j, k)) * s.rdt4)
#define MT_D(a) ((X3(a, i + 1, j, k) + X3(a, i + 1, j - 1, k) - X3(a, i - 1, j, k) - X3(a, i - 1, j - 1, k)) * s.rdt4)
#define MP_B(a) ((X3(a, i, j + 1, k + 1) + X3(a, i, j, k + 1) - X3(a, i, j + 1, k - 1) - X3(a, i, j, k - 1)) * s.rdp4)
#define MP_D(a) ((X3(a, i, j, k + 1) + X3(a, i, j - 1, k + 1) - X3(a, i, j, k - 1) - X3(a, i, j - 1, k - 1)) * s.rdp4)
#define MR_B(a) ((X3(a, i, j + 1, k) - X3(a, i, j, k)) * s.rdr)
#define MR_D(a) ((X3(a, i, j, k) - X3(a, i, j - 1, k)) * s.rdr)
#pragma omp parallel for collapse(2)
  for (int j = 2; j <= npsi - 1; ++j)
    for (int k = 2; k <= nzeta; ++k)
      for (int i = 2; i <= nthe - 1; ++i) {
        const Pos a = geom(DT_A(x), DT_A(y), DT_A(z), DP_A(x), DP_A(y), DP_A(z), DR_A(x), DR_A(y), DR_A(z));
        const Pos b = geom(MT_B(x), MT_B(y), MT_B(z), MP_B(x), MP_B(y), MP_B(z), MR_B(x), MR_B(y), MR_B(z));
        const Pos c = geom(DT_C(x), DT_C(y), DT_C(z), DP_C(x), DP_C(y), DP_C(z), DR_C(x), DR_C(y), DR_C(z));
        const Pos d = geom(MT_D(x), MT_D(y), MT_D(z), MP_D(x), MP_D(y), MP_D(z), MR_D(x), MR_D(y), MR_D(z));
        const double v1a = (sq(a.gpgt) - a.gps * a.gts) * a.aj * s.rdtsq;
        const double v1c = (sq(c.gpgt) - c.gps * c.gts) * c.aj * s.rdtsq;
        const double v2a = (a.grgp * a.gpgt - a.gps * a.gtgr) * a.aj * s.rdtdr4;
        const double v2b = (b.grgp * b.gpgt - b.gps * b.gtgr) * b.aj * s.rdtdr4;
        const double v2c = (c.grgp * c.gpgt - c.gps * c.gtgr) * c.aj * s.rdtdr4;
        const double v2d = (d.grgp * d.gpgt - d.gps * d.gtgr) * d.aj * s.rdtdr4;
        const double v3b = (sq(b.grgp) - b.grs * b.gps) * b.aj * s.rdrsq;
        const double v3d = (sq(d.grgp) - d.grs * d.gps) * d.aj * s.rdrsq;
        X3(vecd, i, j, k) = (v1a + v1c + v3b + v3d);
        X3(vec1, i, j, k) = (v2c + v2d);
        X3(vec2, i, j, k) = ((v2c - v2a) + v3d);
        X3(vec3, i, j, k) = -(v2a + v2d);
        X3(vec4, i, j, k) = (v1c + (v2d - v2b));
        X3(vec6, i, j, k) = (v1a + (v2b - v2d));
        X3(vec7, i, j, k) = -(v2c + v2b);
        X3(vec8, i, j, k) = (v3b + (v2a - v2c));
        X3(vec9, i, j, k) = (v2b + v2a);
      }
}

// ---- newk / newj, src/ModScbEquation.f90:546-665 (Picard; isotropy 0 or 1) -----------
void newk(Scb* o) {
  DIMS
  const int isotropy = o->I("isotropy");
  const double *jac = o->D("jacobian"), *f = o->D("f");
  double* vecx = o->D("vecx");
  if (isotropy == 1) {
    const double* dPdAlpha = o->D("dPdAlpha");
    for (int j = 1; j <= npsi; ++j)
      for (int k = 1; k <= nzeta; ++k)
        for (int i = 1; i <= nthe; ++i) X3(vecx, i, j, k) = -X3(dPdAlpha, i, j, k) * X3(jac, i, j, k) / (f[j - 1] * f[j - 1]);
    return;
  }
  const double *fzet = o->D("fzet"), *bsq = o->D("bsq"), *sigma = o->D("sigma");
  const double *GRS = o->D("GradRhoSq"), *GZS = o->D("GradZetaSq"), *GRGT = o->D("GradRhoGradTheta"), *GRGZ = o->D("GradRhoGradZeta"),
               *GTGZ = o->D("GradThetaGradZeta");
  const double *dPZ = o->D("dPPerdZeta"), *dPT = o->D("dPPerdTheta"), *dBZ = o->D("dBsqdZeta"), *dBT = o->D("dBsqdTheta");
  for (int i = 1; i <= nthe; ++i)
    for (int j = 1; j <= npsi; ++j)
      for (int k = 1; k <= nzeta; ++k) {
        const double xpz = (X3(GRS, i, j, k) * X3(GZS, i, j, k) - sq(X3(GRGZ, i, j, k)));
        const double xpt = (X3(GRS, i, j, k) * X3(GTGZ, i, j, k) - X3(GRGZ, i, j, k) * X3(GRGT, i, j, k));
        const double c0 = -(f[j - 1] * f[j - 1] * fzet[k - 1]) / X3(sigma, i, j, k) / X3(bsq, i, j, k);
        const double tz = X3(dPZ, i, j, k) + 0.5 * (1. - X3(sigma, i, j, k)) * X3(dBZ, i, j, k);
        const double tt = X3(dPT, i, j, k) + 0.5 * (1. - X3(sigma, i, j, k)) * X3(dBT, i, j, k);
        X3(vecx, i, j, k) = X3(jac, i, j, k) / (f[j - 1] * f[j - 1]) * c0 * (tz * xpz + tt * xpt);
      }
}
void newj(Scb* o) {
  DIMS
  const int isotropy = o->I("isotropy");
  const double *jac = o->D("jacobian"), *fzet = o->D("fzet");
  double* vecr = o->D("vecr");
  if (isotropy == 1) {
    const double* dPdPsi = o->D("dPdPsi");
    for (int k = 1; k <= nzeta; ++k)
      for (int j = 1; j <= npsi; ++j)
        for (int i = 1; i <= nthe; ++i) X3(vecr, i, j, k) = X3(jac, i, j, k) * X3(dPdPsi, i, j, k) / (fzet[k - 1] * fzet[k - 1]);
    return;
  }
  const double *f = o->D("f"), *bsq = o->D("bsq"), *sigma = o->D("sigma");
  const double *GRS = o->D("GradRhoSq"), *GZS = o->D("GradZetaSq"), *GRGT = o->D("GradRhoGradTheta"), *GRGZ = o->D("GradRhoGradZeta"),
               *GTGZ = o->D("GradThetaGradZeta");
  const double *dPR = o->D("dPPerdRho"), *dPT = o->D("dPPerdTheta"), *dBR = o->D("dBsqdRho"), *dBT = o->D("dBsqdTheta");
  for (int k = 1; k <= nzeta; ++k)
    for (int j = 1; j <= npsi; ++j)
      for (int i = 1; i <= nthe; ++i) {
        const double xpr = (sq(X3(GRGZ, i, j, k)) - X3(GRS, i, j, k) * X3(GZS, i, j, k));
        const double xpt = (X3(GRGZ, i, j, k) * X3(GTGZ, i, j, k) - X3(GZS, i, j, k) * X3(GRGT, i, j, k));
        const double c0 = -(f[j - 1] * (fzet[k - 1] * fzet[k - 1])) / X3(sigma, i, j, k) / X3(bsq, i, j, k);
        const double tr = X3(dPR, i, j, k) + 0.5 * (1. - X3(sigma, i, j, k)) * X3(dBR, i, j, k);
        const double tt = X3(dPT, i, j, k) + 0.5 * (1. - X3(sigma, i, j, k)) * X3(dBT, i, j, k);
        X3(vecr, i, j, k) = X3(jac, i, j, k) / (fzet[k - 1] * fzet[k - 1]) * c0 * (tr * xpr + tt * xpt);
      }
}

// extap, src/ModScbFunctions.f90:57-76
inline void extap(double x1, double x2, double x3, double& x4) {
  x4 = 3. * x3 - 3. * x2 + x1;
  const double ddx1 = x3 - x2, ddx2 = x2 - x1;
  double ddx = x4 - x3;
  const double pm = ddx * ddx1;
  if (pm > 0.) return;
  if (std::fabs(ddx2) <= 1e-9) { x4 = 2. * x3 - x2; return; }
  ddx = (ddx1 * ddx1) / ddx2;
  x4 = x3 + ddx;
}

// post-processing shared by iterateAlpha / iteratePsi (:262-292 / :575-605)
void sor_post(Scb* o, double* u, int nT, int nP, double wrap, bool alpha_variant) {
  DIMS
  for (int k = 2; k <= nzeta; ++k)
    for (int i = 1 + nT; i <= nthe - nT; ++i)
      for (int j = nP; j >= 1; --j) extap(X3(u, i, npsi - j - 2, k), X3(u, i, npsi - j - 1, k), X3(u, i, npsi - j + 0, k), X3(u, i, npsi - j + 1, k));
  if (nT == 1) {
    // theChange <= 1 (:272-279 / :585-591): the theta end points are extrapolated with extap instead of the linear fill.
    // The psi variant loops k = 2..nzeta.  The alpha variant loops k = 2..nthe-1 over the ZETA index of an array with
    // nzeta+1 planes (:274, a reference quirk): out of bounds whenever nthe-1 > nzeta+1, as on the default grid (101 vs
    // 98) -- restated with the loop clipped to the planes that exist.  No shipped PARAM uses theChange <= 1 (default 4);
    // the device library answers RSG_ERR_UNSUPPORTED for it.
    const int kend = alpha_variant ? std::min(nthe - 1, nzeta + 1) : nzeta;
    for (int k = 2; k <= kend; ++k)
      for (int j = 1; j <= npsi; ++j) {
        extap(X3(u, nthe - 3, j, k), X3(u, nthe - 2, j, k), X3(u, nthe - 1, j, k), X3(u, nthe, j, k));
        extap(X3(u, 4, j, k), X3(u, 3, j, k), X3(u, 2, j, k), X3(u, 1, j, k));
      }
  } else {
    for (int j = 1; j <= npsi; ++j)
      for (int k = 1; k <= nzeta; ++k)
        for (int i = 1; i <= nT; ++i) {
          X3(u, i, j, k) = X3(u, nT + 1, j, k) + (nT + 1 - i) * (X3(u, 1, j, k) - X3(u, nT + 1, j, k)) / nT;
          X3(u, nthe - i + 1, j, k) = X3(u, nthe - nT - 1, j, k) + (nT + 1 - i) * (X3(u, nthe, j, k) - X3(u, nthe - nT - 1, j, k)) / (nT);
        }
  }
  for (int j = 1; j <= npsi; ++j)
    for (int i = 1; i <= nthe; ++i) {
      X3(u, i, j, 1) = X3(u, i, j, nzeta) - wrap;
      X3(u, i, j, nzeta + 1) = X3(u, i, j, 2) + wrap;
    }
}

// ---- iterateAlpha, src/ModScbEuler.f90:160-299 ----------------------------------------
// outputs: nisave, sumb, sumdb, diffmx (scalars), ni per surface; returns SORFail
int iterateAlpha(Scb* o, int* ni_out) {
  DIMS
  const int nimax = o->I("nimax"), nT = std::max(o->I("theChange"), 1), nP = std::max(o->I("psiChange"), 1);
  const double InCon = o->S("InConAlpha");
  const double *vecd = o->D("vecd"), *vec1 = o->D("vec1"), *vec2 = o->D("vec2"), *vec3 = o->D("vec3"), *vec4 = o->D("vec4"),
               *vec6 = o->D("vec6"), *vec7 = o->D("vec7"), *vec8 = o->D("vec8"), *vec9 = o->D("vec9"), *vecx = o->D("vecx");
  double* alfa = o->D("alfa");
  const size_t n1 = (size_t)nthe * npsi * (nzeta + 1);
  std::vector<double> prev(alfa, alfa + n1), resid(n1, 0.0);
  std::vector<int> ni(npsi + 1, 0);
  const double rjac = 1.0 - 2.0 * PI_D * PI_D / ((double)nzeta * (double)nzeta + (double)nthe * (double)nthe);
  const double omegaOpt = 2.0 / (1.0 + std::sqrt(1.0 - rjac * rjac));
  int fail = 0;
  double* res = resid.data();
#pragma omp parallel for schedule(dynamic, 1)
  for (int jz = 2; jz <= npsi - nP; ++jz) {
    double om = 1.0;
    ni[jz] = 1;
    bool stop = false;
    while (ni[jz] <= nimax && !stop) {
      for (int k = 2; k <= nzeta && !stop; ++k) {
        const int kp = k + 1, km = k - 1;
        for (int iz = 1 + nT; iz <= nthe - nT; ++iz) {
          const int im = iz - 1, ip = iz + 1;
          X3(res, iz, jz, k) = -X3(vecd, iz, jz, k) * X3(alfa, iz, jz, k) + X3(vec1, iz, jz, k) * X3(alfa, im, jz, km) +
                               X3(vec2, iz, jz, k) * X3(alfa, iz, jz, km) + X3(vec3, iz, jz, k) * X3(alfa, ip, jz, km) +
                               X3(vec4, iz, jz, k) * X3(alfa, im, jz, k) + X3(vec6, iz, jz, k) * X3(alfa, ip, jz, k) +
                               X3(vec7, iz, jz, k) * X3(alfa, im, jz, kp) + X3(vec8, iz, jz, k) * X3(alfa, iz, jz, kp) +
                               X3(vec9, iz, jz, k) * X3(alfa, ip, jz, kp) - X3(vecx, iz, jz, k);
          X3(alfa, iz, jz, k) = X3(alfa, iz, jz, k) + om * (X3(res, iz, jz, k) / X3(vecd, iz, jz, k));
          if (std::isnan(X3(alfa, iz, jz, k)) || X3(alfa, iz, jz, k) >= 1e10) {
            X3(alfa, iz, jz, k) = prev[(&X3(alfa, iz, jz, k)) - alfa];
            X3(res, iz, jz, k) = 0.0;
#pragma omp atomic write
            fail = 1;
            stop = true;
            break;
          }
        }
      }
      if (stop) break;
      om = omegaOpt;
      double mx = 0.0;
      for (int k = 2; k <= nzeta; ++k)
        for (int i = 2; i <= nthe - 1; ++i) mx = std::max(mx, std::fabs(X3(res, i, jz, k)));
      if (mx < InCon) break;
      ni[jz] = ni[jz] + 1;
    }
  }
  int nisave = 0;
  for (int jz = 1; jz <= npsi; ++jz) nisave = std::max(nisave, ni[jz]);
  double sumdb = 0.0, sumb = 0.0, diffmx = 0.0;
  for (int k = 2; k <= nzeta; ++k)
    for (int j = 2; j <= npsi - 1; ++j)
      for (int i = 2; i <= nthe - 1; ++i) {
        sumdb += std::fabs(X3(alfa, i, j, k) - X3(prev.data(), i, j, k));
        sumb += std::fabs(X3(alfa, i, j, k));
        diffmx = std::max(diffmx, std::fabs(X3(res, i, j, k)));
      }
  o->s["nisave"] = nisave; o->s["sumdb"] = sumdb; o->s["sumb"] = sumb; o->s["diffmx"] = diffmx;
  if (ni_out)
    for (int jz = 1; jz <= npsi; ++jz) ni_out[jz - 1] = ni[jz];
  sor_post(o, alfa, nT, nP, 2.0 * PI_D, true);
  return fail;
}

// ---- iteratePsi, src/ModScbEuler.f90:469-612 ------------------------------------------
int iteratePsi(Scb* o, int* ni_out) {
  DIMS
  const int nimax = o->I("nimax"), nT = std::max(o->I("theChange"), 1), nP = std::max(o->I("psiChange"), 1);
  const double InCon = o->S("InConPsi");
  const double *vecd = o->D("vecd"), *vec1 = o->D("vec1"), *vec2 = o->D("vec2"), *vec3 = o->D("vec3"), *vec4 = o->D("vec4"),
               *vec6 = o->D("vec6"), *vec7 = o->D("vec7"), *vec8 = o->D("vec8"), *vec9 = o->D("vec9"), *vecr = o->D("vecr");
  double* psi = o->D("psi");
  const size_t n1 = (size_t)nthe * npsi * (nzeta + 1);
  std::vector<double> prev(psi, psi + n1), resid(n1, 0.0);
  std::vector<int> ni(nzeta + 1, 0);
  const double rjac = 1.0 - 2.0 * PI_D * PI_D / ((double)nthe * (double)nthe + (double)npsi * (double)npsi);
  const double omegaOpt = 2.0 / (1.0 + std::sqrt(1.0 - rjac * rjac));
  int fail = 0;
  double* res = resid.data();
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 2; k <= nzeta; ++k) {
    double om = 1.0;
    ni[k] = 1;
    bool stop = false;
    while (ni[k] <= nimax && !stop) {
      for (int jz = 2; jz <= npsi - nP && !stop; ++jz) {
        const int jp = jz + 1, jm = jz - 1;
        for (int iz = 1 + nT; iz <= nthe - nT; ++iz) {
          const int im = iz - 1, ip = iz + 1;
          X3(res, iz, jz, k) = -X3(vecd, iz, jz, k) * X3(psi, iz, jz, k) + X3(vec1, iz, jz, k) * X3(psi, im, jm, k) +
                               X3(vec2, iz, jz, k) * X3(psi, iz, jm, k) + X3(vec3, iz, jz, k) * X3(psi, ip, jm, k) +
                               X3(vec4, iz, jz, k) * X3(psi, im, jz, k) + X3(vec6, iz, jz, k) * X3(psi, ip, jz, k) +
                               X3(vec7, iz, jz, k) * X3(psi, im, jp, k) + X3(vec8, iz, jz, k) * X3(psi, iz, jp, k) +
                               X3(vec9, iz, jz, k) * X3(psi, ip, jp, k) - X3(vecr, iz, jz, k);
          X3(psi, iz, jz, k) = X3(psi, iz, jz, k) + om * X3(res, iz, jz, k) / X3(vecd, iz, jz, k);
          if (std::isnan(X3(psi, iz, jz, k)) || X3(psi, iz, jz, k) >= 1e10) {
            X3(psi, iz, jz, k) = prev[(&X3(psi, iz, jz, k)) - psi];
            X3(res, iz, jz, k) = 0.0;
#pragma omp atomic write
            fail = 1;
            stop = true;
            break;
          }
        }
      }
      if (stop) break;
      om = omegaOpt;
      double mx = 0.0;
      for (int j = 2; j <= npsi - 1; ++j)
        for (int i = 2; i <= nthe - 1; ++i) mx = std::max(mx, std::fabs(X3(res, i, j, k)));
      if (mx < InCon) break;
      ni[k] = ni[k] + 1;
    }
  }
  int nisave = 0;
  for (int k = 1; k <= nzeta; ++k) nisave = std::max(nisave, ni[k]);
  double sumdb = 0.0, sumb = 0.0, diffmx = 0.0;
  for (int k = 2; k <= nzeta; ++k)
    for (int j = 2; j <= npsi - 1; ++j)
      for (int i = 2; i <= nthe - 1; ++i) {
        sumdb += std::fabs(X3(psi, i, j, k) - X3(prev.data(), i, j, k));
        sumb += std::fabs(X3(psi, i, j, k));
        diffmx = std::max(diffmx, std::fabs(X3(res, i, j, k)));
      }
  o->s["nisave"] = nisave; o->s["sumdb"] = sumdb; o->s["sumb"] = sumb; o->s["diffmx"] = diffmx;
  if (ni_out)
    for (int k = 1; k <= nzeta; ++k) ni_out[k - 1] = ni[k];
  sor_post(o, psi, nT, nP, 0.0, false);
  return fail;
}

// ---- GSL_Interpolation_1D, Steffen (src/ModRamGSL.f90:240-311 + interpolation_1d_c,
// src/RamGSL.c:111-174; gsl interpolation/steffen.c steffen_init + steffen_eval) -----------
// The Fortran wrapper first drops abscissae that do not increase (:262-273); the C driver
// extrapolates linearly outside [xa_0, xa_{n-1}] (:159-164, end points included) and evaluates
// the Steffen cubic d + delx*(c + delx*(b + delx*a)) of the interval found by bisection
// (gsl_interp_bsearch: xa[i] <= x < xa[i+1]) inside.  Returns GSLerr (0 = ok).
int interp1d(int n0, const double* x1, const double* f1, int n2, const double* x2, double* f2) {
  std::vector<double> xa(n0), fa(n0);
  int n1 = 1;
  xa[0] = x1[0];
  fa[0] = f1[0];
  for (int i = 1; i < n0; ++i)
    if (x1[i] > xa[n1 - 1]) { xa[n1] = x1[i]; fa[n1] = f1[i]; ++n1; }
  if (n1 < 3) return 1;   // gsl_interp_steffen needs 3 points (gsl_spline_alloc fails)
  std::vector<double> yp(n1), a(n1 - 1), b(n1 - 1);
  {
    const double h0 = xa[1] - xa[0];
    yp[0] = (fa[1] - fa[0]) / h0;
    for (int i = 1; i < n1 - 1; ++i) {
      const double hi = xa[i + 1] - xa[i];
      const double him1 = xa[i] - xa[i - 1];
      const double si = (fa[i + 1] - fa[i]) / hi;
      const double sim1 = (fa[i] - fa[i - 1]) / him1;
      const double pi = (sim1 * hi + si * him1) / (him1 + hi);
      const double m1 = std::fabs(si) < 0.5 * std::fabs(pi) ? std::fabs(si) : 0.5 * std::fabs(pi);
      const double m2 = std::fabs(sim1) < m1 ? std::fabs(sim1) : m1;
      yp[i] = (steffen_copysign(1.0, sim1) + steffen_copysign(1.0, si)) * m2;
    }
    yp[n1 - 1] = (fa[n1 - 1] - fa[n1 - 2]) / (xa[n1 - 1] - xa[n1 - 2]);
    for (int i = 0; i < n1 - 1; ++i) {
      const double hi = xa[i + 1] - xa[i];
      const double si = (fa[i + 1] - fa[i]) / hi;
      a[i] = (yp[i] + yp[i + 1] - 2 * si) / hi / hi;
      b[i] = (3 * si - 2 * yp[i] - yp[i + 1]) / hi;
    }
  }
  for (int q = 0; q < n2; ++q) {
    const double xb = x2[q];
    if (xb <= xa[0]) {
      f2[q] = fa[0] + (xb - xa[0]) / (xa[1] - xa[0]) * (fa[1] - fa[0]);
    } else if (xb >= xa[n1 - 1]) {
      f2[q] = fa[n1 - 1] + (xb - xa[n1 - 1]) / (xa[n1 - 2] - xa[n1 - 1]) * (fa[n1 - 2] - fa[n1 - 1]);
    } else if (xb == xb) {
      int ilo = 0, ihi = n1 - 1;
      while (ihi > ilo + 1) {
        const int i = (ihi + ilo) / 2;
        if (xa[i] > xb) ihi = i; else ilo = i;
      }
      const double delx = xb - xa[ilo];
      f2[q] = fa[ilo] + delx * (yp[ilo] + delx * (b[ilo] + delx * a[ilo]));
    } else {
      return 1;   // NaN abscissa: gsl_spline_eval_e reports GSL_EDOM
    }
  }
  return 0;
}

void wrap_xyz(Scb* o) {   // periodic boundary conditions in zeta (src/ModScbEuler.f90:66-71)
  DIMS
  for (const char* n : {"x", "y", "z"}) {
    double* a = o->D(n);
    for (int j = 1; j <= npsi; ++j)
      for (int i = 1; i <= nthe; ++i) {
        X3(a, i, j, 1) = X3(a, i, j, nzeta);
        X3(a, i, j, nzeta + 1) = X3(a, i, j, 2);
      }
  }
}

// ---- mapAlpha, src/ModScbEuler.f90:97-147: move the grid points along zeta so that the computed
// alfa takes the prescribed values alphaVal(k); then alfges (:82-94) resets alfa --------------
int mapAlpha(Scb* o) {
  DIMS
  double *x = o->D("x"), *y = o->D("y"), *z = o->D("z"), *alfa = o->D("alfa");
  const double* alphaVal = o->D("alphaVal");
  const int n = nzeta + 1;
  std::vector<double> xo(n), yo(n), zo(n), ao(n), out(n);
  for (int j = 1; j <= npsi; ++j)
    for (int i = 1; i <= nthe; ++i) {
      for (int k = 1; k <= n; ++k) { xo[k - 1] = X3(x, i, j, k); yo[k - 1] = X3(y, i, j, k); zo[k - 1] = X3(z, i, j, k); ao[k - 1] = X3(alfa, i, j, k); }
      double* arrs[3] = {x, y, z};
      const double* olds[3] = {xo.data(), yo.data(), zo.data()};
      for (int c = 0; c < 3; ++c) {
        if (interp1d(n, ao.data(), olds[c], nzeta - 1, alphaVal + 1, out.data())) return 1;   // SORFail
        for (int k = 2; k <= nzeta; ++k) X3(arrs[c], i, j, k) = out[k - 2];
      }
    }
  wrap_xyz(o);
  for (int k = 1; k <= n; ++k)
    for (int j = 1; j <= npsi; ++j)
      for (int i = 1; i <= nthe; ++i) X3(alfa, i, j, k) = alphaVal[k - 1];
  return 0;
}

// ---- mapPsi, src/ModScbEuler.f90:403-457 (+ psiges :387-400) ------------------------------------
int mapPsi(Scb* o) {
  DIMS
  double *x = o->D("x"), *y = o->D("y"), *z = o->D("z"), *psi = o->D("psi");
  const double* psiVal = o->D("psiVal");
  std::vector<double> xo(npsi), yo(npsi), zo(npsi), po(npsi), out(npsi);
  for (int k = 2; k <= nzeta; ++k)
    for (int i = 1; i <= nthe; ++i) {
      for (int j = 1; j <= npsi; ++j) { xo[j - 1] = X3(x, i, j, k); yo[j - 1] = X3(y, i, j, k); zo[j - 1] = X3(z, i, j, k); po[j - 1] = X3(psi, i, j, k); }
      double* arrs[3] = {x, y, z};
      const double* olds[3] = {xo.data(), yo.data(), zo.data()};
      for (int c = 0; c < 3; ++c) {
        if (interp1d(npsi, po.data(), olds[c], npsi, psiVal, out.data())) return 1;
        for (int j = 1; j <= npsi; ++j) X3(arrs[c], i, j, k) = out[j - 1];
      }
    }
  wrap_xyz(o);
  for (int k = 1; k <= nzeta + 1; ++k)
    for (int j = 1; j <= npsi; ++j)
      for (int i = 1; i <= nthe; ++i) X3(psi, i, j, k) = psiVal[j - 1];
  return 0;
}

// ---- mapTheta, src/ModScbEuler.f90:15-75: redistribute the points of every field line to the
// prescribed arc-length fractions chiVal -------------------------------------------------------
int mapTheta(Scb* o) {
  DIMS
  double *x = o->D("x"), *y = o->D("y"), *z = o->D("z");
  const double* chiVal = o->D("chiVal");
  std::vector<double> xo(nthe), yo(nthe), zo(nthe), dist(nthe), chiOld(nthe), out(nthe);
  for (int k = 2; k <= nzeta; ++k)
    for (int j = 1; j <= npsi; ++j) {
      dist[0] = 0.0;
      for (int i = 1; i <= nthe; ++i) { xo[i - 1] = X3(x, i, j, k); yo[i - 1] = X3(y, i, j, k); zo[i - 1] = X3(z, i, j, k); }
      for (int i = 2; i <= nthe; ++i)
        dist[i - 1] = dist[i - 2] + std::sqrt(sq(X3(x, i, j, k) - X3(x, i - 1, j, k)) + sq(X3(y, i, j, k) - X3(y, i - 1, j, k)) +
                                               sq(X3(z, i, j, k) - X3(z, i - 1, j, k)));
      for (int i = 0; i < nthe; ++i) chiOld[i] = dist[i] / dist[nthe - 1] * PI_D;
      double* arrs[3] = {x, y, z};
      const double* olds[3] = {xo.data(), yo.data(), zo.data()};
      for (int c = 0; c < 3; ++c) {
        if (interp1d(nthe, chiOld.data(), olds[c], nthe, chiVal, out.data())) return 1;
        for (int i = 1; i <= nthe; ++i) X3(arrs[c], i, j, k) = out[i - 1];
      }
    }
  wrap_xyz(o);
  return 0;
}

// ---- pressure, anisotropic branch, from the normalised equatorial pressures on
// (src/ModScbRun.f90:1087-1175): field-line mapping of pper / ppar with the iLossCone = 1
// (filled loss cone) or 2 (Liemohn 2004) formulas, sigma, tau, the optional reduction of the
// anisotropy to marginal mirror stability (:1127-1160), the Steffen derivatives of pper and bsq
// and their conversion to Euler-potential derivatives.  Inputs: pperEq, pparEq (npsi, nzeta+1),
// bf, bsq of the last computeBandJacob.  `1./6.` at :1133 is a default-real constant expression:
// it is evaluated in single precision and then promoted.
struct AnisoPt { double pper, ppar, sigma, tau; };
inline AnisoPt aniso_point(double pperEq, double pparEq, double bfEq, double bfI, double bf1, double bsqI, int iLossCone,
                           bool clampRatio) {
  const double pEq = (2. * pperEq + pparEq) / 3.;
  const double aratio = pperEq / pparEq - 1.;
  const double aL = -aratio / (aratio + 1);
  double ratioB = bfEq / bfI;
  if (clampRatio) ratioB = ratioB < 1.0 ? ratioB : 1.0;
  AnisoPt r;
  if (iLossCone == 2) {
    const double q = bf1 / bfI;
    const double rBI = q > 1. + 1.E-9 ? q : 1. + 1.E-9;
    const double pparN = pparEq * (1. - (ratioB + aL * ratioB) / (rBI + aL * ratioB));
    const double pperN = pperEq * (1. - (ratioB + aL * ratioB) / (rBI + aL * ratioB));
    const double aN = pparN / pperN - 1.;
    r.ppar = pparN * (aN + 1.) / (1. + aN * ratioB) * std::sqrt((rBI - 1.) / (rBI - ratioB)) * (1. - (1. + aN * ratioB) / (rBI + aN * ratioB));
    r.pper = r.ppar / (1. + aN * ratioB);
  } else {
    const double gParam = 1. / sq(1. + aratio * (1. - ratioB));
    r.ppar = pEq * 1. / (1. + 2. * aratio / 3.) * std::sqrt(gParam);
    r.pper = pEq * (aratio + 1.) / (1. + 2. * aratio / 3.) * gParam;
  }
  r.sigma = 1. + (r.pper - r.ppar) / bsqI;
  r.tau = 1. - 2. * (r.pper - r.ppar) / bsqI * r.pper / r.ppar;
  return r;
}
void pressure_aniso(Scb* o) {
  DIMS
  const int iLossCone = o->I("iLossCone"), iReduce = o->I("iReduceAnisotropy");
  const int eq = (nthe + 1) / 2;                     // nThetaEquator, src/ModScbInit.f90:234
  const double *pperEq = o->D("pperEq"), *pparEq = o->D("pparEq"), *bf = o->D("bf"), *bsq = o->D("bsq");
  const double *f = o->D("f"), *fzet = o->D("fzet");
  double *pper = o->D("pper"), *ppar = o->D("ppar"), *sigma = o->D("sigma"), *tau = o->D("tau");
#define E2(a, j, k) (a)[(size_t)((j)-1) + (size_t)npsi * (size_t)((k)-1)]
  for (int k = 1; k <= nzeta; ++k)
    for (int j = 1; j <= npsi; ++j) {
      for (int i = 1; i <= nthe; ++i) {
        const AnisoPt r = aniso_point(E2(pperEq, j, k), E2(pparEq, j, k), X3(bf, eq, j, k), X3(bf, i, j, k), X3(bf, 1, j, k),
                                      X3(bsq, i, j, k), iLossCone, true);
        X3(pper, i, j, k) = r.pper; X3(ppar, i, j, k) = r.ppar; X3(sigma, i, j, k) = r.sigma; X3(tau, i, j, k) = r.tau;
      }
      if (iReduce == 1 && X3(tau, eq, j, k) < 0.) {   // Mirror_unstable, :1130-1158
        const double pEq = (2. * E2(pperEq, j, k) + E2(pparEq, j, k)) / 3.;
        const double bEqSq = X3(bsq, eq, j, k);
        const double sixth = (double)(1.f / 6.f);
        const double pe = sixth * (3. * pEq - bEqSq + std::sqrt(sq(bEqSq) + 12. * bEqSq * pEq + 9. * sq(pEq)));
        const double pa = 3. * pEq - 2. * pe;
        for (int i = 1; i <= nthe; ++i) {
          AnisoPt r = aniso_point(pe, pa, X3(bf, eq, j, k), X3(bf, i, j, k), X3(bf, 1, j, k), X3(bsq, i, j, k), iLossCone, false);
          if (iLossCone == 1) {                       // :1149-1151 keeps pEq = press(j,k) of the first pass
            const double aratio = pe / pa - 1.;
            const double ratioB = X3(bf, eq, j, k) / X3(bf, i, j, k);
            const double gParam = 1. / sq(1. + aratio * (1. - ratioB));
            r.ppar = pEq * 1. / (1. + 2. * aratio / 3.) * std::sqrt(gParam);
            r.pper = pEq * (aratio + 1.) / (1. + 2. * aratio / 3.) * gParam;
            r.sigma = 1.0 + (r.pper - r.ppar) / X3(bsq, i, j, k);
            r.tau = 1. - 2. * (r.pper - r.ppar) / X3(bsq, i, j, k) * r.pper / r.ppar;
          }
          X3(pper, i, j, k) = r.pper; X3(ppar, i, j, k) = r.ppar; X3(sigma, i, j, k) = r.sigma; X3(tau, i, j, k) = r.tau;
        }
      }
    }
#undef E2
  derivs3d(o, pper, o->D("dPPerdTheta"), o->D("dPPerdRho"), o->D("dPPerdZeta"));
  derivs3d(o, bsq, o->D("dBsqdTheta"), o->D("dBsqdRho"), o->D("dBsqdZeta"));
  double *dPR = o->D("dPPerdRho"), *dBR = o->D("dBsqdRho"), *dPZ = o->D("dPPerdZeta"), *dBZ = o->D("dBsqdZeta");
  double *dPP = o->D("dPPerdPsi"), *dBP = o->D("dBsqdPsi"), *dPA = o->D("dPPerdAlpha"), *dBA = o->D("dBsqdAlpha");
  for (int k = 1; k <= nzeta; ++k)
    for (int j = 1; j <= npsi; ++j)
      for (int i = 1; i <= nthe; ++i) {
        X3(dPP, i, j, k) = 1. / f[j - 1] * X3(dPR, i, j, k);
        X3(dBP, i, j, k) = 1. / f[j - 1] * X3(dBR, i, j, k);
        X3(dPA, i, j, k) = 1. / fzet[k - 1] * X3(dPZ, i, j, k);
        X3(dBA, i, j, k) = 1. / fzet[k - 1] * X3(dBZ, i, j, k);
      }
}

// ---- Compute_convergence, src/ModScbCompute.f90:499-754 (isotropy 0 and 1) ----------------
// produces jGradRho/jGradZeta/jGradTheta, Jx..Jz, GradPx..GradPz, jCrossB, GradP and
// the three norms
int compute_convergence(Scb* o) {
  DIMS
  const int isotropy = o->I("isotropy");
  const Spacing s = spacing(o);
  const double *dPdAlpha = o->D("dPdAlpha"), *dPdPsi = o->D("dPdPsi");
  const double bnormal = o->S("bnormal"), pnormal = o->S("pnormal"), pjconst = o->S("pjconst");
  const double *f = o->D("f"), *fzet = o->D("fzet"), *jac = o->D("jacobian"), *bsq = o->D("bsq"), *sigma = o->D("sigma");
  const double *GRS = o->D("GradRhoSq"), *GTS = o->D("GradThetaSq"), *GZS = o->D("GradZetaSq"), *GRGT = o->D("GradRhoGradTheta"),
               *GRGZ = o->D("GradRhoGradZeta"), *GTGZ = o->D("GradThetaGradZeta");
  const double *dPA = o->D("dPPerdAlpha"), *dPP = o->D("dPPerdPsi"), *dPT = o->D("dPPerdTheta"), *dPR = o->D("dPPerdRho"),
               *dPZ = o->D("dPPerdZeta"), *dBA = o->D("dBsqdAlpha"), *dBP = o->D("dBsqdPsi"), *dBT = o->D("dBsqdTheta");
  const double *pper = o->D("pper"), *ppar = o->D("ppar");
  const double *dXT = o->D("derivXTheta"), *dXR = o->D("derivXRho"), *dXZ = o->D("derivXZeta");
  const double *dYT = o->D("derivYTheta"), *dYR = o->D("derivYRho"), *dYZ = o->D("derivYZeta");
  const double *dZT = o->D("derivZTheta"), *dZR = o->D("derivZRho"), *dZZ = o->D("derivZZeta");
  double *jGR = o->D("jGradRho"), *jGZ = o->D("jGradZeta"), *jGT = o->D("jGradTheta");
  double *Jx = o->D("Jx"), *Jy = o->D("Jy"), *Jz = o->D("Jz"), *GPx = o->D("GradPx"), *GPy = o->D("GradPy"), *GPz = o->D("GradPz");
  double *jCrossB = o->D("jCrossB"), *GradP = o->D("GradP");
  const size_t n = (size_t)nthe * npsi * nzeta;
  std::vector<double> pR(n), pZ(n), dpR(n), dpZ(n), nu1(n), nu2(n), dDiff(n), jdiff(n), jCBsq(n), gPsq(n);
  for (int i = 1; i <= nthe; ++i)
    for (int j = 1; j <= npsi; ++j)
      for (int k = 1; k <= nzeta; ++k) {
        const double sg = X3(sigma, i, j, k), fj = f[j - 1], fk = fzet[k - 1];
        if (isotropy == 1) {                                   // :553-556, :575-578
          X3(jGR, i, j, k) = -1.0 / fj * X3(dPdAlpha, i, j, k);
          X3(jGZ, i, j, k) = 1.0 / fk * X3(dPdPsi, i, j, k);
          continue;
        }
        X3(jGR, i, j, k) = 1.0 / fj * (-1. / sg * X3(dPA, i, j, k) -
                                       1. / (sg * X3(bsq, i, j, k)) * (fj * fj) * fk *
                                           (X3(GRS, i, j, k) * X3(GTGZ, i, j, k) - X3(GRGT, i, j, k) * X3(GRGZ, i, j, k)) *
                                           (X3(dPT, i, j, k) + (1. - sg) * 0.5 * X3(dBT, i, j, k)) -
                                       (1. - sg) / sg * 0.5 * X3(dBA, i, j, k));
        X3(jGZ, i, j, k) = 1.0 / fk * (1. / sg * X3(dPP, i, j, k) -
                                       1. / (sg * X3(bsq, i, j, k)) * fj * (fk * fk) *
                                           (X3(GRGZ, i, j, k) * X3(GTGZ, i, j, k) - X3(GRGT, i, j, k) * X3(GZS, i, j, k)) *
                                           (X3(dPT, i, j, k) + (1. - sg) * 0.5 * X3(dBT, i, j, k)) +
                                       (1. - sg) / sg * 0.5 * X3(dBP, i, j, k));
      }
  for (int j = 1; j <= npsi; ++j)
    for (int k = 1; k <= nzeta; ++k)
      for (int i = 1; i <= nthe; ++i) {
        X3(pR.data(), i, j, k) = X3(jac, i, j, k) * f[j - 1] * fzet[k - 1] * (X3(GRGT, i, j, k) * X3(GRGZ, i, j, k) - X3(GTGZ, i, j, k) * X3(GRS, i, j, k));
        X3(pZ.data(), i, j, k) = X3(jac, i, j, k) * f[j - 1] * fzet[k - 1] * (X3(GRGT, i, j, k) * X3(GZS, i, j, k) - X3(GRGZ, i, j, k) * X3(GTGZ, i, j, k));
      }
  derivs3d(o, pR.data(), nu1.data(), dpR.data(), nu2.data());
  derivs3d(o, pZ.data(), nu1.data(), nu2.data(), dpZ.data());
  for (size_t q = 0; q < n; ++q) jGT[q] = (dpR[q] + dpZ[q]) / jac[q];
  for (size_t q = 0; q < n; ++q) {
    // pper/ppar are (nthe,npsi,nzeta+1): same linear index for k <= nzeta
    jdiff[q] = jac[q] * (pper[q] - ppar[q]);
  }
  derivs3d(o, jdiff.data(), dDiff.data(), nu1.data(), nu2.data());
  for (size_t q = 0; q < n; ++q) {
    Jx[q] = (jGR[q] * dXR[q] + jGZ[q] * dXZ[q] + jGT[q] * dXT[q]);
    Jy[q] = (jGR[q] * dYR[q] + jGZ[q] * dYZ[q] + jGT[q] * dYT[q]);
    Jz[q] = (jGR[q] * dZR[q] + jGZ[q] * dZZ[q] + jGT[q] * dZT[q]);
  }
  for (int k = 1; k <= nzeta; ++k)
    for (int j = 1; j <= npsi; ++j)
      for (int i = 1; i <= nthe; ++i) {
        const size_t q = &X3(jac, i, j, k) - jac;
        const double fj = f[j - 1], fk = fzet[k - 1];
        jCBsq[q] = (fj * fj) * (fk * fk) *
                   (GRS[q] * sq(jGZ[q]) + GZS[q] * sq(jGR[q]) - 2.0 * jGZ[q] * jGR[q] * GRGZ[q]);
        gPsq[q] = GRS[q] * sq(dPR[q]) + GZS[q] * sq(dPZ[q]) + GTS[q] * sq(dPT[q]) + 2. * dPR[q] * dPZ[q] * GRGZ[q] +
                  2. * dPR[q] * dPT[q] * GRGT[q] + 2. * dPZ[q] * dPT[q] * GTGZ[q] + sq(dDiff[q] / jac[q]) -
                  2. * dPT[q] * dDiff[q] / jac[q];
        if (isotropy == 1) {                                   // :671-690
          const double a1 = (fj * dPdPsi[q] * GRS[q] + fk * dPdAlpha[q] * GRGZ[q]);
          const double a2 = (fj * dPdPsi[q] * GRGZ[q] + fk * dPdAlpha[q] * GZS[q]);
          const double a3 = (fj * dPdPsi[q] * GRGT[q] + fk * dPdAlpha[q] * GTGZ[q]);
          GPx[q] = a1 * dXR[q] + a2 * dXZ[q] + a3 * dXT[q];
          GPy[q] = a1 * dYR[q] + a2 * dYZ[q] + a3 * dYT[q];
          GPz[q] = a1 * dZR[q] + a2 * dZZ[q] + a3 * dZT[q];
          continue;
        }
        const double t1 = (dPR[q] * GRS[q] + dPZ[q] * GRGZ[q] + dPT[q] * GRGT[q]);
        const double t2 = (dPR[q] * GRGZ[q] + dPZ[q] * GZS[q] + dPT[q] * GTGZ[q]);
        const double t3 = (dPR[q] * GRGT[q] + dPZ[q] * GTGZ[q] + dPT[q] * GTS[q]);
        GPx[q] = t1 * dXR[q] + t2 * dXZ[q] + t3 * dXT[q] + dDiff[q] * GRGT[q] * dXR[q] + dDiff[q] * GTGZ[q] * dXZ[q] + dDiff[q] * GTS[q] * dXT[q];
        GPy[q] = t1 * dYR[q] + t2 * dYZ[q] + t3 * dYT[q] + dDiff[q] * GRGT[q] * dYR[q] + dDiff[q] * GTGZ[q] * dYZ[q] + dDiff[q] * GTS[q] * dYT[q];
        GPz[q] = t1 * dZR[q] + t2 * dZZ[q] + t3 * dZT[q] + dDiff[q] * GRGT[q] * dZR[q] + dDiff[q] * GTGZ[q] * dZZ[q] + dDiff[q] * GTS[q] * dZT[q];
      }
  for (size_t q = 0; q < n; ++q) {
    jCrossB[q] = std::sqrt(jCBsq[q]) * bnormal * pjconst;
    GradP[q] = std::sqrt(std::fabs(gPsq[q])) * pnormal / 6.4;
  }
  double normDiff = 0, normJxB = 0, normGradP = 0, volume = 0;
  for (int i = 2; i <= nthe - 1; ++i)
    for (int j = 2; j <= npsi - 1; ++j)
      for (int k = 2; k <= nzeta; ++k) {
        const size_t q = &X3(jac, i, j, k) - jac;
        normDiff = normDiff + jac[q] * s.dr * s.dpPrime * s.dt * (jCrossB[q] - GradP[q]);
        normJxB = normJxB + jac[q] * s.dr * s.dpPrime * s.dt * jCrossB[q];
        normGradP = normGradP + jac[q] * s.dr * s.dpPrime * s.dt * GradP[q];
        volume = volume + jac[q] * s.dr * s.dpPrime * s.dt;
      }
  normDiff /= volume; normJxB /= volume; normGradP /= volume;
  o->s["normDiff"] = normDiff; o->s["normJxB"] = normJxB; o->s["normGradP"] = normGradP; o->s["volume"] = volume;
  return (std::isnan(normDiff) || std::isnan(normJxB) || std::isnan(normGradP)) ? 1 : 0;
}

}  // namespace

// ---- `pressure`, 2-D front end (src/ModScbRun.f90:838-1086, anisotropic branch with RAM pressures) ----------------------------
// Part 1 (independent of the SCB geometry): the RAM pressures of the species with species%SCB summed on the RAM grid
// (:858-875), the radial extension of the grid (:877-881) and of the pressures (PressMode SKD / ROE / EXT / FLT, :884-938), the
// smoothing (iSm2: 1 SavGol7, 3 Gaussian, 4 both; :947-980 with SavGol7 of src/ModScbFunctions.f90:137-230 and gaussian_kernel /
// convolve of srcExternal/gaussian_filter.f90).  Outputs rad2(nXRawExt) = radRawExt**2, azim(nAz), per / par(nXRawExt, nAz).
namespace {
void gaussian_kernel9(double* w);                                                   // defined with the computehI tail below
void convolve9(int rows, int cols, const double* in, const double* w, double* out);
void savgol7(int nrad, int nphi, int iters, std::vector<double>& pres) {
  static const double vals[49] = {32, 5, 1, -2, -2, -1, 5, 15, 4, 3, 3, 1, 0, -3, 3, 3, 4, 6, 3, 1, -6, -4, 2, 4, 7, 4, 2, -4,
                                  -6, 1, 3, 6, 4, 3, 3, -3, 0, 1, 3, 3, 4, 15, 5, -1, -2, -2, 1, 5, 32};
  auto B = [&](int r, int c) { return vals[(c - 1) * 7 + (r - 1)]; };       // RESHAPE fills by columns
  auto P = [&](std::vector<double>& a, int j, int k) -> double& { return a[(size_t)(j - 1) + (size_t)nrad * (k - 1)]; };
  double mn = pres[0];
  for (double v : pres) mn = v < mn ? v : mn;                                // MINVAL(pres)
  std::vector<double> pres0 = pres, pres1(pres.size()), pres2(pres.size());
  for (int it = 0; it < iters; ++it) {
    for (int k = 1; k <= nphi; ++k)
      for (int j = 1; j <= nrad; ++j) {
        if (j > 3 && j < nrad - 2) {
          double sum = 0.0;
          for (int m = 1; m <= 7; ++m) sum += (B(4, m) / 21.) * P(pres0, j - 4 + m, k);
          P(pres1, j, k) = sum;
        } else if (j <= 3) {
          P(pres1, j, k) = P(pres0, j, k);
        } else {
          const int row = (j == nrad - 2) ? 5 : ((j == nrad - 1) ? 6 : 7);
          const double den = (row == 7) ? 42. : 14.;
          double sum = 0.0;
          for (int m = 1; m <= 7; ++m) sum += (B(row, m) / den) * P(pres0, nrad - 7 + m, k);
          P(pres1, j, k) = sum;
        }
      }
    for (int j = 1; j <= nrad; ++j)
      for (int k = 1; k <= nphi; ++k) {
        int idx[7];
        if (k > 3 && k < nphi - 2) { for (int m = 0; m < 7; ++m) idx[m] = k - 3 + m; }
        else if (k == 1 || k == nphi) { const int t[7] = {nphi - 3, nphi - 2, nphi - 1, 1, 2, 3, 4}; for (int m = 0; m < 7; ++m) idx[m] = t[m]; }
        else if (k == 2) { const int t[7] = {nphi - 2, nphi - 1, 1, 2, 3, 4, 5}; for (int m = 0; m < 7; ++m) idx[m] = t[m]; }
        else if (k == nphi - 1) { const int t[7] = {nphi - 4, nphi - 3, nphi - 2, nphi - 1, 1, 2, 3}; for (int m = 0; m < 7; ++m) idx[m] = t[m]; }
        else if (k == 3) { const int t[7] = {nphi - 1, 1, 2, 3, 4, 5, 6}; for (int m = 0; m < 7; ++m) idx[m] = t[m]; }
        else { const int t[7] = {nphi - 5, nphi - 4, nphi - 3, nphi - 2, nphi - 1, 1, 2}; for (int m = 0; m < 7; ++m) idx[m] = t[m]; }
        double sum = 0.0;
        for (int m = 1; m <= 7; ++m) sum += (B(4, m) / 21.) * P(pres1, j, idx[m - 1]);
        P(pres2, j, k) = sum;
      }
    pres0 = pres2;
  }
  pres = pres2;
  for (double& v : pres) if (v < 0) v = mn;
}
}  // namespace

extern "C" {
// returns 0, or 1 for an unsupported PressMode / iSm2
int scbo_pressure_raw(int nS, int NR, int NT, const double* PPerT, const double* PParT, const int* scb, const double* LZ,
                      const double* PHI, int PressMode, int iSm2, int SavGolIters, double* rad2, double* azim, double* per, double* par) {
  const int nXRaw = NR - 1, nAz = NT, nXRawExt = NR + 2 * (int)std::floor(1.5 / (5. / NR));
  std::vector<double> radRaw(nXRaw), radExt(nXRawExt);
  for (int j1 = 1; j1 <= nXRaw; ++j1) radRaw[j1 - 1] = LZ[j1];                       // LZ(j1+1)
  for (int k1 = 1; k1 <= nAz; ++k1) azim[k1 - 1] = ((PHI[k1 - 1] * 12 / PI_D) * 360. / 24) * PI_D / 180.;
  for (int j1 = 1; j1 <= nXRaw; ++j1) radExt[j1 - 1] = radRaw[j1 - 1];
  for (int j1 = nXRaw + 1; j1 <= nXRawExt; ++j1)
    radExt[j1 - 1] = radRaw[nXRaw - 1] + (double)(j1 - nXRaw) * (radRaw[nXRaw - 1] - radRaw[0]) / ((double)(nXRaw - 1));
  std::vector<double> pe((size_t)nXRawExt * nAz, 0.0), pa((size_t)nXRawExt * nAz, 0.0);
  auto E = [&](std::vector<double>& a, int j, int k) -> double& { return a[(size_t)(j - 1) + (size_t)nXRawExt * (k - 1)]; };
  for (int k1 = 1; k1 <= nAz; ++k1)
    for (int j1 = 1; j1 <= nXRaw; ++j1) {
      double sp = 0.0, sa = 0.0;
      for (int iS = 1; iS <= nS; ++iS)
        if (scb[iS - 1]) {
          sp = sp + PPerT[(size_t)(iS - 1) + (size_t)nS * (j1 + (size_t)NR * (k1 - 1))];   // PPerT(iS, j1+1, k1)
          sa = sa + PParT[(size_t)(iS - 1) + (size_t)nS * (j1 + (size_t)NR * (k1 - 1))];
        }
      E(pe, j1, k1) = 0.16 * sp;                                                      // keV/cm^3 -> nPa
      E(pa, j1, k1) = 0.16 * sa;
    }
  auto R = [&](int j) { return radExt[j - 1]; };
  for (int k1 = 1; k1 <= nAz; ++k1) {
    if (PressMode == 0) {                                                              // SKD (:884-897)
      const double den = 89. * std::exp(-0.59 * R(nXRaw - 2)) + 8.9 * std::pow(R(nXRaw - 2), -1.53);
      for (int j1 = nXRaw - 1; j1 <= nXRawExt; ++j1) {
        const double num = 89. * std::exp(-0.59 * R(j1)) + 8.9 * std::pow(R(j1), -1.53);
        E(pe, j1, k1) = E(pe, nXRaw - 2, k1) * num / den;
        E(pa, j1, k1) = E(pa, nXRaw - 2, k1) * num / den;
      }
    } else if (PressMode == 1) {                                                       // ROE (:898-909), pRoeRad ModScbFunctions.f90:118-134
      auto roe = [](double r) { return 8.4027 * std::exp(-1.7845 * r) + 90.3150 * std::exp(-0.7659 * r); };
      for (int j1 = nXRaw + 1; j1 <= nXRawExt; ++j1) {
        E(pe, j1, k1) = E(pe, nXRaw, k1) * roe(R(j1)) / roe(R(nXRaw));
        E(pa, j1, k1) = E(pa, nXRaw, k1) * roe(R(j1)) / roe(R(nXRaw));
      }
    } else if (PressMode == 2) {                                                       // EXT (:910-925)
      for (int j1 = nXRaw + 1; j1 <= nXRawExt; ++j1) {
        E(pe, j1, k1) = E(pe, j1 - 1, k1) + (R(j1) - R(j1 - 1)) / (R(j1 - 2) - R(j1 - 1)) * (E(pe, j1 - 2, k1) - E(pe, j1 - 1, k1));
        E(pa, j1, k1) = E(pa, j1 - 1, k1) + (R(j1) - R(j1 - 1)) / (R(j1 - 2) - R(j1 - 1)) * (E(pa, j1 - 2, k1) - E(pa, j1 - 1, k1));
      }
    } else if (PressMode == 3) {                                                       // FLT (:926-938)
      for (int j1 = nXRaw + 1; j1 <= nXRawExt - 1; ++j1) { E(pe, j1, k1) = E(pe, nXRaw, k1); E(pa, j1, k1) = E(pa, nXRaw, k1); }
      E(pe, nXRawExt, k1) = 0.0;
      E(pa, nXRawExt, k1) = 0.0;
    } else return 1;
  }
  if (iSm2 == 1 || iSm2 == 4) { savgol7(nXRawExt, nAz, SavGolIters, pe); savgol7(nXRawExt, nAz, SavGolIters, pa); }
  if (iSm2 == 3 || iSm2 == 4) {
    double w[81];
    gaussian_kernel9(w);
    std::vector<double> out(pe.size());
    convolve9(nXRawExt, nAz, pe.data(), w, out.data()); pe = out;
    convolve9(nXRawExt, nAz, pa.data(), w, out.data()); pa = out;
  } else if (iSm2 != 0 && iSm2 != 1) return 1;                                         // 2 = B-spline fit (GSL bspline): not restated
  for (int j1 = 0; j1 < nXRawExt; ++j1) rad2[j1] = radExt[j1] * radExt[j1];
  for (size_t q = 0; q < pe.size(); ++q) { per[q] = pe[q]; par[q] = pa[q]; }
  return 0;
}

// Part 2, per call of `pressure`: the equatorial foot points of the SCB lines -> radius / angle (:838-850), bilinear
// interpolation of the extended RAM pressures in (r**2, azimuth) (GSL_Interpolation_2D -> gsl_interp2d_bilinear with
// eval_extrap, :1060-1065), extap inside 2 RE (:1069-1076), the <= 0 floor, periodic columns, normalisation (:1077-1086).
// xEq, yEq, pperEq, pparEq: (npsi, nzeta+1).
void scbo_pressure_eq(int npsi, int nzeta, const double* xEq, const double* yEq, int nX, int nAz, const double* rad2, const double* azim,
                      const double* per, const double* par, double pnormal, double* pperEq, double* pparEq) {
  auto Q = [&](const double* a, int j, int k) { return a[(size_t)(j - 1) + (size_t)npsi * (k - 1)]; };
  auto W = [&](double* a, int j, int k) -> double& { return a[(size_t)(j - 1) + (size_t)npsi * (k - 1)]; };
  auto bsearch = [](const double* xa, double x, int n) {
    int ilo = 0, ihi = n - 1;
    while (ihi > ilo + 1) { const int i = (ihi + ilo) / 2; if (xa[i] > x) ihi = i; else ilo = i; }
    return ilo;
  };
  auto bilin = [&](const double* za, double x, double y) {
    const int xi = bsearch(rad2, x, nX), yi = bsearch(azim, y, nAz);
    const double xmin = rad2[xi], xmax = rad2[xi + 1], ymin = azim[yi], ymax = azim[yi + 1];
    const double zminmin = za[(size_t)yi * nX + xi], zminmax = za[(size_t)(yi + 1) * nX + xi];
    const double zmaxmin = za[(size_t)yi * nX + xi + 1], zmaxmax = za[(size_t)(yi + 1) * nX + xi + 1];
    const double dx = xmax - xmin, dy = ymax - ymin;
    const double t = (x - xmin) / dx, u = (y - ymin) / dy;
    return (1. - t) * (1. - u) * zminmin + t * (1. - u) * zmaxmin + (1. - t) * u * zminmax + t * u * zmaxmax;
  };
  std::vector<double> radGrid((size_t)npsi * (nzeta + 1), 0.0);
  for (size_t q = 0; q < (size_t)npsi * (nzeta + 1); ++q) { pperEq[q] = 0.0; pparEq[q] = 0.0; }
  for (int k = 2; k <= nzeta; ++k)
    for (int j = 1; j <= npsi; ++j) {
      const double xe = Q(xEq, j, k), ye = Q(yEq, j, k);
      const double radius = std::sqrt(xe * xe + ye * ye);
      double angle = std::asin(ye / radius) + PI_D;
      if (xe <= 0 && ye >= 0) angle = 2.0 * PI_D - std::asin(ye / radius);
      if (xe <= 0 && ye <= 0) angle = -std::asin(ye / radius);
      radGrid[(size_t)(j - 1) + (size_t)npsi * (k - 1)] = radius;
      W(pperEq, j, k) = bilin(per, radius * radius, angle);
      W(pparEq, j, k) = bilin(par, radius * radius, angle);
    }
  for (int k = 1; k <= nzeta; ++k)
    for (int j = 10; j >= 1; --j)
      if (radGrid[(size_t)(j - 1) + (size_t)npsi * (k - 1)] < 2.0) {
        extap(Q(pperEq, j + 3, k), Q(pperEq, j + 2, k), Q(pperEq, j + 1, k), W(pperEq, j, k));
        extap(Q(pparEq, j + 3, k), Q(pparEq, j + 2, k), Q(pparEq, j + 1, k), W(pparEq, j, k));
      }
  for (size_t q = 0; q < (size_t)npsi * (nzeta + 1); ++q) {
    if (pperEq[q] <= 0.0) pperEq[q] = 1e-1 / pnormal;
    if (pparEq[q] <= 0.0) pparEq[q] = 1e-1 / pnormal;
  }
  for (int j = 1; j <= npsi; ++j) {
    W(pperEq, j, nzeta + 1) = Q(pperEq, j, 2); W(pparEq, j, nzeta + 1) = Q(pparEq, j, 2);
    W(pperEq, j, 1) = Q(pperEq, j, nzeta); W(pparEq, j, 1) = Q(pparEq, j, nzeta);
  }
  for (size_t q = 0; q < (size_t)npsi * (nzeta + 1); ++q) { pperEq[q] = pperEq[q] / pnormal; pparEq[q] = pparEq[q] / pnormal; }
}
}  // extern "C"

extern "C" {
void* scbo_create(int nthe, int npsi, int nzeta) {
  Scb* o = new Scb();
  o->nthe = nthe; o->npsi = npsi; o->nzeta = nzeta;
  o->iv["nimax"] = 5001;        // src/ModScbMain.f90:45
  o->iv["theChange"] = 4;       // src/ModScbParams.f90
  o->iv["psiChange"] = 0;
  o->iv["isotropy"] = 0;
  o->iv["iLossCone"] = 1;           // src/ModScbParams.f90:76
  o->iv["iReduceAnisotropy"] = 0;   // :52
  o->s["InConAlpha"] = 1e-6;    // src/ModScbParams.f90:37-38
  o->s["InConPsi"] = 1e-6;
  const double xzero3 = 6.6 * 6.6 * 6.6;   // src/ModScbInit.f90:246-273
  const double bnormal = 0.31 / xzero3 * 1.E5;
  o->s["bnormal"] = bnormal;
  o->s["pnormal"] = bnormal * bnormal / (4. * PI_D * 1.E-7) * 1.E-9;
  o->s["pjconst"] = 1.e6 * 0.31E-4 / (xzero3 * 4. * PI_D * 1.E-7 * 6.4E6);
  return o;
}
void scbo_destroy(void* h) { delete (Scb*)h; }
void scbo_set_array(void* h, const char* n, double* p) { ((Scb*)h)->d[n] = p; }
void scbo_set_scalar(void* h, const char* n, double v) { ((Scb*)h)->s[n] = v; }
void scbo_set_int(void* h, const char* n, int v) { ((Scb*)h)->iv[n] = v; }
double scbo_get_scalar(void* h, const char* n) { return ((Scb*)h)->S(n); }
int scbo_bandjacob(void* h) { return computeBandJacob((Scb*)h); }
void scbo_metrica(void* h) { metrica((Scb*)h); }
void scbo_metric(void* h) { metric((Scb*)h); }
void scbo_newk(void* h) { newk((Scb*)h); }
void scbo_newj(void* h) { newj((Scb*)h); }
int scbo_iterate_alpha(void* h, int* ni) { return iterateAlpha((Scb*)h, ni); }
int scbo_iterate_psi(void* h, int* ni) { return iteratePsi((Scb*)h, ni); }
int scbo_convergence(void* h) { return compute_convergence((Scb*)h); }
void scbo_derivs3d(void* h, const double* f, double* dT, double* dR, double* dZ) { derivs3d((Scb*)h, f, dT, dR, dZ); }
void scbo_pressure_aniso(void* h) { pressure_aniso((Scb*)h); }
int scbo_map_alpha(void* h) { return mapAlpha((Scb*)h); }
int scbo_map_psi(void* h) { return mapPsi((Scb*)h); }
int scbo_map_theta(void* h) { return mapTheta((Scb*)h); }
int scbo_interp1d(int n0, const double* x1, const double* f1, int n2, const double* x2, double* f2) { return interp1d(n0, x1, f1, n2, x2, f2); }
void scbo_steffen(int n, const double* xa, const double* ya, double* dx) {
  std::vector<double> yp;
  steffen_derivs(n, xa, ya, dx, yp);
}
}

// ---- computehI, everything after the integral block (src/ModRamScb.f90:413-632) -------------------------------
// outer-boundary scaling (:413-470), MLT continuity (:472-476), near-90-degree corrections (:478-487), negative and
// "too large" repairs (:489-514), Steffen interpolation of h, I onto PAbn (:516-529), Gaussian smoothing
// (:539-563 with gaussian_kernel / convolve of srcExternal/gaussian_filter.f90:19-56, :98-187; 8-byte reals under the
// reference's -fdefault-real-8), the RAM variables and their time derivatives (:566-606), the I = 1 row (:607-622) and the
// NaN repair (:625-637).  EIR / EIP(1,J) = 0 (:611-612) are E-field arrays and stay with the caller.  All arrays in
// the reference's shapes, Fortran order; ScaleAt holds 1-based radial indices (0 = line inside the SCB domain).
namespace {
inline size_t h3(int i, int j, int L, int n1, int n2) { return (size_t)i + (size_t)n1 * (j + (size_t)n2 * L); }   // 0-based

void gaussian_kernel9(double* w) {     // sigma = 1.0, truncate 4 -> radius 4
  const int radius = (int)(4 * 1.0 + 0.5);
  const double s = 1.0 * 1.0;
  double sum = 0.0;
  for (int j = -radius; j <= radius; ++j)
    for (int i = -radius; i <= radius; ++i) {
      const double x = i, y = j;
      const double v = 2.0 * std::exp(-0.5 * (x * x + y * y) / s);
      w[(i + radius) + 9 * (j + radius)] = v;
      sum += v;                         // SUM(kernel): array element order
    }
  for (int q = 0; q < 81; ++q) w[q] = w[q] / sum;
}

// convolve without mask: 3x3 reflected tiling, output(i,j) = sum(weights * overlapping) in array element order
void convolve9(int rows, int cols, const double* in, const double* w, double* out) {
  auto refl = [](int p, int n) { return p < 0 ? -1 - p : (p >= n ? 2 * n - 1 - p : p); };   // 0-based mirror with edge repeat
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i) {
      double sum = 0.0;
      for (int dj = -4; dj <= 4; ++dj)
        for (int di = -4; di <= 4; ++di) sum += w[(di + 4) + 9 * (dj + 4)] * in[refl(i + di, rows) + (size_t)rows * refl(j + dj, cols)];
      out[i + (size_t)rows * j] = sum;
    }
}
}  // namespace

extern "C" {
void scbo_gaussian_kernel9(double* w) { gaussian_kernel9(w); }

int scbo_hi_tail(int nR, int nT, int nPa, double* I_cart, double* H_cart, double* D_cart, double* bZEq, const int* ScaleAt,
                 const int* outsideMGNP, const double* Lz, const double* PA, const double* PAbn, int integral_smooth, double DthI,
                 double* FNHS, double* FNIS, double* BOUNHS, double* BOUNIS, double* HDNS, double* BNES, double* dIdt, double* dHdt,
                 double* dIbndt, double* dBdt, double* h_interp_out, double* I_interp_out) {
#define C3(a, i, j, L) a[h3((i) - 1, (j) - 1, (L) - 1, nR, nT)]
#define R3(a, i, j, L) a[h3((i) - 1, (j) - 1, (L) - 1, nR + 1, nT)]
#define C2(a, i, j) a[((i) - 1) + (size_t)nR * ((j) - 1)]
#define R2(a, i, j) a[((i) - 1) + (size_t)(nR + 1) * ((j) - 1)]
  const size_t n3 = (size_t)nR * nT * nPa;
  std::vector<double> hI(n3, 0.0), iI(n3, 0.0);
  double scalingI = 0, scalingH = 0, scalingD = 0;
  for (int j = 2; j <= nT; ++j) {                                                               // :417-470
    if (ScaleAt[j - 1] == 0) continue;
    const int ii = ScaleAt[j - 1];
    for (int L = 2; L <= nPa; ++L) {
      const double fr = (Lz[ii - 1] - Lz[ii - 2]) / (Lz[ii - 3] - Lz[ii - 2]);
      const double I_Temp = C3(I_cart, ii - 1, j, L) + fr * (C3(I_cart, ii - 2, j, L) - C3(I_cart, ii - 1, j, L));
      scalingI = (I_Temp <= 0) ? C3(I_cart, ii - 1, j, L) / C3(I_cart, ii, j, L) : I_Temp / C3(I_cart, ii, j, L);
      const double H_Temp = C3(H_cart, ii - 1, j, L) + fr * (C3(H_cart, ii - 2, j, L) - C3(H_cart, ii - 1, j, L));
      scalingH = (H_Temp <= 0) ? C3(H_cart, ii - 1, j, L) / C3(H_cart, ii, j, L) : H_Temp / C3(H_cart, ii, j, L);
      const double D_Temp = C3(D_cart, ii - 1, j, L) + fr * (C3(D_cart, ii - 2, j, L) - C3(D_cart, ii - 1, j, L));
      scalingD = (D_Temp <= 0) ? C3(D_cart, ii - 1, j, L) / C3(D_cart, ii, j, L) : D_Temp / C3(D_cart, ii, j, L);
      for (int i = ii; i <= nR; ++i) {
        if (C2(outsideMGNP, i, j) == 0) {
          C3(I_cart, i, j, L) = C3(I_cart, i, j, L) * scalingI;
          C3(H_cart, i, j, L) = C3(H_cart, i, j, L) * scalingH;
          C3(D_cart, i, j, L) = C3(D_cart, i, j, L) * scalingD;
          C2(bZEq, i, j) = C2(bZEq, i - 1, j);
        } else {
          C3(I_cart, i, j, L) = C3(I_cart, i - 1, j, L);
          C3(H_cart, i, j, L) = C3(H_cart, i - 1, j, L);
          C3(D_cart, i, j, L) = C3(D_cart, i - 1, j, L);
          C2(bZEq, i, j) = C2(bZEq, i - 1, j);
        }
      }
    }
  }
  for (int i = 1; i <= nR; ++i) {                                                               // :472-476
    for (int L = 1; L <= nPa; ++L) {
      C3(I_cart, i, 1, L) = C3(I_cart, i, nT, L);
      C3(H_cart, i, 1, L) = C3(H_cart, i, nT, L);
      C3(D_cart, i, 1, L) = C3(D_cart, i, nT, L);
    }
    C2(bZEq, i, 1) = C2(bZEq, i, nT);
  }
  for (int j = 1; j <= nT; ++j)                                                                 // :478-487
    for (int i = 1; i <= nR; ++i) {
      C3(I_cart, i, j, 3) = 0.50 * C3(I_cart, i, j, 4);
      C3(I_cart, i, j, 2) = 0.20 * C3(I_cart, i, j, 3);
      C3(I_cart, i, j, 1) = 0.0;
      C3(H_cart, i, j, 3) = 0.99 * C3(H_cart, i, j, 4);
      C3(H_cart, i, j, 2) = 0.99 * C3(H_cart, i, j, 3);
      C3(H_cart, i, j, 1) = 0.99 * C3(H_cart, i, j, 2);
      C3(D_cart, i, j, 3) = 0.999 * C3(D_cart, i, j, 4);
      C3(D_cart, i, j, 2) = 0.999 * C3(D_cart, i, j, 3);
      C3(D_cart, i, j, 1) = 0.999 * C3(D_cart, i, j, 2);
    }
  bool neg = false;                                                                             // :489-508
  for (size_t q = 0; q < n3; ++q) neg = neg || H_cart[q] < 0.0 || I_cart[q] < 0.0 || D_cart[q] < 0.0;
  if (neg)
    for (int j = 1; j <= nT; ++j)
      for (int i = 2; i <= nR; ++i)
        for (int L = 1; L <= nPa; ++L) {
          if (C3(H_cart, i, j, L) < 0) C3(H_cart, i, j, L) = C3(H_cart, i - 1, j, L);
          if (C3(I_cart, i, j, L) < 0) C3(I_cart, i, j, L) = C3(I_cart, i - 1, j, L);
          if (C3(D_cart, i, j, L) < 0) C3(D_cart, i, j, L) = C3(D_cart, i - 1, j, L);
        }
  for (int i = 1; i <= nR; ++i)                                                                 // :509-517
    for (int j = 1; j <= nT; ++j)
      for (int L = nPa - 1; L >= 1; --L) {
        if (C3(I_cart, i, j, L) > C3(I_cart, i, j, L + 1)) C3(I_cart, i, j, L) = 0.99 * C3(I_cart, i, j, L + 1);
        if (C3(H_cart, i, j, L) > C3(H_cart, i, j, L + 1)) C3(H_cart, i, j, L) = 0.99 * C3(H_cart, i, j, L + 1);
        if (C3(D_cart, i, j, L) > C3(D_cart, i, j, L + 1)) C3(D_cart, i, j, L) = 0.999 * C3(D_cart, i, j, L + 1);
      }
  {                                                                                             // :519-532
    std::vector<double> xa(nPa), fh(nPa), fi(nPa), xb(nPa - 2), oh(nPa - 2), oi(nPa - 2);
    for (int q = 0; q < nPa; ++q) xa[q] = PA[nPa - 1 - q];                                      // PA(NPA:1:-1)
    for (int q = 0; q < nPa - 2; ++q) xb[q] = PAbn[nPa - 2 - q];                                // PAbn(NPA-1:2:-1)
    for (int j = 1; j <= nT; ++j)
      for (int i = 1; i <= nR; ++i) {
        for (int q = 0; q < nPa; ++q) { fh[q] = C3(H_cart, i, j, nPa - q); fi[q] = C3(I_cart, i, j, nPa - q); }
        if (interp1d(nPa, xa.data(), fh.data(), nPa - 2, xb.data(), oh.data())) return 1;
        if (interp1d(nPa, xa.data(), fi.data(), nPa - 2, xb.data(), oi.data())) return 1;
        for (int q = 0; q < nPa - 2; ++q) { C3(hI, i, j, nPa - 1 - q) = oh[q]; C3(iI, i, j, nPa - 1 - q) = oi[q]; }
        C3(hI, i, j, nPa) = C3(hI, i, j, nPa - 1);
        C3(iI, i, j, nPa) = C3(iI, i, j, nPa - 1);
        C3(hI, i, j, 1) = C3(hI, i, j, 2);
        C3(iI, i, j, 1) = C3(iI, i, j, 2);
      }
  }
  if (integral_smooth) {                                                                        // :539-563
    double w[81];
    gaussian_kernel9(w);
    std::vector<double> out((size_t)nR * nT);
    double* arrs[5] = {H_cart, I_cart, hI.data(), iI.data(), D_cart};
    for (int L = 2; L <= nPa; ++L)
      for (double* a : arrs) {
        double* pl = a + (size_t)nR * nT * (L - 1);
        convolve9(nR, nT, pl, w, out.data());
        std::copy(out.begin(), out.end(), pl);
      }
  }
  for (int I = 2; I <= nR + 1; ++I)                                                             // :566-606
    for (int J = 1; J <= nT; ++J) {
      const double BNESPrev = R2(BNES, I, J);
      for (int L = 1; L <= nPa; ++L) {
        const double FNISPrev = R3(FNIS, I, J, L), FNHSPrev = R3(FNHS, I, J, L);
        const double BOUNISPrev = R3(BOUNIS, I, J, L), BOUNHSPrev = R3(BOUNHS, I, J, L);
        R3(FNHS, I, J, L) = C3(H_cart, I - 1, J, L);
        R3(FNIS, I, J, L) = C3(I_cart, I - 1, J, L);
        R2(BNES, I, J) = C2(bZEq, I - 1, J);
        R3(HDNS, I, J, L) = C3(D_cart, I - 1, J, L);
        R3(BOUNHS, I, J, L) = C3(hI, I - 1, J, L);
        R3(BOUNIS, I, J, L) = C3(iI, I - 1, J, L);
        if (std::fabs(DthI) <= 1e-9) {
          R3(dIdt, I, J, L) = 0; R3(dHdt, I, J, L) = 0; R3(dIbndt, I, J, L) = 0;
        } else {
          R3(dIdt, I, J, L) = (R3(FNIS, I, J, L) - FNISPrev) / DthI;
          R3(dHdt, I, J, L) = (R3(FNHS, I, J, L) - FNHSPrev) / DthI;
          R3(dIbndt, I, J, L) = (R3(BOUNIS, I, J, L) - BOUNISPrev) / DthI;
        }
      }
      R2(BNES, I, J) = R2(BNES, I, J) / 1e9;
      R2(dBdt, I, J) = (std::fabs(DthI) <= 1e-9) ? 0.0 : (R2(BNES, I, J) - BNESPrev) / DthI;
    }
  for (int J = 1; J <= nT; ++J) {                                                               // :607-622
    R2(BNES, 1, J) = 0.32 / (Lz[0] * Lz[0] * Lz[0]) / 1.e4;
    R2(dBdt, 1, J) = 0.0;
    for (int L = 1; L <= nPa; ++L) {
      R3(FNHS, 1, J, L) = R3(FNHS, 2, J, L);
      R3(FNIS, 1, J, L) = R3(FNIS, 2, J, L);
      R3(BOUNHS, 1, J, L) = R3(BOUNHS, 2, J, L);
      R3(BOUNIS, 1, J, L) = R3(BOUNIS, 2, J, L);
      R3(HDNS, 1, J, L) = R3(HDNS, 2, J, L);
      R3(dIdt, 1, J, L) = 0; R3(dHdt, 1, J, L) = 0; R3(dIbndt, 1, J, L) = 0;
    }
  }
  for (int i = 2; i <= nR + 1; ++i)                                                             // :625-637
    for (int j = 1; j <= nT; ++j)
      for (int L = 1; L <= nPa; ++L) {
        if (std::isnan(R3(FNIS, i, j, L))) R3(FNIS, i, j, L) = R3(FNIS, i - 1, j, L);
        if (std::isnan(R3(FNHS, i, j, L))) R3(FNHS, i, j, L) = R3(FNHS, i - 1, j, L);
        if (std::isnan(R3(BOUNIS, i, j, L))) R3(BOUNIS, i, j, L) = R3(BOUNIS, i - 1, j, L);
        if (std::isnan(R3(BOUNHS, i, j, L))) R3(BOUNHS, i, j, L) = R3(BOUNHS, i - 1, j, L);
        if (std::isnan(R3(HDNS, i, j, L))) R3(HDNS, i, j, L) = R3(HDNS, i - 1, j, L);
        if (std::isnan(R3(dIdt, i, j, L))) R3(dIdt, i, j, L) = 0;
        if (std::isnan(R3(dIbndt, i, j, L))) R3(dIbndt, i, j, L) = 0;
      }
  if (h_interp_out) std::copy(hI.begin(), hI.end(), h_interp_out);
  if (I_interp_out) std::copy(iI.begin(), iI.end(), I_interp_out);
  return 0;
#undef C3
#undef R3
#undef C2
#undef R2
}
}
