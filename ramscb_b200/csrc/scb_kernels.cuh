// SCB hot-path kernels for sm_100a (FP64).  Compiled with -fmad=false: every
// expression keeps the reference's operation order, so the pointwise kernels
// (Steffen derivatives, computeBandJacob, metrica/metric, newk/newj) and the
// lexicographic SOR are bit-identical to the CPU oracle.
//
// Layout: the reference's Fortran arrays are mirrored as they are, theta
// fastest: a(i,j,k) -> a[i + nthe*(j + npsi*k)] (0-based), so consecutive
// threads walk theta and every access is coalesced.  Arrays with a periodic
// ghost plane have nzeta+1 planes, the others nzeta (src/ModScbInit.f90:22-119).
//
// Reference routines restated: src/ModScbCompute.f90:412-754,
// src/ModScbEquation.f90:18-665, src/ModScbEuler.f90:160-299,469-612,
// src/ModScbFunctions.f90:57-76, src/RamGSL.c:228-291 (GSL steffen spline).
#pragma once
#include <cuda_runtime.h>

struct ScbDev {
  int nthe, npsi, nzeta;
  // spacing constants, src/ModScbInit.f90:131-148
  double rdr, rdt, rdp, rdrsq, rdtsq, rdpsq, rdr2, rdt2, rdp2, rdr4, rdt4, rdp4, rdpdt4, rdtdr4, dr, dt, dpPrime;
  const double *thetaVal, *rhoVal, *zetaVal, *f, *fzet;
  double *x, *y, *z, *alfa, *psi, *pper, *ppar, *sigma, *bsq, *bf;  // (nthe,npsi,nzeta+1)
  double *dXT, *dXR, *dXZ, *dYT, *dYR, *dYZ, *dZT, *dZR, *dZZ, *jac;
  double *gRX, *gRY, *gRZ, *gZX, *gZY, *gZZ, *gTX, *gTY, *gTZ;
  double *GRS, *GTS, *GZS, *GRGT, *GRGZ, *GTGZ, *Bx, *By, *Bz;
  double *vecd, *vec1, *vec2, *vec3, *vec4, *vec6, *vec7, *vec8, *vec9, *vecx, *vecr;
  double *dPT, *dPR, *dPZ, *dBT, *dBR, *dBZ, *dPP, *dPA, *dBP, *dBA, *dPdAlpha, *dPdPsi;
  double *jGR, *jGZ, *jGT, *Jx, *Jy, *Jz, *GPx, *GPy, *GPz, *jCrossB, *GradP;
  double *w1, *w2, *w3, *w4, *w5;   // scratch (nthe,npsi,nzeta)
};

#define S3(a, i, j, k) (a)[(size_t)(i) + (size_t)d.nthe * ((size_t)(j) + (size_t)d.npsi * (size_t)(k))]   /* 0-based */

__device__ __forceinline__ double sq(double x) { return x * x; }

// ---- GSL steffen spline: derivative at node `i` of the data ya(0..n-1) on the
// abscissae xa.  ya is addressed as base[idx*stride].
__device__ __forceinline__ double steffen_sgn(double y) { return (y < 0) ? -1.0 : 1.0; }   // steffen_copysign(1.0, y)
__device__ __forceinline__ double steffen_yp(const double* xa, const double* base, size_t stride, int n, int i) {
  if (i == 0) return (base[stride] - base[0]) / (xa[1] - xa[0]);
  if (i == n - 1) return (base[(size_t)(n - 1) * stride] - base[(size_t)(n - 2) * stride]) / (xa[n - 1] - xa[n - 2]);
  const double hi = xa[i + 1] - xa[i];
  const double him1 = xa[i] - xa[i - 1];
  const double yi = base[(size_t)i * stride];
  const double si = (base[(size_t)(i + 1) * stride] - yi) / hi;
  const double sim1 = (yi - base[(size_t)(i - 1) * stride]) / him1;
  const double pi = (sim1 * hi + si * him1) / (him1 + hi);
  const double m1 = fabs(si) < 0.5 * fabs(pi) ? fabs(si) : 0.5 * fabs(pi);
  const double m2 = fabs(sim1) < m1 ? fabs(sim1) : m1;
  return (steffen_sgn(sim1) + steffen_sgn(si)) * m2;
}
__device__ __forceinline__ double steffen_node(const double* xa, const double* base, size_t stride, int n, int i) {
  double dx;
  if (i < n - 1) {
    dx = steffen_yp(xa, base, stride, n, i);
  } else {
    // last node: evaluated through the cubic of interval n-2 at delx = h (src/RamGSL.c:274-275)
    const int q = n - 2;
    const double hi = xa[q + 1] - xa[q];
    const double si = (base[(size_t)(q + 1) * stride] - base[(size_t)q * stride]) / hi;
    const double ypq = steffen_yp(xa, base, stride, n, q), ypq1 = steffen_yp(xa, base, stride, n, q + 1);
    const double a = (ypq + ypq1 - 2 * si) / hi / hi;
    const double b = (3 * si - 2 * ypq - ypq1) / hi;
    const double c = ypq;
    const double delx = xa[n - 1] - xa[q];
    dx = c + delx * (2.0 * b + delx * 3.0 * a);
  }
  if (dx == 0.0) dx = dx + 1e-31;
  return dx;
}

// GSL_Derivs of one field: any of the three outputs may be NULL.
__global__ void __launch_bounds__(128) k_scb_derivs(ScbDev d, const double* __restrict__ f, double* __restrict__ dT,
                                                    double* __restrict__ dR, double* __restrict__ dZ) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;
  if (i >= d.nthe) return;
  const size_t sj = d.nthe, sk = (size_t)d.nthe * d.npsi;
  const size_t q = i + sj * j + sk * k;
  if (dT) dT[q] = steffen_node(d.thetaVal, f + sj * j + sk * k, 1, d.nthe, i);
  if (dR) dR[q] = steffen_node(d.rhoVal, f + i + sk * k, sj, d.npsi, j);
  if (dZ) dZ[q] = steffen_node(d.zetaVal, f + i + sj * j, sk, d.nzeta, k);
}

// computeBandJacob (src/ModScbCompute.f90:435-483): derivatives of x,y,z, Jacobian,
// contravariant gradients, metric products, B (k >= 2).  One thread per point.
__global__ void __launch_bounds__(128) k_scb_bandjacob(ScbDev d, int* __restrict__ fail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;   // k = 0..nzeta-1
  if (i >= d.nthe) return;
  const size_t sj = d.nthe, sk = (size_t)d.nthe * d.npsi;
  const size_t q = i + sj * j + sk * k;
  const size_t oT = sj * j + sk * k, oR = i + sk * k, oZ = i + sj * j;
  const double xT = steffen_node(d.thetaVal, d.x + oT, 1, d.nthe, i), xR = steffen_node(d.rhoVal, d.x + oR, sj, d.npsi, j),
               xZ = steffen_node(d.zetaVal, d.x + oZ, sk, d.nzeta, k);
  const double yT = steffen_node(d.thetaVal, d.y + oT, 1, d.nthe, i), yR = steffen_node(d.rhoVal, d.y + oR, sj, d.npsi, j),
               yZ = steffen_node(d.zetaVal, d.y + oZ, sk, d.nzeta, k);
  const double zT = steffen_node(d.thetaVal, d.z + oT, 1, d.nthe, i), zR = steffen_node(d.rhoVal, d.z + oR, sj, d.npsi, j),
               zZ = steffen_node(d.zetaVal, d.z + oZ, sk, d.nzeta, k);
  d.dXT[q] = xT; d.dXR[q] = xR; d.dXZ[q] = xZ;
  d.dYT[q] = yT; d.dYR[q] = yR; d.dYZ[q] = yZ;
  d.dZT[q] = zT; d.dZR[q] = zR; d.dZZ[q] = zZ;
  const double jac = xR * (yZ * zT - yT * zZ) + xZ * (yT * zR - yR * zT) + xT * (yR * zZ - yZ * zR);
  d.jac[q] = jac;
  const double gRX = (yZ * zT - yT * zZ) / jac, gRY = (zZ * xT - zT * xZ) / jac, gRZ = (xZ * yT - xT * yZ) / jac;
  const double gZX = (yT * zR - yR * zT) / jac, gZY = (zT * xR - zR * xT) / jac, gZZ = (xT * yR - xR * yT) / jac;
  const double gTX = (yR * zZ - yZ * zR) / jac, gTY = (zR * xZ - zZ * xR) / jac, gTZ = (xR * yZ - xZ * yR) / jac;
  d.gRX[q] = gRX; d.gRY[q] = gRY; d.gRZ[q] = gRZ;
  d.gZX[q] = gZX; d.gZY[q] = gZY; d.gZZ[q] = gZZ;
  d.gTX[q] = gTX; d.gTY[q] = gTY; d.gTZ[q] = gTZ;
  const double GRS = gRX * gRX + gRY * gRY + gRZ * gRZ;
  const double GRGZ = gRX * gZX + gRY * gZY + gRZ * gZZ;
  const double GRGT = gRX * gTX + gRY * gTY + gRZ * gTZ;
  const double GTS = gTX * gTX + gTY * gTY + gTZ * gTZ;
  const double GTGZ = gTX * gZX + gTY * gZY + gTZ * gZZ;
  const double GZS = gZX * gZX + gZY * gZY + gZZ * gZZ;
  d.GRS[q] = GRS; d.GRGZ[q] = GRGZ; d.GRGT[q] = GRGT; d.GTS[q] = GTS; d.GTGZ[q] = GTGZ; d.GZS[q] = GZS;
  if (k >= 1) {   // Fortran k = 2..nzeta
    const double ff = d.f[j] * d.fzet[k];
    d.Bx[q] = (ff * xT / jac);
    d.By[q] = (ff * yT / jac);
    d.Bz[q] = (ff * zT / jac);
    const double bsq = (GRS * GZS - sq(GRGZ)) * sq(ff);
    d.bsq[q] = bsq;
    const double bf = sqrt(bsq);
    d.bf[q] = bf;
    if (isnan(bf)) *fail = 1;
  }
}
// bX(:,:,1) = bX(:,:,nZeta) etc. (:490-494)
__global__ void k_scb_bwrap(ScbDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= d.nthe) return;
  const size_t q0 = i + (size_t)d.nthe * j, qn = q0 + (size_t)d.nthe * d.npsi * (d.nzeta - 1);
  d.Bx[q0] = d.Bx[qn]; d.By[q0] = d.By[qn]; d.Bz[q0] = d.Bz[qn]; d.bf[q0] = d.bf[qn]; d.bsq[q0] = d.bsq[qn];
}

// ---- one stencil position of metrica / metric (src/ModScbEquation.f90:154-250) ----------
struct Pos {
  double aj, grs, gps, gts, grgp, gpgt, gtgr;
};
__device__ __forceinline__ Pos geom(double xt, double yt, double zt, double xp, double yp, double zp, double xr, double yr,
                                    double zr) {
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

// difference stencils (0-based i,j,k of the centre), a = x, y or z
#define V(a, di, dj, dk) S3(a, i + (di), j + (dj), k + (dk))
#define DT_A(a) ((V(a, 1, 0, 0) - V(a, 0, 0, 0)) * d.rdt)
#define DT_C(a) ((V(a, 0, 0, 0) - V(a, -1, 0, 0)) * d.rdt)
#define DP_A(a) ((V(a, 1, 0, 1) + V(a, 0, 0, 1) - V(a, 1, 0, -1) - V(a, 0, 0, -1)) * d.rdp4)
#define DP_C(a) ((V(a, 0, 0, 1) + V(a, -1, 0, 1) - V(a, 0, 0, -1) - V(a, -1, 0, -1)) * d.rdp4)
#define DR_A(a) ((V(a, 1, 1, 0) + V(a, 0, 1, 0) - V(a, 1, -1, 0) - V(a, 0, -1, 0)) * d.rdr4)
#define DR_C(a) ((V(a, 0, 1, 0) + V(a, -1, 1, 0) - V(a, 0, -1, 0) - V(a, -1, -1, 0)) * d.rdr4)
// metrica b,d = (k +- 1/2)
#define AT_B(a) ((V(a, 1, 0, 1) + V(a, 1, 0, 0) - V(a, -1, 0, 1) - V(a, -1, 0, 0)) * d.rdt4)
#define AT_D(a) ((V(a, 1, 0, 0) + V(a, 1, 0, -1) - V(a, -1, 0, 0) - V(a, -1, 0, -1)) * d.rdt4)
#define AP_B(a) ((V(a, 0, 0, 1) - V(a, 0, 0, 0)) * d.rdp)
#define AP_D(a) ((V(a, 0, 0, 0) - V(a, 0, 0, -1)) * d.rdp)
#define AR_B(a) ((V(a, 0, 1, 1) + V(a, 0, 1, 0) - V(a, 0, -1, 1) - V(a, 0, -1, 0)) * d.rdr4)
#define AR_D(a) ((V(a, 0, 1, -1) + V(a, 0, 1, 0) - V(a, 0, -1, -1) - V(a, 0, -1, 0)) * d.rdr4)
// metric b,d = (j +- 1/2)
#define MT_B(a) ((V(a, 1, 1, 0) + V(a, 1, 0, 0) - V(a, -1, 1, 0) - V(a, -1, 0, 0)) * d.rdt4)
#define MT_D(a) ((V(a, 1, 0, 0) + V(a, 1, -1, 0) - V(a, -1, 0, 0) - V(a, -1, -1, 0)) * d.rdt4)
#define MP_B(a) ((V(a, 0, 1, 1) + V(a, 0, 0, 1) - V(a, 0, 1, -1) - V(a, 0, 0, -1)) * d.rdp4)
#define MP_D(a) ((V(a, 0, 0, 1) + V(a, 0, -1, 1) - V(a, 0, 0, -1) - V(a, 0, -1, -1)) * d.rdp4)
#define MR_B(a) ((V(a, 0, 1, 0) - V(a, 0, 0, 0)) * d.rdr)
#define MR_D(a) ((V(a, 0, 0, 0) - V(a, 0, -1, 0)) * d.rdr)

// metrica (ALPHA=true, src/ModScbEquation.f90:18-280) / metric (ALPHA=false, :283-540).
// One thread per interior point j=2..npsi-1, k=2..nzeta, i=2..nthe-1 (Fortran); the
// 27-point neighbourhood of x,y,z is served by L1/L2.
template <bool ALPHA>
__global__ void __launch_bounds__(128) k_scb_metric(ScbDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;   // 1..nthe-2
  const int j = blockIdx.y + 1;                              // 1..npsi-2
  const int k = blockIdx.z + 1;                              // 1..nzeta-1
  if (i > d.nthe - 2) return;
  const double *x = d.x, *y = d.y, *z = d.z;
  const Pos a = geom(DT_A(x), DT_A(y), DT_A(z), DP_A(x), DP_A(y), DP_A(z), DR_A(x), DR_A(y), DR_A(z));
  const Pos c = geom(DT_C(x), DT_C(y), DT_C(z), DP_C(x), DP_C(y), DP_C(z), DR_C(x), DR_C(y), DR_C(z));
  const size_t q = (size_t)i + (size_t)d.nthe * ((size_t)j + (size_t)d.npsi * (size_t)k);
  if (ALPHA) {
    const Pos b = geom(AT_B(x), AT_B(y), AT_B(z), AP_B(x), AP_B(y), AP_B(z), AR_B(x), AR_B(y), AR_B(z));
    const Pos e = geom(AT_D(x), AT_D(y), AT_D(z), AP_D(x), AP_D(y), AP_D(z), AR_D(x), AR_D(y), AR_D(z));
    const double v1a = (a.grs * a.gts - sq(a.gtgr)) * a.aj * d.rdtsq;
    const double v1c = (c.grs * c.gts - sq(c.gtgr)) * c.aj * d.rdtsq;
    const double v2a = (a.grs * a.gpgt - a.grgp * a.gtgr) * a.aj * d.rdpdt4;
    const double v2b = (b.grs * b.gpgt - b.grgp * b.gtgr) * b.aj * d.rdpdt4;
    const double v2c = (c.grs * c.gpgt - c.grgp * c.gtgr) * c.aj * d.rdpdt4;
    const double v2d = (e.grs * e.gpgt - e.grgp * e.gtgr) * e.aj * d.rdpdt4;
    const double v3b = (b.grs * b.gps - sq(b.grgp)) * b.aj * d.rdpsq;
    const double v3d = (e.grs * e.gps - sq(e.grgp)) * e.aj * d.rdpsq;
    d.vecd[q] = (v1a + v1c) + (v3b + v3d);
    d.vec1[q] = (v2c + v2d);
    d.vec2[q] = (v2c - v2a) + v3d;
    d.vec3[q] = -(v2a + v2d);
    d.vec4[q] = v1c + (v2d - v2b);
    d.vec6[q] = v1a + (v2b - v2d);
    d.vec7[q] = -(v2c + v2b);
    d.vec8[q] = v3b + (v2a - v2c);
    d.vec9[q] = (v2a + v2b);
  } else {
    const Pos b = geom(MT_B(x), MT_B(y), MT_B(z), MP_B(x), MP_B(y), MP_B(z), MR_B(x), MR_B(y), MR_B(z));
    const Pos e = geom(MT_D(x), MT_D(y), MT_D(z), MP_D(x), MP_D(y), MP_D(z), MR_D(x), MR_D(y), MR_D(z));
    const double v1a = (sq(a.gpgt) - a.gps * a.gts) * a.aj * d.rdtsq;
    const double v1c = (sq(c.gpgt) - c.gps * c.gts) * c.aj * d.rdtsq;
    const double v2a = (a.grgp * a.gpgt - a.gps * a.gtgr) * a.aj * d.rdtdr4;
    const double v2b = (b.grgp * b.gpgt - b.gps * b.gtgr) * b.aj * d.rdtdr4;
    const double v2c = (c.grgp * c.gpgt - c.gps * c.gtgr) * c.aj * d.rdtdr4;
    const double v2d = (e.grgp * e.gpgt - e.gps * e.gtgr) * e.aj * d.rdtdr4;
    const double v3b = (sq(b.grgp) - b.grs * b.gps) * b.aj * d.rdrsq;
    const double v3d = (sq(e.grgp) - e.grs * e.gps) * e.aj * d.rdrsq;
    d.vecd[q] = (v1a + v1c + v3b + v3d);
    d.vec1[q] = (v2c + v2d);
    d.vec2[q] = ((v2c - v2a) + v3d);
    d.vec3[q] = -(v2a + v2d);
    d.vec4[q] = (v1c + (v2d - v2b));
    d.vec6[q] = (v1a + (v2b - v2d));
    d.vec7[q] = -(v2c + v2b);
    d.vec8[q] = (v3b + (v2a - v2c));
    d.vec9[q] = (v2b + v2a);
  }
}

// newk (ALPHA) / newj, Picard, src/ModScbEquation.f90:546-665.  All points of (nthe,npsi,nzeta).
template <bool ALPHA>
__global__ void __launch_bounds__(128) k_scb_rhs(ScbDev d, int isotropy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;
  if (i >= d.nthe) return;
  const size_t q = (size_t)i + (size_t)d.nthe * ((size_t)j + (size_t)d.npsi * (size_t)k);
  const double fj = d.f[j], fk = d.fzet[k];
  if (isotropy == 1) {
    if (ALPHA) d.vecx[q] = -d.dPdAlpha[q] * d.jac[q] / (fj * fj);
    else d.vecr[q] = d.jac[q] * d.dPdPsi[q] / (fk * fk);
    return;
  }
  const double sg = d.sigma[q];
  const double tt = d.dPT[q] + 0.5 * (1. - sg) * d.dBT[q];
  if (ALPHA) {
    const double xpz = (d.GRS[q] * d.GZS[q] - sq(d.GRGZ[q]));
    const double xpt = (d.GRS[q] * d.GTGZ[q] - d.GRGZ[q] * d.GRGT[q]);
    const double c0 = -(fj * fj * fk) / sg / d.bsq[q];
    const double tz = d.dPZ[q] + 0.5 * (1. - sg) * d.dBZ[q];
    d.vecx[q] = d.jac[q] / (fj * fj) * c0 * (tz * xpz + tt * xpt);
  } else {
    const double xpr = (sq(d.GRGZ[q]) - d.GRS[q] * d.GZS[q]);
    const double xpt = (d.GRGZ[q] * d.GTGZ[q] - d.GZS[q] * d.GRGT[q]);
    const double c0 = -(fj * (fk * fk)) / sg / d.bsq[q];
    const double tr = d.dPR[q] + 0.5 * (1. - sg) * d.dBR[q];
    d.vecr[q] = d.jac[q] / (fk * fk) * c0 * (tr * xpr + tt * xpt);
  }
}

// =============================================================================
// SOR.  Each sub-problem (a psi surface for alpha, a zeta slice for psi) is a 2-D
// 9-point problem owned by one CTA; the unknown plane lives in shared memory,
// the 10 coefficient planes are streamed from L2 (the whole coefficient set of
// the default grid is 39 MB: L2 resident).  The CTA iterates to convergence in
// one launch (in-kernel convergence test, per-sub-problem early exit).
//
// Plane indexing: column c = theta (0..nthe-1), row r = the second stencil
// direction (zeta planes 0..nzeta for alpha, psi 0..npsi-1 for psi).
//   ORDER 0 (LEX): the reference's lexicographic Gauss-Seidel order reproduced
//     as a skewed wavefront -- thread = row, at step t row r updates column
//     t-2r -- bit-identical iterates, iteration counts and residuals.
//   ORDER 1 (COLOR4): 4-colour ordering ((c mod 2, r mod 2)); the 9-point
//     stencil has corner couplings, so 4 colours (not red-black) are needed for
//     a race-free Gauss-Seidel; converges to the same fixed point.
// =============================================================================
struct SorArgs {
  double tol, omegaOpt;
  int nimax, nT, nP;
  int sub0;         // first sub-problem of this launch (ranks sharding the independent sub-problems)
  int* ni;          // per sub-problem iteration count (Fortran ni(jz) / ni(k))
  double* resmax;   // per sub-problem max|resid| of the last sweep
  int* fail;
};

template <bool ALPHA, int ORDER>
__global__ void __launch_bounds__(1024) k_scb_sor(ScbDev d, SorArgs a) {
  extern __shared__ double su[];
  __shared__ double s_red[32];
  __shared__ int s_stop;
  const int nthe = d.nthe, npsi = d.npsi, nzeta = d.nzeta;
  const int tid = threadIdx.x, T = blockDim.x;
  // sub-problem and its plane geometry
  const int sub = a.sub0 + blockIdx.x;              // alpha: jz-2 (Fortran jz = 2..npsi-nP); psi: k-2 (k = 2..nzeta)
  const int nrows = ALPHA ? nzeta + 1 : npsi;       // rows held in shared memory
  const int r0 = 1, r1 = ALPHA ? nzeta - 1 : npsi - a.nP - 1;   // updated rows (0-based, inclusive)
  const int c0 = a.nT, c1 = nthe - a.nT - 1;                    // updated columns
  const size_t sj = nthe, sk = (size_t)nthe * npsi;
  double* u = ALPHA ? d.alfa : d.psi;
  const size_t base = ALPHA ? sj * (size_t)(sub + 1) : sk * (size_t)(sub + 1);   // plane origin (jz or k fixed)
  const size_t rstride = ALPHA ? sk : sj;                                        // between rows
  for (int q = tid; q < nrows * nthe; q += T) {
    const int r = q / nthe, c = q - r * nthe;
    su[q] = u[base + (size_t)r * rstride + c];
  }
  if (tid == 0) s_stop = 0;
  __syncthreads();
  const double* rhs = ALPHA ? d.vecx : d.vecr;
  double om = 1.0;
  int ni = 1;
  double lastmax = 0.0;
  bool failed = false;
  while (ni <= a.nimax) {
    double rmax = 0.0;
    if (ORDER == 0) {
      const int nr = r1 - r0 + 1, nc = c1 - c0 + 1;
      const int r = r0 + tid;
      const bool active = tid < nr;
      const size_t rowoff = base + (size_t)r * rstride;
      for (int t = 0; t < 2 * (nr - 1) + nc; ++t) {
        const int c = c0 + t - 2 * tid;
        if (active && c >= c0 && c <= c1) {
          const size_t q = rowoff + c;
          const double* um = su + (r - 1) * nthe + c;
          const double* uc = su + r * nthe + c;
          const double* up = su + (r + 1) * nthe + c;
          const double vd = d.vecd[q];
          const double res = -vd * uc[0] + d.vec1[q] * um[-1] + d.vec2[q] * um[0] + d.vec3[q] * um[1] + d.vec4[q] * uc[-1] +
                             d.vec6[q] * uc[1] + d.vec7[q] * up[-1] + d.vec8[q] * up[0] + d.vec9[q] * up[1] - rhs[q];
          double un = ALPHA ? (uc[0] + om * (res / vd)) : (uc[0] + om * res / vd);
          double rr = res;
          if (isnan(un) || un >= 1e10) {   // :226-240 / :539-553
            un = u[q];                      // global memory still holds the pre-solve value
            rr = 0.0;
            failed = true;
          }
          su[r * nthe + c] = un;
          if (c >= 1 && c <= nthe - 2) rmax = fmax(rmax, fabs(rr));
        }
        __syncthreads();
      }
    } else {
      for (int col = 0; col < 4; ++col) {
        const int pc = col & 1, pr = col >> 1;
        // points of this colour: c in [c0,c1] with c%2==pc, r in [r0,r1] with r%2==pr
        const int cs = c0 + (((c0 & 1) == pc) ? 0 : 1), rs = r0 + (((r0 & 1) == pr) ? 0 : 1);
        const int ncc = (c1 - cs) / 2 + 1, nrr = (r1 - rs) / 2 + 1;
        if (cs <= c1 && rs <= r1)
          for (int w = tid; w < ncc * nrr; w += T) {
            const int rr_ = w / ncc, cc = w - rr_ * ncc;
            const int r = rs + 2 * rr_, c = cs + 2 * cc;
            const size_t q = base + (size_t)r * rstride + c;
            const double* um = su + (r - 1) * nthe + c;
            const double* uc = su + r * nthe + c;
            const double* up = su + (r + 1) * nthe + c;
            const double vd = d.vecd[q];
            const double res = -vd * uc[0] + d.vec1[q] * um[-1] + d.vec2[q] * um[0] + d.vec3[q] * um[1] + d.vec4[q] * uc[-1] +
                               d.vec6[q] * uc[1] + d.vec7[q] * up[-1] + d.vec8[q] * up[0] + d.vec9[q] * up[1] - rhs[q];
            double un = ALPHA ? (uc[0] + om * (res / vd)) : (uc[0] + om * res / vd);
            double rr = res;
            if (isnan(un) || un >= 1e10) {
              un = u[q];
              rr = 0.0;
              failed = true;
            }
            su[r * nthe + c] = un;
            if (c >= 1 && c <= nthe - 2) rmax = fmax(rmax, fabs(rr));
          }
        __syncthreads();
      }
    }
    // block max of |resid| over the sweep, failure flag
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    if ((tid & 31) == 0) s_red[tid >> 5] = rmax;
    if (failed) s_stop = 1;
    __syncthreads();
    double m = 0.0;
    for (int q = 0; q < (T + 31) / 32; ++q) m = fmax(m, s_red[q]);
    lastmax = m;
    const int stop = s_stop;
    __syncthreads();
    if (stop) break;              // EXIT Iterations on failure (ni not advanced)
    om = a.omegaOpt;
    if (m < a.tol) break;         // converged
    ni = ni + 1;
  }
  for (int q = tid; q < nrows * nthe; q += T) {
    const int r = q / nthe, c = q - r * nthe;
    if (r >= r0 && r <= r1 && c >= c0 && c <= c1) u[base + (size_t)r * rstride + c] = su[q];
  }
  if (tid == 0) {
    a.ni[sub] = ni;
    a.resmax[sub] = lastmax;
    if (s_stop) *a.fail = 1;
  }
}

// =============================================================================
// iterateAlpha sharded along zeta (the north-star scheme): a rank owns the zeta planes
// (rows) [k0, k0+nk) of EVERY psi surface and holds rows k0-1 and k0+nk as halo.  The 4-colour sweep
// of k_scb_sor<true,1> is cut at the row parity: colours 0,1 touch even rows and read odd rows +
// the row itself, colours 2,3 the other way round -- so ONE halo exchange per HALF-sweep (the rows
// of that parity at the slab edges) keeps every rank's iterates bit-identical to the single-GPU
// 4-colour solve.  The unknown stays in global memory (L2): a launch is one half-sweep of all
// surfaces that are still iterating, one CTA per (row, surface); the per-surface residual maxima are
// accumulated with atomicMax on the bit pattern (non-negative doubles order like integers) into
// state[0..nsub), failures into state[nsub..2 nsub) -- the vector the ranks all-reduce with MAX.
// k_scb_zcommit then applies the loop control of src/ModScbEuler.f90:204-262 per surface.
// =============================================================================
struct ZArgs {
  double tol, om;
  int nT, k0, nk, nsub, nimax;
  double* state;        // [resmax accumulators | failure flags], 2*nsub
  int* done;            // per surface: finished (converged, failed or nimax reached)
  const double* u0;     // alfa before the solve (restored at a point that blows up, :226-240)
};

__global__ void __launch_bounds__(64) k_scb_zhalf(ScbDev d, ZArgs a, int parity) {
  __shared__ double s_red[2];
  __shared__ int s_fail;
  const int sub = blockIdx.y;
  if (a.done[sub]) return;
  const int rs = a.k0 + (((a.k0 & 1) == parity) ? 0 : 1);
  const int r = rs + 2 * (int)blockIdx.x;
  if (r >= a.k0 + a.nk) return;
  const int nthe = d.nthe, tid = threadIdx.x, T = blockDim.x;
  const int c0 = a.nT, c1 = nthe - a.nT - 1;
  const size_t sk = (size_t)nthe * d.npsi;
  const size_t rowoff = (size_t)nthe * (size_t)(sub + 1) + (size_t)r * sk;
  double* u = d.alfa;
  if (tid == 0) s_fail = 0;
  double rmax = 0.0;
  bool failed = false;
  for (int pc = 0; pc < 2; ++pc) {
    const int cs = c0 + (((c0 & 1) == pc) ? 0 : 1);
    for (int c = cs + 2 * tid; c <= c1; c += 2 * T) {
      const size_t q = rowoff + c;
      const double* um = u + q - sk;
      const double* uc = u + q;
      const double* up = u + q + sk;
      const double vd = d.vecd[q];
      const double res = -vd * uc[0] + d.vec1[q] * um[-1] + d.vec2[q] * um[0] + d.vec3[q] * um[1] + d.vec4[q] * uc[-1] +
                         d.vec6[q] * uc[1] + d.vec7[q] * up[-1] + d.vec8[q] * up[0] + d.vec9[q] * up[1] - d.vecx[q];
      double un = uc[0] + a.om * (res / vd);
      double rr = res;
      if (isnan(un) || un >= 1e10) {
        un = a.u0[q];
        rr = 0.0;
        failed = true;
      }
      u[q] = un;
      if (c >= 1 && c <= nthe - 2) rmax = fmax(rmax, fabs(rr));
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = rmax;
  if (failed) s_fail = 1;
  __syncthreads();
  if (tid == 0) {
    double m = s_red[0];
    if (T > 32) m = fmax(m, s_red[1]);
    atomicMax((unsigned long long*)(a.state + sub), (unsigned long long)__double_as_longlong(m));
    if (s_fail) a.state[a.nsub + sub] = 1.0;
  }
}

// after the MAX all-reduce of `state`: per surface  IF fail EXIT (ni kept); IF max|resid| < tol EXIT;
// ni = ni + 1 (loop ends at ni > nimax); the accumulators are cleared for the next sweep
__global__ void k_scb_zcommit(ZArgs a, int* __restrict__ ni, double* __restrict__ resmax, int* __restrict__ fail,
                              int* __restrict__ pending) {
  __shared__ int s_pend;
  if (threadIdx.x == 0) s_pend = 0;
  __syncthreads();
  for (int sub = threadIdx.x; sub < a.nsub; sub += blockDim.x) {
    if (!a.done[sub]) {
      const double m = a.state[sub];
      resmax[sub] = m;
      if (a.state[a.nsub + sub] != 0.0) {
        a.done[sub] = 1;
        *fail = 1;
      } else if (m < a.tol) {
        a.done[sub] = 1;
      } else {
        const int n = ni[sub] + 1;
        ni[sub] = n;
        if (n > a.nimax) a.done[sub] = 1;
      }
      if (!a.done[sub]) atomicAdd(&s_pend, 1);
    }
    a.state[sub] = 0.0;
    a.state[a.nsub + sub] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) *pending = s_pend;
}

// equatorial foot points x(nThetaEquator,j,k), y(nThetaEquator,j,k) of every field line -> out[0..n2), out[n2..2 n2)
// (what `pressure` needs on the host to look the RAM pressures up, src/ModScbRun.f90:1001-1010)
__global__ void k_scb_gather_eq(ScbDev d, int ieq, double* __restrict__ out) {
  const size_t n2 = (size_t)d.npsi * (d.nzeta + 1);
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n2) return;
  out[q] = d.x[(size_t)ieq + (size_t)d.nthe * q];
  out[n2 + q] = d.y[(size_t)ieq + (size_t)d.nthe * q];
}

// sumb, sumdb over (2:nthe-1, 2:npsi-1, 2:nzeta) (Fortran), one CTA per zeta plane -> partials
__global__ void __launch_bounds__(256) k_scb_sums(ScbDev d, const double* __restrict__ u, const double* __restrict__ uprev,
                                                  double* __restrict__ part) {
  __shared__ double sm[2][32];
  const int k = blockIdx.x + 1;
  double sb = 0.0, sdb = 0.0;
  const int ni = d.nthe - 2, nj = d.npsi - 2;
  for (int w = threadIdx.x; w < ni * nj; w += blockDim.x) {
    const int jj = w / ni, ii = w - jj * ni;
    const size_t q = (size_t)(ii + 1) + (size_t)d.nthe * ((size_t)(jj + 1) + (size_t)d.npsi * (size_t)k);
    sb += fabs(u[q]);
    sdb += fabs(u[q] - uprev[q]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
    sdb += __shfl_xor_sync(0xffffffffu, sdb, o);
  }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = sb; sm[1][threadIdx.x >> 5] = sdb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { a += sm[0][q]; b += sm[1][q]; }
    part[2 * blockIdx.x] = a;
    part[2 * blockIdx.x + 1] = b;
  }
}

// =============================================================================
// Glue of the SCB outer iteration (src/ModScbRun.f90:232-262, 418-440) so that the 3-D arrays stay
// on the device between the solves: the blend of the new potential with the one saved at the
// start of the iteration,  u = uNew*blend + uSav*(1 - blend)  (:236, :422; two roundings per
// product and one for the sum, as the Fortran expression), and MINVAL(jacobian(2:nthe-1,
// 2:npsi-1, 2:nzeta)) (:248, :434) whose sign decides whether the moved points are kept.
// =============================================================================
__global__ void __launch_bounds__(256) k_scb_blend(double* __restrict__ u, const double* __restrict__ unew,
                                                   const double* __restrict__ usav, double blend, size_t n) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) u[q] = unew[q] * blend + usav[q] * (1.0 - blend);
}
// one CTA per zeta plane k = 2..nzeta -> part[k-2]; the host takes the min of nzeta-1 values
__global__ void __launch_bounds__(256) k_scb_minjac(ScbDev d, double* __restrict__ part) {
  __shared__ double sm[32];
  const int k = blockIdx.x + 1;
  const int ni = d.nthe - 2, nj = d.npsi - 2;
  double m = 1e300;
  bool nan = false;
  for (int w = threadIdx.x; w < ni * nj; w += blockDim.x) {
    const int jj = w / ni, ii = w - jj * ni;
    const double v = d.jac[(size_t)(ii + 1) + (size_t)d.nthe * ((size_t)(jj + 1) + (size_t)d.npsi * (size_t)k)];
    if (v != v) nan = true;
    m = v < m ? v : m;
  }
  if (nan) m = -1e300;          // a NaN Jacobian must not pass the `< 0` test as "fine"
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t < m ? t : m;
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < (int)(blockDim.x >> 5); ++q) m = sm[q] < m ? sm[q] : m;
    part[blockIdx.x] = m;
  }
}

// =============================================================================
// k_scb_press_aniso: the anisotropic branch of `pressure` from the normalised equatorial pressures
// on (src/ModScbRun.f90:1087-1160): pper, ppar mapped along the field line with the iLossCone = 1
// (filled loss cone, :1107-1110) or 2 (Liemohn 2004, :1098-1106) formulas, sigma, tau, and the
// optional reduction to marginal mirror stability (:1127-1160) of lines whose equatorial tau < 0
// (every thread of such a line re-derives the equatorial test itself: no second pass).
// bf, bsq come from computeBandJacob and never leave the device.  Thread per (i, j, k <= nzeta).
// The Steffen derivatives and the 1/f, 1/fzet scalings (:1162-1175) follow in k_scb_derivs /
// k_scb_press_scale.  Reference operation order; `1./6.` (:1133) is a single-precision constant.
// =============================================================================
struct AnisoPt { double pper, ppar, sigma, tau; };
__device__ __forceinline__ AnisoPt aniso_point(double pperEq, double pparEq, double bfEq, double bfI, double bf1, double bsqI,
                                               int iLossCone, bool clampRatio) {
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
    r.ppar = pparN * (aN + 1.) / (1. + aN * ratioB) * sqrt((rBI - 1.) / (rBI - ratioB)) * (1. - (1. + aN * ratioB) / (rBI + aN * ratioB));
    r.pper = r.ppar / (1. + aN * ratioB);
  } else {
    const double gParam = 1. / sq(1. + aratio * (1. - ratioB));
    r.ppar = pEq * 1. / (1. + 2. * aratio / 3.) * sqrt(gParam);
    r.pper = pEq * (aratio + 1.) / (1. + 2. * aratio / 3.) * gParam;
  }
  r.sigma = 1. + (r.pper - r.ppar) / bsqI;
  r.tau = 1. - 2. * (r.pper - r.ppar) / bsqI * r.pper / r.ppar;
  return r;
}
__global__ void __launch_bounds__(128) k_scb_press_aniso(ScbDev d, const double* __restrict__ pperEq, const double* __restrict__ pparEq,
                                                         double* __restrict__ tau, int iLossCone, int iReduce, int eq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;
  if (i >= d.nthe) return;
  const size_t e2 = (size_t)j + (size_t)d.npsi * k;
  const double pe0 = pperEq[e2], pa0 = pparEq[e2];
  const double bfEq = S3(d.bf, eq, j, k), bfI = S3(d.bf, i, j, k), bf1 = S3(d.bf, 0, j, k), bsqI = S3(d.bsq, i, j, k);
  AnisoPt r = aniso_point(pe0, pa0, bfEq, bfI, bf1, bsqI, iLossCone, true);
  if (iReduce == 1) {
    const double bEqSq = S3(d.bsq, eq, j, k);
    const AnisoPt q = aniso_point(pe0, pa0, bfEq, bfEq, bf1, bEqSq, iLossCone, true);
    if (q.tau < 0.) {
      const double pEq = (2. * pe0 + pa0) / 3.;
      const double sixth = (double)(1.f / 6.f);
      const double pe = sixth * (3. * pEq - bEqSq + sqrt(sq(bEqSq) + 12. * bEqSq * pEq + 9. * sq(pEq)));
      const double pa = 3. * pEq - 2. * pe;
      r = aniso_point(pe, pa, bfEq, bfI, bf1, bsqI, iLossCone, false);
      if (iLossCone == 1) {                              // :1149-1151 keep pEq = press(j,k) of the first pass
        const double aratio = pe / pa - 1.;
        const double ratioB = bfEq / bfI;
        const double gParam = 1. / sq(1. + aratio * (1. - ratioB));
        r.ppar = pEq * 1. / (1. + 2. * aratio / 3.) * sqrt(gParam);
        r.pper = pEq * (aratio + 1.) / (1. + 2. * aratio / 3.) * gParam;
        r.sigma = 1.0 + (r.pper - r.ppar) / bsqI;
        r.tau = 1. - 2. * (r.pper - r.ppar) / bsqI * r.pper / r.ppar;
      }
    }
  }
  S3(d.pper, i, j, k) = r.pper;
  S3(d.ppar, i, j, k) = r.ppar;
  S3(d.sigma, i, j, k) = r.sigma;
  S3(tau, i, j, k) = r.tau;
}
// dPperdPsi = 1/f(j) * dPperdRho, dPperdAlpha = 1/fzet(k) * dPperdZeta, same for bsq (:1168-1175)
__global__ void __launch_bounds__(128) k_scb_press_scale(ScbDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;
  if (i >= d.nthe) return;
  const double rf = 1. / d.f[j], rz = 1. / d.fzet[k];
  S3(d.dPP, i, j, k) = rf * S3(d.dPR, i, j, k);
  S3(d.dBP, i, j, k) = rf * S3(d.dBR, i, j, k);
  S3(d.dPA, i, j, k) = rz * S3(d.dPZ, i, j, k);
  S3(d.dBA, i, j, k) = rz * S3(d.dBZ, i, j, k);
}

// extap, src/ModScbFunctions.f90:57-76
__device__ __forceinline__ double extap(double x1, double x2, double x3) {
  double x4 = 3. * x3 - 3. * x2 + x1;
  const double ddx1 = x3 - x2, ddx2 = x2 - x1;
  double ddx = x4 - x3;
  const double pm = ddx * ddx1;
  if (pm > 0.) return x4;
  if (fabs(ddx2) <= 1e-9) return 2. * x3 - x2;
  ddx = (ddx1 * ddx1) / ddx2;
  return x3 + ddx;
}
// post-processing of iterateAlpha/iteratePsi (src/ModScbEuler.f90:262-292, :575-605), stage 1:
// extrapolation onto the outer psi surfaces; one thread per (i, k)
__global__ void k_scb_post_extap(ScbDev d, double* __restrict__ u, int nT, int nP) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0-based theta
  const int k = blockIdx.y + 1;                           // Fortran k = 2..nzeta
  if (i < nT || i > d.nthe - nT - 1) return;
  for (int j = nP; j >= 1; --j) {
    // extap(u(i,npsi-j-2), u(i,npsi-j-1), u(i,npsi-j), u(i,npsi-j+1)), Fortran 1-based psi index
    const int j4 = d.npsi - j + 1 - 1;
    S3(u, i, j4, k) = extap(S3(u, i, j4 - 3, k), S3(u, i, j4 - 2, k), S3(u, i, j4 - 1, k));
  }
}
// stage 2: linear fill of the theta ends for k = 1..nzeta (Fortran), all j
__global__ void k_scb_post_theta(ScbDev d, double* __restrict__ u, int nT) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;   // 0..nzeta-1
  if (j >= d.npsi) return;
  const int nthe = d.nthe;
  for (int i = 1; i <= nT; ++i) {   // Fortran i
    // u(i) = u(nT+1) + (nT+1-i)*(u(1)-u(nT+1))/nT ; u(nthe-i+1) = u(nthe-nT-1) + (nT+1-i)*(u(nthe)-u(nthe-nT-1))/nT
    S3(u, i - 1, j, k) = S3(u, nT, j, k) + (nT + 1 - i) * (S3(u, 0, j, k) - S3(u, nT, j, k)) / nT;
    S3(u, nthe - i, j, k) = S3(u, nthe - nT - 2, j, k) + (nT + 1 - i) * (S3(u, nthe - 1, j, k) - S3(u, nthe - nT - 2, j, k)) / (nT);
  }
}
// stage 3: periodic wrap u(:,:,1) = u(:,:,nzeta) - wrap ; u(:,:,nzeta+1) = u(:,:,2) + wrap
__global__ void k_scb_post_wrap(ScbDev d, double* __restrict__ u, double wrap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= d.nthe) return;
  S3(u, i, j, 0) = S3(u, i, j, d.nzeta - 1) - wrap;
  S3(u, i, j, d.nzeta) = S3(u, i, j, 1) + wrap;
}

// ---- Compute_convergence (src/ModScbCompute.f90:553-733), isotropy 0 and 1 ----------------
// stage 1: j.gradRho, j.gradZeta, and the two flux-like fields whose derivatives give j.gradTheta,
// plus jacobian*(pper-ppar)
__global__ void __launch_bounds__(128) k_scb_conv1(ScbDev d, int isotropy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;
  if (i >= d.nthe) return;
  const size_t q = (size_t)i + (size_t)d.nthe * ((size_t)j + (size_t)d.npsi * (size_t)k);
  const double sg = d.sigma[q], fj = d.f[j], fk = d.fzet[k];
  if (isotropy == 1) {                                         // :553-556, :575-578
    d.jGR[q] = -1.0 / fj * d.dPdAlpha[q];
    d.jGZ[q] = 1.0 / fk * d.dPdPsi[q];
  } else {
  d.jGR[q] = 1.0 / fj * (-1. / sg * d.dPA[q] -
                         1. / (sg * d.bsq[q]) * (fj * fj) * fk * (d.GRS[q] * d.GTGZ[q] - d.GRGT[q] * d.GRGZ[q]) *
                             (d.dPT[q] + (1. - sg) * 0.5 * d.dBT[q]) -
                         (1. - sg) / sg * 0.5 * d.dBA[q]);
  d.jGZ[q] = 1.0 / fk * (1. / sg * d.dPP[q] -
                         1. / (sg * d.bsq[q]) * fj * (fk * fk) * (d.GRGZ[q] * d.GTGZ[q] - d.GRGT[q] * d.GZS[q]) *
                             (d.dPT[q] + (1. - sg) * 0.5 * d.dBT[q]) +
                         (1. - sg) / sg * 0.5 * d.dBP[q]);
  }
  d.w1[q] = d.jac[q] * fj * fk * (d.GRGT[q] * d.GRGZ[q] - d.GTGZ[q] * d.GRS[q]);   // jGradThetaPartialRho
  d.w2[q] = d.jac[q] * fj * fk * (d.GRGT[q] * d.GZS[q] - d.GRGZ[q] * d.GTGZ[q]);   // jGradThetaPartialZeta
  d.w3[q] = d.jac[q] * (d.pper[q] - d.ppar[q]);
}
// stage 2 (after the derivative passes: w4 = d(w1)/drho, w5 = d(w2)/dzeta, w1 <- d(w3)/dtheta):
// J, grad P, |J x B|, |grad P| and the per-plane partial sums of the four norms
__global__ void __launch_bounds__(128) k_scb_conv2(ScbDev d, double bnormal, double pnormal, double pjconst, double* __restrict__ part,
                                                   int isotropy) {
  __shared__ double sm[4][4];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, k = blockIdx.z;
  double acc[4] = {0, 0, 0, 0};
  if (i < d.nthe) {
    const size_t q = (size_t)i + (size_t)d.nthe * ((size_t)j + (size_t)d.npsi * (size_t)k);
    const double jac = d.jac[q];
    const double jGT = (d.w4[q] + d.w5[q]) / jac;
    d.jGT[q] = jGT;
    const double jGR = d.jGR[q], jGZ = d.jGZ[q], dD = d.w1[q];
    d.Jx[q] = (jGR * d.dXR[q] + jGZ * d.dXZ[q] + jGT * d.dXT[q]);
    d.Jy[q] = (jGR * d.dYR[q] + jGZ * d.dYZ[q] + jGT * d.dYT[q]);
    d.Jz[q] = (jGR * d.dZR[q] + jGZ * d.dZZ[q] + jGT * d.dZT[q]);
    const double fj = d.f[j], fk = d.fzet[k];
    const double GRS = d.GRS[q], GZS = d.GZS[q], GTS = d.GTS[q], GRGZ = d.GRGZ[q], GRGT = d.GRGT[q], GTGZ = d.GTGZ[q];
    const double dPR = d.dPR[q], dPZ = d.dPZ[q], dPT = d.dPT[q];
    const double jCBsq = (fj * fj) * (fk * fk) * (GRS * sq(jGZ) + GZS * sq(jGR) - 2.0 * jGZ * jGR * GRGZ);
    const double gPsq = GRS * sq(dPR) + GZS * sq(dPZ) + GTS * sq(dPT) + 2. * dPR * dPZ * GRGZ + 2. * dPR * dPT * GRGT +
                        2. * dPZ * dPT * GTGZ + sq(dD / jac) - 2. * dPT * dD / jac;
    if (isotropy == 1) {                                       // :671-690
      const double pP = d.dPdPsi[q], pA = d.dPdAlpha[q];
      const double a1 = (fj * pP * GRS + fk * pA * GRGZ);
      const double a2 = (fj * pP * GRGZ + fk * pA * GZS);
      const double a3 = (fj * pP * GRGT + fk * pA * GTGZ);
      d.GPx[q] = a1 * d.dXR[q] + a2 * d.dXZ[q] + a3 * d.dXT[q];
      d.GPy[q] = a1 * d.dYR[q] + a2 * d.dYZ[q] + a3 * d.dYT[q];
      d.GPz[q] = a1 * d.dZR[q] + a2 * d.dZZ[q] + a3 * d.dZT[q];
    } else {
    const double t1 = (dPR * GRS + dPZ * GRGZ + dPT * GRGT);
    const double t2 = (dPR * GRGZ + dPZ * GZS + dPT * GTGZ);
    const double t3 = (dPR * GRGT + dPZ * GTGZ + dPT * GTS);
    d.GPx[q] = t1 * d.dXR[q] + t2 * d.dXZ[q] + t3 * d.dXT[q] + dD * GRGT * d.dXR[q] + dD * GTGZ * d.dXZ[q] + dD * GTS * d.dXT[q];
    d.GPy[q] = t1 * d.dYR[q] + t2 * d.dYZ[q] + t3 * d.dYT[q] + dD * GRGT * d.dYR[q] + dD * GTGZ * d.dYZ[q] + dD * GTS * d.dYT[q];
    d.GPz[q] = t1 * d.dZR[q] + t2 * d.dZZ[q] + t3 * d.dZT[q] + dD * GRGT * d.dZR[q] + dD * GTGZ * d.dZZ[q] + dD * GTS * d.dZT[q];
    }
    const double jCB = sqrt(jCBsq) * bnormal * pjconst;
    const double gP = sqrt(fabs(gPsq)) * pnormal / 6.4;
    d.jCrossB[q] = jCB;
    d.GradP[q] = gP;
    if (i >= 1 && i <= d.nthe - 2 && j >= 1 && j <= d.npsi - 2 && k >= 1) {
      const double vol = jac * d.dr * d.dpPrime * d.dt;
      acc[0] = vol * (jCB - gP);
      acc[1] = vol * jCB;
      acc[2] = vol * gP;
      acc[3] = vol;
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double v = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[m][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t cta = blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z);
    for (int m = 0; m < 4; ++m) {
      double v = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sm[m][w];
      part[cta * 4 + m] = v;
    }
  }
}

// =============================================================================
// k_scb_map<MODE>: mapAlpha (MODE 0, src/ModScbEuler.f90:97-147), mapPsi (1, :403-457) and
// mapTheta (2, :15-75): move the grid points x,y,z along one coordinate line so that the Euler
// potential just computed (alfa / psi) -- or the arc-length fraction of the field line -- takes
// its prescribed node values again.  Per line: GSL_Interpolation_1D with the Steffen spline
// (src/ModRamGSL.f90:240-311, src/RamGSL.c:111-174): drop non-increasing abscissae, Steffen node
// slopes, bisection for the interval, Horner evaluation, linear extrapolation outside the data.
// The three coordinates of a line share abscissae, interval search and the filter.
// One thread per line; the line's copies (xOld/yOld/zOld of the reference) and slopes live in a
// workspace W[7][n][nlines] (coalesced across lines).  The thread also does the periodic wrap of
// its line (:136-141) and alfges / psiges (:82-94, :387-400).  Operation order of the reference
// (-fmad=false): bit-identical to the oracle.
// =============================================================================
template <int MODE>
__global__ void __launch_bounds__(64) k_scb_map(ScbDev d, const double* __restrict__ tgt, double* __restrict__ W, int nlines,
                                               int* __restrict__ fail) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= nlines) return;
  const int nthe = d.nthe, npsi = d.npsi, nzeta = d.nzeta;
  const size_t sj = nthe, sk = (size_t)nthe * npsi;
  // line geometry: n source points from `base` with stride `st`; n2 targets written from point `o0` on
  size_t base, st;
  int n, n2, o0, kpl = 0;
  if (MODE == 0) { base = line; st = sk; n = nzeta + 1; n2 = nzeta - 1; o0 = 1; }                        // line = i + nthe*j
  else if (MODE == 1) { const int i = line % nthe; kpl = 1 + line / nthe; base = i + sk * kpl; st = sj; n = npsi; n2 = npsi; o0 = 0; }
  else { const int j = line % npsi; kpl = 1 + line / npsi; base = sj * j + sk * kpl; st = 1; n = nthe; n2 = nthe; o0 = 0; }
  const size_t NL = nlines;
  double* Wx = W + line;                    // abscissae
  double* Wf[3] = {W + (size_t)1 * n * NL + line, W + (size_t)2 * n * NL + line, W + (size_t)3 * n * NL + line};
  double* Wp[3] = {W + (size_t)4 * n * NL + line, W + (size_t)5 * n * NL + line, W + (size_t)6 * n * NL + line};
  double* xyz[3] = {d.x, d.y, d.z};
  const double* absc = (MODE == 0) ? d.alfa : d.psi;
  double total = 0.0;
  if (MODE == 2) {                          // arc length along the line (:43-46)
    double dist = 0.0;
    Wx[0] = 0.0;
    for (int q = 1; q < n; ++q) {
      const size_t o = base + q * st;
      dist = dist + sqrt(sq(d.x[o] - d.x[o - st]) + sq(d.y[o] - d.y[o - st]) + sq(d.z[o] - d.z[o - st]));
      Wx[(size_t)q * NL] = dist;
    }
    total = dist;
  }
  // copy the line, dropping abscissae that do not increase (src/ModRamGSL.f90:262-273)
  int n1 = 0;
  double last = 0.0;
  for (int q = 0; q < n; ++q) {
    const size_t o = base + q * st;
    const double xv = (MODE == 2) ? Wx[(size_t)q * NL] / total * 3.141592653589793238462643383279502884197 : absc[o];   // pi_d
    if (q == 0 || xv > last) {
      Wx[(size_t)n1 * NL] = xv;
#pragma unroll
      for (int c = 0; c < 3; ++c) Wf[c][(size_t)n1 * NL] = xyz[c][o];
      last = xv;
      ++n1;
    }
  }
  if (n1 < 3) { atomicAdd(fail, 1); return; }            // GSLerr > 0 => SORFail
  // Steffen node slopes (gsl interpolation/steffen.c, steffen_init)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double* fa = Wf[c];
    Wp[c][0] = (fa[NL] - fa[0]) / (Wx[NL] - Wx[0]);
    for (int i = 1; i < n1 - 1; ++i) {
      const double hi = Wx[(size_t)(i + 1) * NL] - Wx[(size_t)i * NL];
      const double him1 = Wx[(size_t)i * NL] - Wx[(size_t)(i - 1) * NL];
      const double si = (fa[(size_t)(i + 1) * NL] - fa[(size_t)i * NL]) / hi;
      const double sim1 = (fa[(size_t)i * NL] - fa[(size_t)(i - 1) * NL]) / him1;
      const double pi = (sim1 * hi + si * him1) / (him1 + hi);
      const double m1 = fabs(si) < 0.5 * fabs(pi) ? fabs(si) : 0.5 * fabs(pi);
      const double m2 = fabs(sim1) < m1 ? fabs(sim1) : m1;
      Wp[c][(size_t)i * NL] = (steffen_sgn(sim1) + steffen_sgn(si)) * m2;
    }
    Wp[c][(size_t)(n1 - 1) * NL] = (fa[(size_t)(n1 - 1) * NL] - fa[(size_t)(n1 - 2) * NL]) / (Wx[(size_t)(n1 - 1) * NL] - Wx[(size_t)(n1 - 2) * NL]);
  }
  // evaluate at the prescribed node values
  const double xa0 = Wx[0], xa1 = Wx[NL], xaN = Wx[(size_t)(n1 - 1) * NL], xaM = Wx[(size_t)(n1 - 2) * NL];
  bool bad = false;
  for (int q = 0; q < n2; ++q) {
    const double xb = tgt[(MODE == 0) ? q + 1 : q];
    const size_t o = base + (size_t)(o0 + q) * st;
    if (xb <= xa0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) xyz[c][o] = Wf[c][0] + (xb - xa0) / (xa1 - xa0) * (Wf[c][NL] - Wf[c][0]);
    } else if (xb >= xaN) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        xyz[c][o] = Wf[c][(size_t)(n1 - 1) * NL] + (xb - xaN) / (xaM - xaN) * (Wf[c][(size_t)(n1 - 2) * NL] - Wf[c][(size_t)(n1 - 1) * NL]);
    } else if (xb == xb) {
      int ilo = 0, ihi = n1 - 1;
      while (ihi > ilo + 1) {                            // gsl_interp_bsearch
        const int i = (ihi + ilo) / 2;
        if (Wx[(size_t)i * NL] > xb) ihi = i; else ilo = i;
      }
      const double xl = Wx[(size_t)ilo * NL];
      const double hi = Wx[(size_t)(ilo + 1) * NL] - xl;
      const double delx = xb - xl;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double f0 = Wf[c][(size_t)ilo * NL];
        const double si = (Wf[c][(size_t)(ilo + 1) * NL] - f0) / hi;
        const double y0 = Wp[c][(size_t)ilo * NL], y1 = Wp[c][(size_t)(ilo + 1) * NL];
        const double a = (y0 + y1 - 2 * si) / hi / hi;
        const double b = (3 * si - 2 * y0 - y1) / hi;
        xyz[c][o] = f0 + delx * (y0 + delx * (b + delx * a));
      }
    } else {
      bad = true;
    }
  }
  if (bad) { atomicAdd(fail, 1); return; }
  // periodic planes (zeta = 1 <- nzeta, nzeta+1 <- 2) and the reset of the potential
  if (MODE == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      xyz[c][base] = xyz[c][base + (size_t)(nzeta - 1) * st];
      xyz[c][base + (size_t)nzeta * st] = xyz[c][base + st];
    }
    for (int k = 0; k <= nzeta; ++k) d.alfa[base + (size_t)k * st] = tgt[k];          // alfges
  } else {
    const long long shift = (kpl == nzeta - 1) ? -(long long)sk * (nzeta - 1) : ((kpl == 1) ? (long long)sk * (nzeta - 1) : 0);
    for (int q = 0; q < n; ++q) {
      const size_t o = base + q * st;
      if (shift != 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) xyz[c][(size_t)((long long)o + shift)] = xyz[c][o];
      }
      if (MODE == 1) {                                                                // psiges
        d.psi[o] = tgt[q];
        if (shift != 0) d.psi[(size_t)((long long)o + shift)] = tgt[q];
      }
    }
  }
}

// =============================================================================
// k_scb_sor_cluster: the 4-colour SOR of iterateAlpha / iteratePsi with the WHOLE
// problem resident on chip.  A thread-block cluster of CL CTAs owns one independent
// sub-problem (a psi surface for alpha, a zeta plane for psi); the updated rows are cut
// into CL slabs.  Each CTA keeps in shared memory, for the life of the solve,
//   * its rows of the unknown plus one halo row on either side, and
//   * the nine stencil coefficients and the right-hand side of its points, packed
//     colour-major (unit-stride, conflict-free reads),
// i.e. after the initial load no sweep touches L2/HBM ("coefficients held in
// (distributed) shared memory for the whole solve").  After a colour phase the boundary-row
// updates are pushed into the neighbour CTA's halo row through distributed shared memory;
// two cluster barriers per sweep (at the row-parity switches of the colour sequence) make them
// visible; the per-sweep max|resid| and the failure flag travel the same way.  Per point the arithmetic and the colour order are those of
// k_scb_sor<.,1>: potentials, iteration counts and residual maxima are bit-identical.
// grid: nsub*CL CTAs, cluster (CL,1,1); smem: see scb_gpu.cu
// =============================================================================
#include <cooperative_groups.h>

template <bool ALPHA>
__global__ void __launch_bounds__(1024) k_scb_sor_cluster(ScbDev d, SorArgs a, int nloc_max, int npc_max) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ double sm[];
  __shared__ double s_red[32];
  __shared__ double s_xmax[8];   // per-rank max|resid| of the sweep (every rank's written by its owner)
  __shared__ int s_xfail[8];
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int nthe = d.nthe, npsi = d.npsi, nzeta = d.nzeta;
  const int tid = threadIdx.x, T = blockDim.x;
  const int sub = a.sub0 + blockIdx.x / CL;
  const int r0 = 1, r1 = ALPHA ? nzeta - 1 : npsi - a.nP - 1;   // updated rows (0-based, inclusive)
  const int c0 = a.nT, c1 = nthe - a.nT - 1;                    // updated columns
  const int nr = r1 - r0 + 1;
  const int nloc_nom = (nr + CL - 1) / CL;
  const int rb = r0 + rank * nloc_nom;                          // my first row
  const int nloc = max(0, min(nloc_nom, r1 - rb + 1));          // my rows: rb .. rb+nloc-1
  const size_t sj = nthe, sk = (size_t)nthe * npsi;
  double* u = ALPHA ? d.alfa : d.psi;
  const size_t base = ALPHA ? sj * (size_t)(sub + 1) : sk * (size_t)(sub + 1);
  const size_t rstride = ALPHA ? sk : sj;
  // shared layout: su[(nloc_max+2)][nthe] | coef[4 colours][10][npc_max]
  double* su = sm;
  double* sc = sm + (size_t)(nloc_max + 2) * nthe;
  // local row lr = 0 is global row rb-1, lr = 1..nloc my rows, lr = nloc+1 global row rb+nloc
  for (int q = tid; q < (nloc + 2) * nthe; q += T) {
    const int lr = q / nthe, c = q - lr * nthe;
    su[q] = u[base + (size_t)(rb - 1 + lr) * rstride + c];
  }
  const double* rhs = ALPHA ? d.vecx : d.vecr;
  const double* cf[10] = {d.vecd, d.vec1, d.vec2, d.vec3, d.vec4, d.vec6, d.vec7, d.vec8, d.vec9, rhs};
  int rs_[4], cs_[4], ncc_[4], np_[4];
  for (int col = 0; col < 4; ++col) {
    const int pc = col & 1, pr = col >> 1;
    const int cs = c0 + (((c0 & 1) == pc) ? 0 : 1);
    const int rs = rb + (((rb & 1) == pr) ? 0 : 1);             // global row parity, as the one-CTA kernel
    const int rend = rb + nloc - 1;
    const int ncc = (cs <= c1) ? (c1 - cs) / 2 + 1 : 0;
    const int nrr = (rs <= rend) ? (rend - rs) / 2 + 1 : 0;
    rs_[col] = rs; cs_[col] = cs; ncc_[col] = ncc; np_[col] = ncc * nrr;
    for (int w = tid; w < ncc * nrr; w += T) {
      const int rr = w / ncc, cc = w - rr * ncc;
      const size_t q = base + (size_t)(rs + 2 * rr) * rstride + (cs + 2 * cc);
#pragma unroll
      for (int m = 0; m < 10; ++m) sc[((size_t)col * 10 + m) * npc_max + w] = cf[m][q];
    }
  }
  if (tid < 8) { s_xmax[tid] = 0.0; s_xfail[tid] = 0; }
  cluster.sync();
  double* su_prev = (rank > 0) ? cluster.map_shared_rank(su, rank - 1) : nullptr;
  double* su_next = (rank < CL - 1) ? cluster.map_shared_rank(su, rank + 1) : nullptr;
  const int nloc_prev = nloc_nom;                               // every rank before the last has nloc_nom rows
  double om = 1.0;
  int ni = 1;
  double lastmax = 0.0;
  bool failed = false, stopped = false;
  while (ni <= a.nimax) {
    double rmax = 0.0;
    for (int col = 0; col < 4; ++col) {
      const int rs = rs_[col], cs = cs_[col], ncc = ncc_[col], np = np_[col];
      const double* k0 = sc + (size_t)col * 10 * npc_max;
      for (int w = tid; w < np; w += T) {
        const int rr = w / ncc, cc = w - rr * ncc;
        const int r = rs + 2 * rr, c = cs + 2 * cc;
        const int lr = r - rb + 1;
        const double* um = su + (lr - 1) * nthe + c;
        double* uc = su + lr * nthe + c;
        const double* up = su + (lr + 1) * nthe + c;
        const double vd = k0[w];
        const double res = -vd * uc[0] + k0[npc_max + w] * um[-1] + k0[2 * npc_max + w] * um[0] + k0[3 * npc_max + w] * um[1] +
                           k0[4 * npc_max + w] * uc[-1] + k0[5 * npc_max + w] * uc[1] + k0[6 * npc_max + w] * up[-1] +
                           k0[7 * npc_max + w] * up[0] + k0[8 * npc_max + w] * up[1] - k0[9 * npc_max + w];
        double un = ALPHA ? (uc[0] + om * (res / vd)) : (uc[0] + om * res / vd);
        double rr2 = res;
        if (isnan(un) || un >= 1e10) {   // :226-240 / :539-553
          un = u[base + (size_t)r * rstride + c];
          rr2 = 0.0;
          failed = true;
        }
        uc[0] = un;
        if (lr == 1 && su_prev) su_prev[(nloc_prev + 1) * nthe + c] = un;     // neighbour's lower halo row
        if (lr == nloc && su_next) su_next[c] = un;                           // neighbour's upper halo row
        if (c >= 1 && c <= nthe - 2) rmax = fmax(rmax, fabs(rr2));
      }
      // Colours 0,1 update the even rows, 2,3 the odd rows.  A point reads the neighbour CTA's
      // rows only through its halo row, whose parity is the opposite of the point's own row: the
      // halo values a colour needs were written two or three colours earlier.  So a CTA barrier
      // separates 0|1 and 2|3, and only the row-parity switches 1|2 and 3|0 need the cluster.
      if (col == 0 || col == 2) {
        __syncthreads();
      } else if (col == 1) {
        cluster.sync();
      } else {
        // last colour: CTA max of |resid| over the sweep + failure flag, published to every rank
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = rmax;
        const int anyfail = __syncthreads_or(failed ? 1 : 0);
        if (tid < CL) {
          double m = 0.0;
          for (int q = 0; q < (T + 31) / 32; ++q) m = fmax(m, s_red[q]);
          double* xm = cluster.map_shared_rank(s_xmax, tid);
          int* xf = cluster.map_shared_rank(s_xfail, tid);
          xm[rank] = m;
          xf[rank] = anyfail;
        }
        cluster.sync();
      }
    }
    double m = 0.0;
    int stop = 0;
    for (int q = 0; q < CL; ++q) { m = fmax(m, s_xmax[q]); stop |= s_xfail[q]; }
    lastmax = m;
    if (stop) { stopped = true; break; }   // EXIT Iterations on failure (ni not advanced)
    om = a.omegaOpt;
    if (m < a.tol) break;                  // converged
    ni = ni + 1;
  }
  for (int q = tid; q < nloc * nthe; q += T) {
    const int lr = 1 + q / nthe, c = q % nthe;
    if (c >= c0 && c <= c1) u[base + (size_t)(rb + lr - 1) * rstride + c] = su[lr * nthe + c];
  }
  if (rank == 0 && tid == 0) {
    a.ni[sub] = ni;
    a.resmax[sub] = lastmax;
    if (stopped) *a.fail = 1;
  }
  cluster.sync();   // no CTA may exit while a neighbour can still write into its shared memory
}

// =============================================================================
// k_scb_sor_cluster_reg: the same cluster solve with the coefficients in REGISTERS.  When a
// CTA has at most one point per colour per thread (<= 576 points per colour), thread t owns
// the t-th point of each of the four colours for the whole solve: its 4 x 10 coefficients
// stay in 80 registers, and shared memory holds only the unknown (rows + 2 halo rows).  A
// sweep then is 9 shared loads, the stencil arithmetic and one store per point: no
// coefficient traffic at all, not even from shared memory.  Same arithmetic, colour order,
// halo pushes and barriers as k_scb_sor_cluster: bit-identical results.
// =============================================================================
template <bool ALPHA, int MAXT = 576>
__global__ void __launch_bounds__(MAXT, 1) k_scb_sor_cluster_reg(ScbDev d, SorArgs a, int nloc_max) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ double su[];
  __shared__ double s_red[32];
  __shared__ double s_xmax[8];
  __shared__ int s_xfail[8];
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int nthe = d.nthe, npsi = d.npsi, nzeta = d.nzeta;
  const int tid = threadIdx.x, T = blockDim.x;
  const int sub = a.sub0 + blockIdx.x / CL;
  const int r0 = 1, r1 = ALPHA ? nzeta - 1 : npsi - a.nP - 1;
  const int c0 = a.nT, c1 = nthe - a.nT - 1;
  const int nr = r1 - r0 + 1;
  const int nloc_nom = (nr + CL - 1) / CL;
  const int rb = r0 + rank * nloc_nom;
  const int nloc = max(0, min(nloc_nom, r1 - rb + 1));
  const size_t sj = nthe, sk = (size_t)nthe * npsi;
  double* u = ALPHA ? d.alfa : d.psi;
  const size_t base = ALPHA ? sj * (size_t)(sub + 1) : sk * (size_t)(sub + 1);
  const size_t rstride = ALPHA ? sk : sj;
  for (int q = tid; q < (nloc + 2) * nthe; q += T) {
    const int lr = q / nthe, c = q - lr * nthe;
    su[q] = u[base + (size_t)(rb - 1 + lr) * rstride + c];
  }
  const double* rhs = ALPHA ? d.vecx : d.vecr;
  const double* cf[10] = {d.vecd, d.vec1, d.vec2, d.vec3, d.vec4, d.vec6, d.vec7, d.vec8, d.vec9, rhs};
  double* su_prev = (rank > 0) ? cluster.map_shared_rank(su, rank - 1) : nullptr;
  double* su_next = (rank < CL - 1) ? cluster.map_shared_rank(su, rank + 1) : nullptr;
  const int nloc_prev = nloc_nom;
  // my point of every colour: coefficients, shared-memory offset, global index, halo targets
  double kk[4][10];
  int off[4];       // shared-memory offset of the point (0: no point of this colour)
  int pushoff[4];   // > 0: offset in the next CTA's block, < 0: -(offset in the previous CTA's block), 0: none
#pragma unroll
  for (int col = 0; col < 4; ++col) {
    const int pc = col & 1, pr = col >> 1;
    const int cs = c0 + (((c0 & 1) == pc) ? 0 : 1);
    const int rs = rb + (((rb & 1) == pr) ? 0 : 1);
    const int rend = rb + nloc - 1;
    const int ncc = (cs <= c1) ? (c1 - cs) / 2 + 1 : 0;
    const int nrr = (rs <= rend) ? (rend - rs) / 2 + 1 : 0;
    const bool valid = tid < ncc * nrr;
    off[col] = 0; pushoff[col] = 0;
#pragma unroll
    for (int m = 0; m < 10; ++m) kk[col][m] = (m == 0) ? 1.0 : 0.0;
    if (valid) {
      const int rr = tid / ncc, cc = tid - rr * ncc;
      const int r = rs + 2 * rr, c = cs + 2 * cc;
      const int lr = r - rb + 1;
      off[col] = lr * nthe + c;                                 // >= nthe
      const size_t gq = base + (size_t)r * rstride + c;
#pragma unroll
      for (int m = 0; m < 10; ++m) kk[col][m] = cf[m][gq];
      if (lr == 1 && su_prev) pushoff[col] = -((nloc_prev + 1) * nthe + c);
      if (lr == nloc && su_next) pushoff[col] = c;              // (a one-row slab is excluded by the host)
    }
  }
  if (tid < 8) { s_xmax[tid] = 0.0; s_xfail[tid] = 0; }
  cluster.sync();
  double om = 1.0;
  int ni = 1;
  double lastmax = 0.0;
  bool failed = false, stopped = false;
  while (ni <= a.nimax) {
    double rmax = 0.0;
#pragma unroll
    for (int col = 0; col < 4; ++col) {
      if (off[col]) {
        double* uc = su + off[col];
        const double* um = uc - nthe;
        const double* up = uc + nthe;
        const double vd = kk[col][0];
        const double res = -vd * uc[0] + kk[col][1] * um[-1] + kk[col][2] * um[0] + kk[col][3] * um[1] + kk[col][4] * uc[-1] +
                           kk[col][5] * uc[1] + kk[col][6] * up[-1] + kk[col][7] * up[0] + kk[col][8] * up[1] - kk[col][9];
        double un = ALPHA ? (uc[0] + om * (res / vd)) : (uc[0] + om * res / vd);
        double rr2 = res;
        if (isnan(un) || un >= 1e10) {
          const int lr = off[col] / nthe, c = off[col] - lr * nthe;
          un = u[base + (size_t)(rb + lr - 1) * rstride + c];
          rr2 = 0.0;
          failed = true;
        }
        uc[0] = un;
        if (pushoff[col] > 0) su_next[pushoff[col]] = un;
        else if (pushoff[col] < 0) su_prev[-pushoff[col]] = un;
        rmax = fmax(rmax, fabs(rr2));      // every updated column lies in 1..nthe-2 (nT >= 2)
      }
      if (col == 0 || col == 2) {
        __syncthreads();
      } else if (col == 1) {
        cluster.sync();
      } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = rmax;
        const int anyfail = __syncthreads_or(failed ? 1 : 0);
        if (tid < CL) {
          double m = 0.0;
          for (int q = 0; q < (T + 31) / 32; ++q) m = fmax(m, s_red[q]);
          double* xm = cluster.map_shared_rank(s_xmax, tid);
          int* xf = cluster.map_shared_rank(s_xfail, tid);
          xm[rank] = m;
          xf[rank] = anyfail;
        }
        cluster.sync();
      }
    }
    double m = 0.0;
    int stop = 0;
    for (int q = 0; q < CL; ++q) { m = fmax(m, s_xmax[q]); stop |= s_xfail[q]; }
    lastmax = m;
    if (stop) { stopped = true; break; }
    om = a.omegaOpt;
    if (m < a.tol) break;
    ni = ni + 1;
  }
  for (int q = tid; q < nloc * nthe; q += T) {
    const int lr = 1 + q / nthe, c = q % nthe;
    if (c >= c0 && c <= c1) u[base + (size_t)(rb + lr - 1) * rstride + c] = su[lr * nthe + c];
  }
  if (rank == 0 && tid == 0) {
    a.ni[sub] = ni;
    a.resmax[sub] = lastmax;
    if (stopped) *a.fail = 1;
  }
  cluster.sync();
}


// =============================================================================
// k_scb_map_w<MODE>: the same re-gridding maps as k_scb_map (mapAlpha 0, mapPsi 1, mapTheta 2), one WARP per grid line
// with the line in shared memory.  k_scb_map walks a line with one thread through a global workspace -- ~100 dependent
// global round trips per line and only a few thousand threads on the device: it was 26 % of scb_run (ncu launch list,
// profiles/r2).  Here the CTA loads LPB lines as one coalesced tile, lane 0 of each warp does the two inherently serial
// prefixes (arc length, monotonicity filter) in shared memory, and the Steffen slopes and the evaluation at the target
// node values run one point per lane.  Per point the arithmetic is k_scb_map's, operation for operation: bit-identical.
// block = 32 * LPB threads; dynamic shared memory: LPB * 7 * nmax doubles.  grid: x = ceil(nlines / LPB)
// =============================================================================
template <int MODE, int LPB>
__global__ void __launch_bounds__(32 * LPB) k_scb_map_w(ScbDev d, const double* __restrict__ tgt, int nlines, int* __restrict__ fail) {
  extern __shared__ double mw_sm[];
  __shared__ int s_n1[LPB];
  __shared__ double s_tot[LPB];
  const int nthe = d.nthe, npsi = d.npsi, nzeta = d.nzeta;
  const size_t sj = nthe, sk = (size_t)nthe * npsi;
  const int n = (MODE == 0) ? nzeta + 1 : ((MODE == 1) ? npsi : nthe);
  const int n2 = (MODE == 0) ? nzeta - 1 : n, o0 = (MODE == 0) ? 1 : 0;
  const size_t st = (MODE == 0) ? sk : ((MODE == 1) ? sj : 1);
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, w = tid >> 5;
  const int line0 = blockIdx.x * LPB;
  const int nl = min(LPB, nlines - line0);
  auto line_base = [&](int line, int* kpl) -> size_t {
    if (MODE == 0) { *kpl = 0; return (size_t)line; }
    if (MODE == 1) { const int i = line % nthe; *kpl = 1 + line / nthe; return (size_t)i + sk * (*kpl); }
    const int j = line % npsi; *kpl = 1 + line / npsi; return sj * j + sk * (*kpl);
  };
  double* L = mw_sm;                                          // [line][7][n]: x-abscissa, f[3], yp[3]
  auto arr = [&](int lw, int a) -> double* { return L + ((size_t)lw * 7 + a) * n; };
  double* xyz[3] = {d.x, d.y, d.z};
  const double* absc = (MODE == 0) ? d.alfa : d.psi;
  // ---- load the tile: lines are adjacent along the fastest index for MODE 0 / 1, contiguous themselves for MODE 2
  for (int e = tid; e < nl * n; e += T) {
    int lw, q;
    if (MODE == 2) { lw = e / n; q = e - lw * n; } else { q = e / nl; lw = e - q * nl; }
    int kpl;
    const size_t o = line_base(line0 + lw, &kpl) + (size_t)q * st;
    arr(lw, 1)[q] = d.x[o]; arr(lw, 2)[q] = d.y[o]; arr(lw, 3)[q] = d.z[o];
    if (MODE != 2) arr(lw, 0)[q] = absc[o];
  }
  __syncthreads();
  const bool have = w < nl;
  // ---- serial prefixes, lane 0 of the line's warp
  if (have && lane == 0) {
    double* Wx = arr(w, 0);
    double *f0 = arr(w, 1), *f1 = arr(w, 2), *f2 = arr(w, 3);
    double total = 0.0;
    if (MODE == 2) {                                          // arc length along the line (src/ModScbEuler.f90:43-46)
      double dist = 0.0;
      Wx[0] = 0.0;
      for (int q = 1; q < n; ++q) {
        dist = dist + sqrt(sq(f0[q] - f0[q - 1]) + sq(f1[q] - f1[q - 1]) + sq(f2[q] - f2[q - 1]));
        Wx[q] = dist;
      }
      total = dist;
    }
    int n1 = 0;                                               // drop abscissae that do not increase (src/ModRamGSL.f90:262-273)
    double last = 0.0;
    for (int q = 0; q < n; ++q) {
      const double xv = (MODE == 2) ? Wx[q] / total * 3.141592653589793238462643383279502884197 : Wx[q];
      if (q == 0 || xv > last) {
        Wx[n1] = xv; f0[n1] = f0[q]; f1[n1] = f1[q]; f2[n1] = f2[q];
        last = xv;
        ++n1;
      }
    }
    s_n1[w] = n1;
    s_tot[w] = total;
    if (n1 < 3) atomicAdd(fail, 1);                           // GSLerr > 0 => SORFail
  }
  __syncthreads();
  const int n1 = have ? s_n1[w] : 0;
  const bool ok = have && n1 >= 3;
  // ---- Steffen node slopes (gsl interpolation/steffen.c, steffen_init), a node per lane
  if (ok) {
    const double* Wx = arr(w, 0);
    for (int i = lane; i < n1; i += 32) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double* fa = arr(w, 1 + c);
        double v;
        if (i == 0) v = (fa[1] - fa[0]) / (Wx[1] - Wx[0]);
        else if (i == n1 - 1) v = (fa[n1 - 1] - fa[n1 - 2]) / (Wx[n1 - 1] - Wx[n1 - 2]);
        else {
          const double hi = Wx[i + 1] - Wx[i];
          const double him1 = Wx[i] - Wx[i - 1];
          const double si = (fa[i + 1] - fa[i]) / hi;
          const double sim1 = (fa[i] - fa[i - 1]) / him1;
          const double pi = (sim1 * hi + si * him1) / (him1 + hi);
          const double m1 = fabs(si) < 0.5 * fabs(pi) ? fabs(si) : 0.5 * fabs(pi);
          const double m2 = fabs(sim1) < m1 ? fabs(sim1) : m1;
          v = (steffen_sgn(sim1) + steffen_sgn(si)) * m2;
        }
        arr(w, 4 + c)[i] = v;
      }
    }
  }
  __syncthreads();
  // ---- evaluate at the prescribed node values, a target per lane; results straight to global memory
  int kpl = 0;
  const size_t base = have ? line_base(line0 + w, &kpl) : 0;
  bool bad = false;
  if (ok) {
    const double* Wx = arr(w, 0);
    const double xa0 = Wx[0], xa1 = Wx[1], xaN = Wx[n1 - 1], xaM = Wx[n1 - 2];
    for (int q = lane; q < n2; q += 32) {
      const double xb = tgt[(MODE == 0) ? q + 1 : q];
      const size_t o = base + (size_t)(o0 + q) * st;
      if (xb <= xa0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { const double* Wf = arr(w, 1 + c); xyz[c][o] = Wf[0] + (xb - xa0) / (xa1 - xa0) * (Wf[1] - Wf[0]); }
      } else if (xb >= xaN) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { const double* Wf = arr(w, 1 + c); xyz[c][o] = Wf[n1 - 1] + (xb - xaN) / (xaM - xaN) * (Wf[n1 - 2] - Wf[n1 - 1]); }
      } else if (xb == xb) {
        int ilo = 0, ihi = n1 - 1;
        while (ihi > ilo + 1) {                              // gsl_interp_bsearch
          const int i = (ihi + ilo) / 2;
          if (Wx[i] > xb) ihi = i; else ilo = i;
        }
        const double xl = Wx[ilo];
        const double hi = Wx[ilo + 1] - xl;
        const double delx = xb - xl;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double* Wf = arr(w, 1 + c);
          const double* Wp = arr(w, 4 + c);
          const double f0 = Wf[ilo];
          const double si = (Wf[ilo + 1] - f0) / hi;
          const double y0 = Wp[ilo], y1 = Wp[ilo + 1];
          const double a = (y0 + y1 - 2 * si) / hi / hi;
          const double b = (3 * si - 2 * y0 - y1) / hi;
          xyz[c][o] = f0 + delx * (y0 + delx * (b + delx * a));
        }
      } else {
        bad = true;
      }
    }
  }
  {
    int b = bad ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b |= __shfl_xor_sync(0xffffffffu, b, o);
    bad = b != 0;
  }
  if (bad && lane == 0) atomicAdd(fail, 1);
  __syncwarp();                                               // orders this warp's global writes before the copies below
  // ---- periodic planes (zeta = 1 <- nzeta, nzeta+1 <- 2) and the reset of the potential; a line that failed keeps its points
  if (ok && !bad) {
    if (MODE == 0) {
      if (lane < 3) {
        double* a = xyz[lane];
        a[base] = a[base + (size_t)(nzeta - 1) * st];
        a[base + (size_t)nzeta * st] = a[base + st];
      }
      for (int k = lane; k <= nzeta; k += 32) d.alfa[base + (size_t)k * st] = tgt[k];          // alfges
    } else {
      const long long shift = (kpl == nzeta - 1) ? -(long long)sk * (nzeta - 1) : ((kpl == 1) ? (long long)sk * (nzeta - 1) : 0);
      for (int q = lane; q < n; q += 32) {
        const size_t o = base + (size_t)q * st;
        if (shift != 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) xyz[c][(size_t)((long long)o + shift)] = xyz[c][o];
        }
        if (MODE == 1) {                                                                        // psiges
          d.psi[o] = tgt[q];
          if (shift != 0) d.psi[(size_t)((long long)o + shift)] = tgt[q];
        }
      }
    }
  }
}
