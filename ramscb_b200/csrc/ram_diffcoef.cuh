// ANISCH, second half (src/ModRamRun.f90:422-605): the pitch-angle diffusion coefficients of WPADIF rebuilt on the
// device, straight into the [l][k][Pp] arrays the WPADIF kernels read.  Runs once per Dt_bc = 300 s of
// simulated time: written for clarity, not for speed.
//   chorus (electrons outside the plasmapause, XNE <= 50): GSL_Interpolation_1D = Steffen spline of log10 <Daa> over the
//     table's pitch angles PA, evaluated at PAbn (src/ModRamGSL.f90:240-311, src/RamGSL.c:111-174)           -> ATAC
//   hiss (electrons inside): GSL_Interpolation_2D = gsl_interp2d_bilinear in (log10 E, fpe/fce) (src/RamGSL.c:178-214) -> ATAW
//   EMIC (H+): the same bilinear rule on the H-band / He-band tables, scaled by I_emic (src/ModRamWPI.f90:720-750)
//                                                                                         -> ATAW_emic_h, ATAW_emic_he
// Operation order of the reference throughout (compiled with -fmad=false): bit-identical to the oracle up to the
// device's log10 / pow.
#pragma once
#include "ram_common.cuh"

struct DiffTabs {
  int ENG, NCF, ENGe, NCFe, use_bas, n1;   // n1: abscissae left after the wrapper's monotonicity filter
  const double *ALENOR, *fpofc, *NDAAJ;    // log10(ENOR) [ENG], [NCF], (NR,ENG,NPA,NCF)
  const double* DAAR;                      // CDAAR or BDAAR (NR,NT,NE,NPA)
  const double *logEe, *fp2c, *DH, *DHE;   // log10(EKEV_emic) [ENGe], [NCFe], (NR,ENGe,NPA,NCFe) x 2
  const double *Ihs, *Ihes;                // (4,NR,NT)
  const double *PAx, *PAbn;                // filtered ascending PA [n1], PAbn [NPA]
  const int* PAidx;                        // index (0-based L) of each kept abscissa [n1]
  const double* XNE;                       // (NR,NT)
  const double* GRELs;                     // GREL(S,:) [NE]
  double RMASs, RMASe, Bw;
  int cls;                                 // I_emic class 1..4 (0: intensities 0)
};

__device__ __forceinline__ int dc_bsearch(const double* xa, double x, int n) {
  int ilo = 0, ihi = n - 1;
  while (ihi > ilo + 1) {
    const int i = (ihi + ilo) / 2;
    if (xa[i] > x) ihi = i; else ilo = i;
  }
  return ilo;
}
// bilinear rule of gsl interp2d/bilinear.c on za[j*n1 + i] = f(x_i, y_j); the table values are fetched through `zf`
template <class ZF>
__device__ __forceinline__ double dc_bilinear(int n1, int m1, const double* xa, const double* ya, ZF zf, double x, double y) {
  const int xi = dc_bsearch(xa, x, n1), yi = dc_bsearch(ya, y, m1);
  const double xmin = xa[xi], xmax = xa[xi + 1], ymin = ya[yi], ymax = ya[yi + 1];
  const double zminmin = zf(xi, yi), zminmax = zf(xi, yi + 1), zmaxmin = zf(xi + 1, yi), zmaxmax = zf(xi + 1, yi + 1);
  const double dx = xmax - xmin, dy = ymax - ymin;
  const double t = (x - xmin) / dx, u = (y - ymin) / dy;
  return (1. - t) * (1. - u) * zminmin + t * (1. - u) * zmaxmin + (1. - t) * u * zminmax + t * u * zmaxmax;
}

// chorus: one warp per (I, J, K) line; shared: per warp 3 * NPA doubles (values, slopes, abscissa copy not needed)
// grid: x = ceil(lines / warps per CTA); block = 128
__global__ void __launch_bounds__(128) k_diffcoef_chorus(const __grid_constant__ RamDev d, const __grid_constant__ DiffTabs t,
                                                         double* __restrict__ ATAC, int* __restrict__ nerr) {
  extern __shared__ double dc_sm[];
  const int NR = d.NR, NT = d.NT, NE = d.NE, NPA = d.NPA, Pp = d.Pp;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* fa = dc_sm + (size_t)w * 2 * NPA;
  double* yp = fa + NPA;
  const long long line = (long long)blockIdx.x * nw + w;
  const long long nlines = (long long)(NR - 1) * NT * (NE - 1);
  if (line >= nlines) return;
  const int i = 1 + (int)(line % (NR - 1));
  const int j = (int)((line / (NR - 1)) % NT);
  const int k = 1 + (int)(line / ((long long)(NR - 1) * NT));
  if (!(t.XNE[(size_t)j * NR + i] <= 50.)) return;
  const int n1 = t.n1;
  const double* xa = t.PAx;
  for (int q = lane; q < n1; q += 32) {
    const int L0 = t.PAidx[q];                                   // 0-based index into DAMR1
    const int Lsrc = t.use_bas ? (NPA - 1 - L0) : L0;            // CDAAR(I,J,K,nPa-L+1) | BDAAR(I,J,K,L)
    fa[q] = log10(t.DAAR[i + (size_t)NR * (j + (size_t)NT * (k + (size_t)NE * Lsrc))]);
  }
  __syncwarp();
  if (n1 < 3) { if (lane == 0) atomicAdd(nerr, 1); return; }
  for (int q = lane; q < n1; q += 32) {                          // steffen_init
    double v;
    if (q == 0) v = (fa[1] - fa[0]) / (xa[1] - xa[0]);
    else if (q == n1 - 1) v = (fa[n1 - 1] - fa[n1 - 2]) / (xa[n1 - 1] - xa[n1 - 2]);
    else {
      const double hi = xa[q + 1] - xa[q], him1 = xa[q] - xa[q - 1];
      const double si = (fa[q + 1] - fa[q]) / hi, sim1 = (fa[q] - fa[q - 1]) / him1;
      const double pi = (sim1 * hi + si * him1) / (him1 + hi);
      const double m1 = fabs(si) < 0.5 * fabs(pi) ? fabs(si) : 0.5 * fabs(pi);
      const double m2 = fabs(sim1) < m1 ? fabs(sim1) : m1;
      v = (((sim1 < 0) ? -1.0 : 1.0) + ((si < 0) ? -1.0 : 1.0)) * m2;
    }
    yp[q] = v;
  }
  __syncwarp();
  bool bad = false;
  for (int L0 = lane; L0 < NPA; L0 += 32) {
    const double xb = t.PAbn[L0];
    double Y;
    if (xb <= xa[0]) Y = fa[0] + (xb - xa[0]) / (xa[1] - xa[0]) * (fa[1] - fa[0]);
    else if (xb >= xa[n1 - 1]) Y = fa[n1 - 1] + (xb - xa[n1 - 1]) / (xa[n1 - 2] - xa[n1 - 1]) * (fa[n1 - 2] - fa[n1 - 1]);
    else if (xb == xb) {
      const int ilo = dc_bsearch(xa, xb, n1);
      const double hi = xa[ilo + 1] - xa[ilo], delx = xb - xa[ilo];
      const double si = (fa[ilo + 1] - fa[ilo]) / hi;
      const double a = (yp[ilo] + yp[ilo + 1] - 2 * si) / hi / hi;
      const double b = (3 * si - 2 * yp[ilo] - yp[ilo + 1]) / hi;
      Y = fa[ilo] + delx * (yp[ilo] + delx * (b + delx * a));
    } else { bad = true; Y = xb; }
    const double MUBOUN = d.MU[L0] + 0.5 * d.WMU[L0];
    double taudaa = pow(10., Y) * 1.0;
    if (taudaa > 1e0) taudaa = 1e-1;
    if (taudaa < 1e-30) taudaa = 1e-30;
    ATAC[((size_t)L0 * NE + k) * Pp + (size_t)j * NR + i] = taudaa * (1. - MUBOUN * MUBOUN) * MUBOUN * d.BOUNHS[((size_t)L0 * NT + j) * d.NR1 + i];
  }
  if (bad && lane == 0) atomicAdd(nerr, 1);
}

// hiss (which = 0) / EMIC (which = 1): one thread per (I, J, K, L)
__global__ void __launch_bounds__(128) k_diffcoef_bilinear(const __grid_constant__ RamDev d, const __grid_constant__ DiffTabs t, int which,
                                                           double* __restrict__ A0, double* __restrict__ A1) {
  const int NR = d.NR, NT = d.NT, NE = d.NE, NPA = d.NPA, Pp = d.Pp;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)(NR - 1) * NT * (NE - 1) * NPA;
  if (q >= n) return;
  const int i = 1 + (int)(q % (NR - 1));
  const int j = (int)((q / (NR - 1)) % NT);
  const int k = 1 + (int)((q / ((long long)(NR - 1) * NT)) % (NE - 1));
  const int L0 = (int)(q / ((long long)(NR - 1) * NT * (NE - 1)));
  const double CS = 2.998E8, PI = 3.1415926535897932384626433832795, Q = 1.602E-19;
  const double xne = t.XNE[(size_t)j * NR + i];
  const double bnes = d.BNES[(size_t)j * d.NR1 + i];
  const double MUBOUN = d.MU[L0] + 0.5 * d.WMU[L0];
  const double ER1 = log10(d.EKEV[k]);
  const size_t o = ((size_t)L0 * NE + k) * Pp + (size_t)j * NR + i;
  if (which == 0) {
    if (!(xne > 50.)) return;
    const double cv = CS * 100, esu = Q * 3E9, gausgam = 1.E-5;
    const double omega = esu * 10 * bnes / (t.RMASs * cv);
    double xfrl = CS * sqrt(xne * t.RMASs * 40 * PI) / 10. / bnes;
    if (xfrl > 18) xfrl = 18.;
    if (xfrl < 2) xfrl = 2.;
    const double fnorm = omega * ((t.Bw * 1e-3) * (t.Bw * 1e-3)) * (gausgam * gausgam) / 1e8 / bnes / bnes;
    const int ENG = t.ENG;
    const double* tab = t.NDAAJ;
    auto zf = [&](int kn, int iz) { return log10(tab[i + (size_t)NR * (kn + (size_t)ENG * (L0 + (size_t)NPA * iz))]); };
    const double Y = dc_bilinear(ENG, t.NCF, t.ALENOR, t.fpofc, zf, ER1, xfrl);
    A0[o] = pow(10., Y) * fnorm / (t.GRELs[k] * t.GRELs[k]) * (1. - MUBOUN * MUBOUN) / MUBOUN;
  } else {
    double xfrl = CS * sqrt(xne * t.RMASe * 40 * PI) / 10. / bnes;
    if (xfrl > 20) xfrl = 20.;
    if (xfrl < 2) xfrl = 2.;
    const double fh = t.cls ? t.Ihs[(t.cls - 1) + 4 * (i + (size_t)NR * j)] : 0.0;
    const double fhe = t.cls ? t.Ihes[(t.cls - 1) + 4 * (i + (size_t)NR * j)] : 0.0;
    const int ENG = t.ENGe;
    const double bh = d.BOUNHS[((size_t)L0 * NT + j) * d.NR1 + i];
    const double *t1 = t.DH, *t2 = t.DHE;
    auto z1 = [&](int kn, int iz) { return log10(t1[i + (size_t)NR * (kn + (size_t)ENG * (L0 + (size_t)NPA * iz))]); };
    auto z2 = [&](int kn, int iz) { return log10(t2[i + (size_t)NR * (kn + (size_t)ENG * (L0 + (size_t)NPA * iz))]); };
    double Y = dc_bilinear(ENG, t.NCFe, t.logEe, t.fp2c, z1, ER1, xfrl);
    double vh = pow(10., Y) * fh * (1. - MUBOUN * MUBOUN) * MUBOUN * bh;
    Y = dc_bilinear(ENG, t.NCFe, t.logEe, t.fp2c, z2, ER1, xfrl);
    double vhe = pow(10., Y) * fhe * (1. - MUBOUN * MUBOUN) * MUBOUN * bh;
    if (vh <= 1.0e-20) vh = 1.0e-31;
    if (vhe <= 1.0e-20) vhe = 1.0e-31;
    A0[o] = vh;
    A1[o] = vhe;
  }
}
