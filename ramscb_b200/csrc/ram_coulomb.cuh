// Coulomb operators of ModRamCoul.f90 on the device (reference operation order, bit-identical
// to the oracle; optional operators, so no FAST variant).
//   COULPARA (:17-125)  -> host tables (ram_gpu.cu: tables_coulomb), COULE/COULI/ATA/GTA [k][l]
//   COULEN   (:133-221) -> k_coulen: energy drag, DRIFTE-like limited flux along K
//   COULMU   (:229-296) -> k_coulmu: pitch-angle scattering, Thomas recurrences along L
#pragma once
#include "ram_kernels.cuh"

// COULEN: one thread per (L, plane position) line, walking K with a rolling window; in place.
// grid: x = tiles of p, y = L-2 (L = 2..NPA); ghosts F(1), F(0) with COULEN's own ratios
// (:182-183: the square roots are the reciprocals of DRIFTE's).
__global__ void __launch_bounds__(128) k_coulen(const __grid_constant__ RamDev d, SpecDev sp, const double* __restrict__ COULE,
                                                const double* __restrict__ COULI, const double* __restrict__ NECR, double GREL1,
                                                double GREL2, double GRZERO, double g1, double g0) {
  const int NE = d.NE, NPA = d.NPA, Pp = d.Pp;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = 2 + blockIdx.y;
  if (p >= d.P) return;
  const int j = p / d.NR, i = p - j * d.NR;
  if (i < 1) return;
  const int I = i + 1, J = j + 1;
  // BANE(L) (:170-173): L >= NPA-10 repeat the value of L = NPA-11
  const int Lb = min(L, NPA - 11);
  const double BANE = (1. - R3(d.FNIS, I, J, Lb) / 2. / R3(d.FNHS, I, J, Lb)) / (1. - d.MU[Lb - 1] * d.MU[Lb - 1]);
  const double XNE = NECR[p] * BANE;
  const double beta = d.BetaLim;
  double* F = sp.F + (size_t)(L - 1) * NE * Pp + p;     // F(K) at F[(K-1)*Pp]
  const double f2 = F[Pp];
  const double Fg1 = f2 * GREL1 / GREL2 * g1;
  const double Fg0 = Fg1 * GRZERO / GREL1 * g0;
#define GETK(K) (((K) > NE) ? 0.0 : (((K) >= 2) ? F[(size_t)((K)-1) * Pp] : (((K) == 1) ? Fg1 : Fg0)))
  double Fm1 = GETK(0), F0 = GETK(1), Fp1 = GETK(2), Fp2 = GETK(3);
  double cprev, FBprev;
  {
    const double c = (COULE[(size_t)0 * NPA + (L - 1)] + COULI[(size_t)0 * NPA + (L - 1)]) * XNE;
    FBprev = limited_flux(Fm1, F0, Fp1, Fp2, c, c / d.DE[0], beta);
    cprev = c;
  }
  for (int K = 2; K <= NE; ++K) {
    const double nn = GETK(K + 2);
    Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = nn;
    const double c = (COULE[(size_t)(K - 1) * NPA + (L - 1)] + COULI[(size_t)(K - 1) * NPA + (L - 1)]) * XNE;
    const double FB = limited_flux(Fm1, F0, Fp1, Fp2, c, c / d.DE[K - 1], beta);
    const double WEK = d.WE[K - 1];
    double fn = F0 - c / WEK * FB + cprev / WEK * FBprev;   // :212-213
    if (fn < 0.0) fn = 1E-15;
    F[(size_t)(K - 1) * Pp] = fn;
    cprev = c; FBprev = FB;
  }
#undef GETK
}

// COULMU: one thread per (K, plane position) line; RK/RL in shared memory [NPA][T].
// grid: x = tiles of (k, p); in place.  T_elapsed > 0 enables the negative clamp (:289).
__global__ void k_coulmu(const __grid_constant__ RamDev d, SpecDev sp, const double* __restrict__ ATA, const double* __restrict__ GTA,
                         const double* __restrict__ NECR, double T_elapsed) {
  extern __shared__ double smem[];
  const int NE = d.NE, NPA = d.NPA, Pp = d.Pp, T = blockDim.x;
  double* RK = smem;             // [NPA][T]
  double* RL = smem + NPA * T;
  const long long t = (long long)blockIdx.x * T + threadIdx.x;
  const int tx = threadIdx.x;
  if (t >= (long long)NE * Pp) return;
  const int k = (int)(t / Pp), p = (int)(t - (long long)k * Pp);
  if (p >= d.P || k < 1) return;
  const int j = p / d.NR, i = p - j * d.NR;
  if (i < 1) return;
  const int I = i + 1, J = j + 1;
  const size_t LS = (size_t)NE * Pp;
  double* F = sp.F + (size_t)k * Pp + p;
  const double XNE = NECR[p];
  RK[tx] = 0.;
  RL[tx] = -1.;
  double rkm = 0., rlm = -1.;
  double BASm = XNE * R3(d.BOUNIS, I, J, 1) / 2. / R3(d.BOUNHS, I, J, 1);   // BASCNE(L-1)
  for (int L = 2; L <= NPA - 1; ++L) {
    const double BOUNHSl = R3(d.BOUNHS, I, J, L), FNHSl = R3(d.FNHS, I, J, L);
    const double BAS = XNE * R3(d.BOUNIS, I, J, L) / 2. / BOUNHSl;
    const double AN = ATA[(size_t)k * NPA + (L - 1)] * BAS / FNHSl * BOUNHSl;
    const double GN = GTA[(size_t)k * NPA + (L - 1)] * BASm / FNHSl * R3(d.BOUNHS, I, J, L - 1);
    const double BN = AN + GN;
    const double RP = F[(size_t)(L - 1) * LS] / FNHSl / d.MU[L - 1];
    const double DENOM = BN + GN * rlm + 1;
    rkm = (RP + GN * rkm) / DENOM;
    rlm = -AN / DENOM;
    RK[(L - 1) * T + tx] = rkm;
    RL[(L - 1) * T + tx] = rlm;
    BASm = BAS;
  }
  auto put = [&](int L, double f) {
    double v = f * R3(d.FNHS, I, J, L) * d.MU[L - 1];
    if ((T_elapsed > 0.0) && (v < 0.0)) v = 1E-15;
    F[(size_t)(L - 1) * LS] = v;
  };
  double f = rkm / (1 + rlm);        // F2(NPA-1)
  put(NPA, f);                       // F2(NPA) = F2(NPA-1)  (:284)
  put(NPA - 1, f);
  for (int L = NPA - 2; L >= 1; --L) {
    f = RK[(L - 1) * T + tx] - RL[(L - 1) * T + tx] * f;
    put(L, f);
  }
}


// =============================================================================
// PARA_FLC (src/ModRamLoss.f90:342-455): field-line-curvature pitch-angle diffusion coefficient
// FLC_coef of one species, built on the device from the equatorial curvature radius and zeta
// parameters (NR,NT; the 2-D output of FLC_Radius) -- the species' (NR,NT,NE,NPA) array no longer
// crosses the bus.  One thread per (plane position, energy); the two pitch-angle loops of the
// reference run in the thread (D(l) is recomputed in the second one with the same operations).
// in: tab = [r_curvEq(P) | zeta1Eq(P) | zeta2Eq(P) | V(S,1:NE) | LZ(1:NR)], out: flc [l][k][Pp]
// (zero-filled by the caller: L = NPA and the lines with epsilon < 0.1 stay 0).
// =============================================================================
__global__ void __launch_bounds__(128) k_para_flc(const __grid_constant__ RamDev d, const double* __restrict__ tab, double rmas,
                                                  double* __restrict__ flc) {
  const int P = d.P, NE = d.NE, NPA = d.NPA, NR = d.NR;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P * NE) return;
  const int k = q / P, p = q - k * P;
  const int j = p / NR, i = p - j * NR;
  const double Q = 1.602E-19, REarth = 6.4 * 1.E6;
  const double Vk = tab[3 * P + k], LZi = tab[3 * P + NE + i];
  const size_t f2 = (size_t)i + (size_t)d.NR1 * j;                    // (NR+1,NT) raw field index
  const double r_gyro = rmas * Vk / fabs(d.BNES[f2] * Q);
  double epsl = r_gyro / tab[p];
  if (epsl > 0.584) epsl = 0.584;
  if (!(epsl >= 0.1)) return;
  const double e1 = 1.0 / epsl, e2 = 1.0 / (epsl * epsl), e3 = 1.0 / (epsl * epsl * epsl);
  const double a1 = -0.35533865 + 0.12800347 * e1 + 0.0017113113 * e2;
  const double a2 = 0.23156321 + 0.15561211 * e1 - 0.001860433 * e2;
  const double ba = -0.51057275 + 0.93651781 * e1 - 0.031690658 * e2;
  const double ca = 1.0663037 - 1.0944973 * e1 + 0.016679378 * e2 - 0.000499 * e3;
  const double da = -0.49667826 - 0.0081941799 * e1 + 0.0013621659 * e2;
  const double omegaa = 1.0513540 + 0.1351358 * epsl - 0.50787555 * (epsl * epsl);
  const double Am = exp(ca) * (pow(tab[P + p], a1) * pow(tab[2 * P + p], a2) + da);
  const size_t n2 = (size_t)d.NR1 * d.NT;
  double Nmin = 1.0e20, nf1 = 0.0;
  int lmin = 0;
  for (int l = 0; l < NPA - 1; ++l) {
    const double MUBOUN = d.MU[l] + 0.5 * d.WMU[l];
    const double alph = acos(MUBOUN);
    const double nf = 1.0 / (sin(omegaa * alph) * pow(MUBOUN, ba));
    if (l == 0) nf1 = nf;
    if (nf <= Nmin) {
      Nmin = nf;
      lmin = l;
    }
  }
  const double nfm = lmin == 0 ? nf1 : Nmin;                           // Nfactor(lmin), also when no comparison held (NaN)
  for (int l = 0; l < NPA - 1; ++l) {
    const double MUBOUN = d.MU[l] + 0.5 * d.WMU[l];
    const double alph = acos(MUBOUN);
    const double bh = d.BOUNHS[f2 + n2 * l];
    const double tau_bounce = 4 * LZi * REarth * bh / Vk;
    const double D = (Am * Am) / (2 * tau_bounce);
    const double sn = sin(omegaa * alph);
    const double Daa = D * (nfm * nfm) * (sn * sn) * pow(MUBOUN, 2 * ba) / ((1 - MUBOUN * MUBOUN) * (MUBOUN * MUBOUN));
    flc[((size_t)l * NE + k) * d.Pp + p] = Daa * (1 - MUBOUN * MUBOUN) * MUBOUN * bh;
  }
}
