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
