// RAM hot-path kernels for sm_100a (FP64, HBM-bound design; no tensor cores:
// nothing here is a dense contraction).  Compiled with -fmad=false: in EXACT
// mode every expression is evaluated in the reference's operation order with
// separately rounded products, so results are bit-identical to the CPU oracle
// (IEEE-754 +,-,*,/ and sqrt are correctly rounded on the device).  FAST-mode
// code asks for fused multiply-adds explicitly with fma().
//
// Reference routines restated here: src/ModRamDrift.f90 (DRIFTPARA/R/P/E/MU),
// src/ModRamLoss.f90 (CEPARA/CHAREXCHANGE/ATMOL), src/ModRamWPI.f90
// (WAVELO/WPADIF), src/ModRamRun.f90 (SUMRC/ANISCH).
#pragma once
#include "ram_common.cuh"

#define OME_EARTH 7.3E-5

// ---- raw-field accessors (1-based Fortran indices) ---------------------------
#define R2(a, I, J) (a)[(size_t)((J)-1) * d.NR1 + ((I)-1)]
#define R3(a, I, J, L) (a)[((size_t)((L)-1) * d.NT + ((J)-1)) * d.NR1 + ((I)-1)]

__device__ __forceinline__ unsigned long long dbl_bits(double x) { return (unsigned long long)__double_as_longlong(x); }

// block-wide min of a positive double, then one atomicMin on the ordered bit
// pattern (min is order independent => deterministic)
__device__ __forceinline__ void block_min_to(unsigned long long* dst, double v) {
  unsigned long long b = dbl_bits(v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
    b = t < b ? t : b;
  }
  if ((threadIdx.x & 31) == 0) atomicMin(dst, b);
}

// compare-select min/max with std::min/std::max semantics (what the oracle and
// Fortran MIN/MAX compute for ordered operands); IEEE fmin/fmax cost 6-7
// instructions each on sm_100 because of their NaN rules.
__device__ __forceinline__ double dmin(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }

// =============================================================================
// flux limiter, SURVEY appendix A.1 (ModRamDrift.f90:170-182, 246-259, 349-361,
// 441-453).  Window: Fm1=F(m-1) F0=F(m) Fp1=F(m+1) Fp2=F(m+2); c = Courant
// number at interface m; chat = c, c/DE(K) or c/DMU(L).
// =============================================================================
__device__ __forceinline__ double limited_flux(double Fm1, double F0, double Fp1, double Fp2, double c, double chat,
                                               double beta) {
  const double sgn = (c < 0.0) ? -1.0 : 1.0;
  const double X = Fp1 - F0;
  const double FUP = 0.5 * ((F0 + Fp1) - sgn * X);
  double FB = FUP;
  if (fabs(X) > 1.E-27) {
    const double num = (c < 0.0) ? (Fp2 - Fp1) : (F0 - Fm1);
    const double R = num / X;
    if (R > 0.0) {
      const double LIM = dmax(dmin(beta * R, 1.0), dmin(R, beta));
      const double CORR = (-0.5 * (chat - sgn)) * X;
      FB = FUP + LIM * CORR;
    }
  }
  return FB;
}

// DRIFTP's interface J = NT-1 with a negative coefficient: the reference's far-upwind index N = J+2 wraps to N-NT+1 = 2, so
// its difference is F(2) - F(1) with the STORED F(1) (src/ModRamDrift.f90:251-254), not F(NT+1) - F(NT).  The two agree
// while every (I,J) input is periodic in MLT; computehI's smoothed field arrays are not (the 9 x 9 Gaussian reflects at
// the MLT edges after the continuity fix of src/ModRamScb.f90:474-477), and then F2(J=1) and F2(J=NT) part between two
// DRIFTP sweeps.  numneg = that difference.
__device__ __forceinline__ double limited_flux_num(double Fm1, double F0, double Fp1, double numneg, double c, double chat,
                                                   double beta) {
  const double sgn = (c < 0.0) ? -1.0 : 1.0;
  const double X = Fp1 - F0;
  const double FUP = 0.5 * ((F0 + Fp1) - sgn * X);
  double FB = FUP;
  if (fabs(X) > 1.E-27) {
    const double num = (c < 0.0) ? numneg : (F0 - Fm1);
    const double R = num / X;
    if (R > 0.0) {
      const double LIM = dmax(dmin(beta * R, 1.0), dmin(R, beta));
      const double CORR = (-0.5 * (chat - sgn)) * X;
      FB = FUP + LIM * CORR;
    }
  }
  return FB;
}

// FAST mode: the same limiter in upwind form, without the division.  With
// U/D/UU the upwind, downwind and far-upwind cells, dU=U-UU, dD=D-U, R=dU/dD:
//   FBND = U + 0.5*(1-|chat|) * LIM(R)*dD,  LIM(R)*|dD| = min(max(|dU|,|dD|), beta*min(|dU|,|dD|))
// (max(min(bR,1),min(R,b)) = min(max(R,1), b*min(R,1)) for b >= 1).  achat = |chat|.
// Mathematically identical to the reference; rounding differs at the 1e-16 level.
// Written on the cell differences so a walking sweep computes each difference once:
//   dm1 = F0-Fm1, d0 = Fp1-F0, dp1 = Fp2-Fp1;  dU = neg ? -dp1 : dm1,  dD = neg ? -d0 : d0.
// |.| and the sign tests are done on the high words (dU*dD > 0 <=> same sign, and a zero dU
// gives lim = 0 anyway; the product's underflow to 0 below 1e-308 is the only difference).
__device__ __forceinline__ double limited_flux_d(double F0, double Fp1, double dm1, double d0, double dp1, bool neg, double achat,
                                                 double beta) {
  // the upwind value as the reference forms it, FUP = 0.5*((F0+Fp1) - sgn*X) (:172 etc.): algebraically U, but its
  // rounding noise (ulp of the LARGER of the two cells) is part of the reference's result next to steep gradients
  const double U = 0.5 * ((F0 + Fp1) + (neg ? d0 : -d0));
  const double ds = neg ? dp1 : dm1;
  const int hs = __double2hiint(ds), h0 = __double2hiint(d0);
  const double an = __hiloint2double(hs & 0x7fffffff, __double2loint(ds));
  const double ad = __hiloint2double(h0 & 0x7fffffff, __double2loint(d0));
  const bool lt = an < ad;
  const double m = lt ? an : ad, M = lt ? ad : an;
  const double bm = beta * m;
  const double lim = (bm < M) ? bm : M;
  const bool use = (ad > 1.E-27) && ((hs ^ h0) >= 0);
  const int sgn = (h0 ^ (neg ? (int)0x80000000 : 0)) & (int)0x80000000;   // sign of dD
  const double limx = use ? __hiloint2double(__double2hiint(lim) | sgn, __double2loint(lim)) : 0.0;
  return fma(fma(-0.5, achat, 0.5), limx, U);
}
__device__ __forceinline__ double limited_flux_fast(double Fm1, double F0, double Fp1, double Fp2, bool neg, double achat,
                                                    double beta) {
  return limited_flux_d(F0, Fp1, F0 - Fm1, Fp1 - F0, Fp2 - Fp1, neg, achat, beta);
}

// FAST mode exp(x) for x <= 0 (loss factors): x = (64 n + j) ln2/64 + r, |r| <= ln2/128,
// exp(x) = 2^n * 2^(j/64) * (1 + r + r^2/2 + ... + r^5/120)   (next term 3.5e-17), ~1 ulp.
// tab = 2^(j/64).  Arguments below -708 (result < 3e-308) are clamped: no denormal scaling.
__device__ __forceinline__ double fast_exp(double x, const double* __restrict__ tab) {
  x = (x < -708.0) ? -708.0 : x;
  const double MAGIC = 6755399441055744.0;                  // 1.5 * 2^52: rint in the low word
  const double t = fma(x, 92.33248261689366, MAGIC);     // 64/ln2
  const int ki = __double2loint(t);
  const double kf = t - MAGIC;
  double r = fma(kf, -0.01083042469326756, x);               // ln2/64 high part (21 trailing zero bits)
  r = fma(kf, -2.9815858269852933e-12, r);                   // low part
  const double T = tab[ki & 63];
  double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
  q = fma(q, r, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  q = fma(q * r, r, r);                                      // r + r^2/2 + ... + r^5/120
  const double v = fma(T, q, T);
  return __hiloint2double(__double2hiint(v) + ((ki >> 6) << 20), __double2loint(v));
}

// =============================================================================
// prep kernels: coefficient pieces that do not depend on energy, evaluated in
// the reference's order so the sweeps only do the K-dependent tail.
// =============================================================================

// DTs-independent pieces; run when the field arrays change (after computehI).
__global__ void k_prep_fields(RamDev d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.NPA * d.Pp) return;
  const int l = t / d.Pp, p = t - l * d.Pp;
  const size_t o = (size_t)l * d.Pp + p;
  if (p >= d.P) {
    d.t1[o] = 0; d.G[o] = 0; d.sFp[o] = 1; d.Gr[o] = 0; d.Gp[o] = 0; d.DRD2[o] = 0; d.DPD2[o] = 0;
    d.dBdt1[o] = 0; d.dIdt1[o] = 0; d.FNHSc[o] = 1; d.Gmr[o] = 0; d.Gmp[o] = 0; d.DRM2[o] = 0; d.DPM2[o] = 0;
    d.dIbndt2[o] = 0; d.BOUNHSc[o] = 1; d.HDNSc[o] = 0;
    d.fRb[o] = 0; d.fPb[o] = 0; d.fEb[o] = 0; d.rFNHS[o] = 1;
    if (l == 0) { d.sB[p] = 1; d.sBp[p] = 1; d.BNESc[p] = 1; d.RLZp[p] = 1; d.outp[p] = 1; }
    return;
  }
  const int j = p / d.NR, i = p - j * d.NR;
  const int I = i + 1, J = j + 1, L = l + 1;
  const int NT = d.NT, NR = d.NR;
  int J0 = J - 1; if (J == 1) J0 = NT - 1;
  int J1 = J + 1; if (J == NT) J1 = 2;
  const double MDR = d.MDR, DPHI = d.DPHI;
  const double RLZI = d.RLZ[i];
  const double *BNES = d.BNES, *FNIS = d.FNIS, *FNHS = d.FNHS, *BOUNIS = d.BOUNIS, *BOUNHS = d.BOUNHS;

  if (l == 0) {
    d.sB[p] = R2(BNES, I, J) + R2(BNES, I + 1, J);
    d.BNESc[p] = R2(BNES, I, J);
    d.RLZp[p] = RLZI;
    d.outp[p] = (unsigned char)(d.outside[(size_t)j * NR + i] != 0);
    d.sBp[p] = (I >= 2 && J >= 2) ? (R2(BNES, I, J) + R2(BNES, I, J1)) : 1.0;
  }
  d.FNHSc[o] = R3(FNHS, I, J, L);
  d.rFNHS[o] = 1.0 / R3(FNHS, I, J, L);
  d.BOUNHSc[o] = R3(BOUNHS, I, J, L);
  d.HDNSc[o] = R3(d.HDNS, I, J, L);

  {  // DRIFTR :140-144 (all I, all J)
    const double CGR1 = R3(FNIS, I + 1, J1, L) + R3(FNIS, I, J1, L) - R3(FNIS, I + 1, J0, L) - R3(FNIS, I, J0, L);
    const double CGR2 = R2(BNES, I + 1, J1) + R2(BNES, I, J1) - R2(BNES, I + 1, J0) - R2(BNES, I, J0);
    const double CGR3 =
        CGR1 + (R3(FNIS, I + 1, J, L) + R3(FNIS, I, J, L) - 2 * R3(FNHS, I + 1, J, L) - 2 * R3(FNHS, I, J, L)) * CGR2 / 2. /
                   (R2(BNES, I + 1, J) + R2(BNES, I, J));
    d.t1[o] = CGR3 / (R3(FNHS, I, J, L) + R3(FNHS, I + 1, J, L));
    // FAST: CGR = P4(K) * fRb
    d.fRb[o] = d.t1[o] / 2. / (R2(BNES, I, J) + R2(BNES, I + 1, J)) / (RLZI + 0.5 * MDR);
  }
  if (I >= 2 && J >= 2) {  // DRIFTP :232-237
    const double GPA1 = R3(FNIS, I, J, L) + R3(FNIS, I, J1, L) +
                        (R3(FNIS, I + 1, J1, L) + R3(FNIS, I + 1, J, L) - R3(FNIS, I - 1, J, L) - R3(FNIS, I - 1, J1, L)) * RLZI / 2. / MDR;
    const double GPA2 = RLZI / 4. / MDR * (R3(FNIS, I, J, L) + R3(FNIS, I, J1, L) - 2 * R3(FNHS, I, J, L) - 2 * R3(FNHS, I, J1, L)) *
                        (R2(BNES, I + 1, J1) + R2(BNES, I + 1, J) - R2(BNES, I - 1, J) - R2(BNES, I - 1, J1)) /
                        (R2(BNES, I, J) + R2(BNES, I, J1));
    d.G[o] = GPA1 + GPA2;
    d.sFp[o] = R3(FNHS, I, J, L) + R3(FNHS, I, J1, L);
    // FAST: CDriftP = fPa - w2(K) * fPb   (P2(I,K) = w2(K)/RLZ(I)**2)
    d.fPb[o] = (GPA1 + GPA2) / (R3(FNHS, I, J, L) + R3(FNHS, I, J1, L)) / (R2(BNES, I, J) + R2(BNES, I, J1)) / (RLZI * RLZI);
  } else {
    d.G[o] = 0;
    d.sFp[o] = 1;
    d.fPb[o] = 0;
  }
  if (I >= 2) {
    // DRIFTE :323-332, :340-341
    const double GPA = (1. - R3(FNIS, I, J, L) / 2. / R3(FNHS, I, J, L)) / R2(BNES, I, J);
    const double GPR1 = GPA * (R2(BNES, I + 1, J) - R2(BNES, I - 1, J)) / 2. / MDR;
    const double GPR2 = -R3(FNIS, I, J, L) / R3(FNHS, I, J, L) / RLZI;
    const double GPR3 = -(R3(FNIS, I + 1, J, L) - R3(FNIS, I - 1, J, L)) / 2. / MDR / R3(FNHS, I, J, L);
    const double GPP1 = GPA * (R2(BNES, I, J1) - R2(BNES, I, J0)) / 2. / DPHI;
    const double GPP2 = -(R3(FNIS, I, J1, L) - R3(FNIS, I, J0, L)) / 2. / DPHI / R3(FNHS, I, J, L);
    d.Gr[o] = GPR1 + GPR2 + GPR3;
    d.Gp[o] = GPP1 + GPP2;
    d.DRD2[o] = (R3(FNIS, I, J1, L) - R3(FNIS, I, J0, L)) / 2. / DPHI +
                (R3(FNIS, I, J, L) - 2 * R3(FNHS, I, J, L)) * (R2(BNES, I, J1) - R2(BNES, I, J0)) / 4 / R2(BNES, I, J) / DPHI;
    d.DPD2[o] = R3(FNIS, I, J, L) + (R3(FNIS, I + 1, J, L) - R3(FNIS, I - 1, J, L)) * RLZI / 2 / MDR +
                RLZI * (R3(FNIS, I, J, L) - 2 * R3(FNHS, I, J, L)) / 4 / MDR * (R2(BNES, I + 1, J) - R2(BNES, I - 1, J)) / R2(BNES, I, J);
    d.dBdt1[o] = R2(d.dBdt, I, J) * (1. - R3(FNIS, I, J, L) / 2. / R3(FNHS, I, J, L)) * RLZI / R2(BNES, I, J);
    d.dIdt1[o] = -R3(d.dIdt, I, J, L) * RLZI / R3(FNHS, I, J, L);
    // FAST: CDriftE = uE(K)*fEa + vE(K)*fEb  (EDOT(I,K) = uE(K)/RLZ(I), EDT1 = eK(K)/QS/(FNHS*RLZ*BNES))
    d.fEb[o] = (d.Gr[o] * d.DRD2[o] * RLZI - d.Gp[o] * d.DPD2[o]) / (R3(FNHS, I, J, L) * RLZI * R2(BNES, I, J)) / RLZI;
    // DRIFTMU :419-428, :432
    const double GMR1 = (R2(BNES, I + 1, J) - R2(BNES, I - 1, J)) / 4 / MDR / R2(BNES, I, J);
    const double GMR2 = 1 / RLZI;
    const double GMR3 = (R3(BOUNIS, I + 1, J, L) - R3(BOUNIS, I - 1, J, L)) / 2 / MDR / R3(BOUNIS, I, J, L);
    const double GMP1 = (R2(BNES, I, J1) - R2(BNES, I, J0)) / 4 / DPHI / R2(BNES, I, J);
    const double GMP2 = (R3(BOUNIS, I, J1, L) - R3(BOUNIS, I, J0, L)) / 2 / DPHI / R3(BOUNIS, I, J, L);
    d.Gmr[o] = GMR1 + GMR2 + GMR3;
    d.Gmp[o] = GMP1 + GMP2;
    d.DRM2[o] = (R3(BOUNIS, I, J1, L) - R3(BOUNIS, I, J0, L)) / 2 / DPHI +
                (R3(BOUNIS, I, J, L) - 2 * R3(BOUNHS, I, J, L)) * (R2(BNES, I, J1) - R2(BNES, I, J0)) / 4 / R2(BNES, I, J) / DPHI;
    d.DPM2[o] = R3(BOUNIS, I, J, L) + (R3(BOUNIS, I + 1, J, L) - R3(BOUNIS, I - 1, J, L)) * RLZI / 2 / MDR +
                (R3(BOUNIS, I, J, L) - 2 * R3(BOUNHS, I, J, L)) * RLZI / 4 / MDR * (R2(BNES, I + 1, J) - R2(BNES, I - 1, J)) / R2(BNES, I, J);
    d.dIbndt2[o] = R3(d.dIbndt, I, J, L) * RLZI / R3(BOUNIS, I, J, L);
  } else {
    d.Gr[o] = 0; d.Gp[o] = 0; d.DRD2[o] = 0; d.DPD2[o] = 0; d.dBdt1[o] = 0; d.dIdt1[o] = 0;
    d.Gmr[o] = 0; d.Gmp[o] = 0; d.DRM2[o] = 0; d.DPM2[o] = 0; d.dIbndt2[o] = 0;
    d.fEb[o] = 0;
  }
}

// DTs- and E-field-dependent pieces; run from DRIFTPARA when DTs or VT/EIR/EIP
// changed (src/ModRamDrift.f90:65-85 VR/P1/MUDOT, :118-127 CR, :236-239, :320-321)
__global__ void k_prep_step(RamDev d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.NPA * d.Pp) return;
  const int l = t / d.Pp, p = t - l * d.Pp;
  const size_t o = (size_t)l * d.Pp + p;
  if (p >= d.P) {
    d.CMUDOT[o] = 0; d.fEa[o] = 0; d.fMa[o] = 0; d.fMb[o] = 0;
    if (l == 0) { d.CR[p] = 0; d.pT1[p] = 0; d.pT3[p] = 0; d.DRD1[p] = 0; d.DPD1[p] = 0; d.dBdt2[p] = 0; d.fPa[p] = 0; }
    return;
  }
  const int j = p / d.NR, i = p - j * d.NR;
  const int I = i + 1, J = j + 1, L = l + 1;
  const int NT = d.NT;
  int J0 = J - 1; if (J == 1) J0 = NT - 1;
  int J1 = J + 1; if (J == NT) J1 = 2;
  const double MDR = d.MDR, DPHI = d.DPHI, DTs = d.DTs;
  const double RLZI = d.RLZ[i];
  const double *BNES = d.BNES, *VT = d.VT, *EIP = d.EIP, *EIR = d.EIR;
  if (l == 0) {
    const double VR = DTs / MDR / (RLZI + 0.5 * MDR) / 2 / DPHI;
    const double P1 = DTs / DPHI / 2 / MDR / RLZI;
    d.CR[p] = VR * (R2(VT, I, J0) + R2(VT, I + 1, J0) - R2(VT, I, J1) - R2(VT, I + 1, J1)) / (R2(BNES, I, J) + R2(BNES, I + 1, J)) +
              (R2(EIP, I, J) + R2(EIP, I + 1, J)) / (R2(BNES, I, J) + R2(BNES, I + 1, J)) * DTs / MDR;
    if (I >= 2) {
      if (J >= 2) {
        d.pT1[p] = (R2(VT, I + 1, J) + R2(VT, I + 1, J1) - R2(VT, I - 1, J) - R2(VT, I - 1, J1)) * P1;
        d.pT3[p] = (R2(EIR, I, J1) + R2(EIR, I, J)) / RLZI * DTs / DPHI;
        d.fPa[p] = (d.pT1[p] - d.pT3[p]) / (R2(BNES, I, J) + R2(BNES, I, J1)) + OME_EARTH * DTs / DPHI;
      } else {
        d.pT1[p] = 0; d.pT3[p] = 0; d.fPa[p] = 0;
      }
      d.DRD1[p] = (R2(EIP, I, J) * RLZI - (R2(VT, I, J1) - R2(VT, I, J0)) / 2. / DPHI) / R2(BNES, I, J);
      d.DPD1[p] = OME_EARTH * RLZI + ((R2(VT, I + 1, J) - R2(VT, I - 1, J)) / 2 / MDR - R2(EIR, I, J)) / R2(BNES, I, J);
      d.dBdt2[p] = R2(d.dBdt, I, J) / 2. / R2(BNES, I, J) * RLZI;
    } else {
      d.pT1[p] = 0; d.pT3[p] = 0; d.DRD1[p] = 0; d.DPD1[p] = 0; d.dBdt2[p] = 0; d.fPa[p] = 0;
    }
  }
  double CMUDOT = 0.;
  if (I >= 2 && L >= 2) {
    double MUDOT = 0.;
    if (L <= d.NPA - 1) {
      const double MUBOUN = d.MU[l] + 0.5 * d.WMU[l];
      MUDOT = (1. - MUBOUN * MUBOUN) * DTs / 2 / MUBOUN / RLZI;
    }
    CMUDOT = MUDOT * R3(d.BOUNIS, I, J, L) / R3(d.BOUNHS, I, J, L);
  }
  d.CMUDOT[o] = CMUDOT;
  // FAST-mode planes that depend on DTs / the E field (same expressions as the
  // l==0 block above, recomputed here so no thread reads another's output)
  if (I >= 2) {
    const double DRD1 = (R2(EIP, I, J) * RLZI - (R2(VT, I, J1) - R2(VT, I, J0)) / 2. / DPHI) / R2(BNES, I, J);
    const double DPD1 = OME_EARTH * RLZI + ((R2(VT, I + 1, J) - R2(VT, I - 1, J)) / 2 / MDR - R2(EIR, I, J)) / R2(BNES, I, J);
    const double dBdt2 = R2(d.dBdt, I, J) / 2. / R2(BNES, I, J) * RLZI;
    d.fEa[o] = (d.Gr[o] * DRD1 + d.Gp[o] * DPD1 + d.dBdt1[o] + d.dIdt1[o]) / RLZI;
    d.fMa[o] = -CMUDOT * (d.Gmr[o] * DRD1 + d.Gmp[o] * DPD1 + dBdt2 + d.dIbndt2[o]);
    d.fMb[o] = -CMUDOT * (d.Gmr[o] * d.DRM2[o] * RLZI - d.Gmp[o] * d.DPM2[o]) / (d.BOUNHSc[o] * RLZI * R2(BNES, I, J));
  } else {
    d.fEa[o] = 0; d.fMa[o] = 0; d.fMb[o] = 0;
  }
}

// =============================================================================
// coefficient tails (EXACT mode): the K-dependent part of each CDrift*, in the
// reference's operation order.
// =============================================================================
// CDriftR = CR + CGR3/(FNHS+FNHS)*P4/2./(BNES+BNES)/(RLZ+0.5*MDR)   :144-146
__device__ __forceinline__ double coef_r(double CR, double t1, double P4, double sB, double rl) {
  return CR + t1 * P4 / 2. / sB / rl;
}
// CDriftP :236-239
__device__ __forceinline__ double coef_p(double pT1, double P2, double G, double sFp, double pT3, double sBp, double OMEt) {
  return (pT1 - P2 * G / sFp - pT3) / sBp + OMEt;
}
// CDriftE :337-342
__device__ __forceinline__ double coef_e(double eK, double FNHS, double RLZI, double BNES, double QS, double DRD1, double DRD2,
                                         double DPD1, double DPD2, double Gr, double Gp, double dBdt1, double dIdt1, double EDOT) {
  const double EDT1 = eK / FNHS / RLZI / BNES / QS;
  const double DRDT = DRD1 + EDT1 * DRD2 * RLZI;
  const double DPDT = DPD1 - EDT1 * DPD2;
  return EDOT * (Gr * DRDT + Gp * DPDT + dBdt1 + dIdt1);
}
// CDriftMu :424-433
__device__ __forceinline__ double coef_mu(double epK, double BOUNHS, double RLZI, double BNES, double QS, double DRM1, double DRM2,
                                          double DPM1, double DPM2, double Gmr, double Gmp, double dBdt2, double dIbndt2,
                                          double CMUDOT) {
  const double EDT = epK / BOUNHS / RLZI / BNES / QS;
  const double DRDM = DRM1 + EDT * DRM2 * RLZI;
  const double DPDM = DPM1 - EDT * DPM2;
  return -CMUDOT * (Gmr * DRDM + Gmp * DPDM + dBdt2 + dIbndt2);
}

// =============================================================================
// Sweep kernels.  Common design:
//  * F2 is double-buffered per species: a sweep reads sp.F and writes sp.Fo
//    (every cell, untouched ones are copied through), so threads never race on
//    the limiter's 4-cell stencil and a line can be split between threads;
//  * the thread index runs along the contiguous (MLT,R) plane index p, so for
//    the strided sweeps (P, E, MU) every step of the walk is a coalesced row
//    access; each thread walks a SEGMENT of a line with a 4-value register
//    window, recomputing the one flux at the segment's lower edge;
//  * DRIFTR's lines lie along p itself: one thread per cell, the neighbour's
//    interface flux comes through a warp shuffle (31 cells + 1 halo lane / warp);
//  * blockIdx.y selects the species: one launch advances all species.
// =============================================================================

// warp-wide min -> one filtered atomicMin per warp (no CTA barrier: warps retire
// independently; the racy pre-read only filters, atomicMin decides)
__device__ __forceinline__ void warp_min_to(unsigned long long* dst, double v) {
  unsigned long long b = dbl_bits(v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
    b = t < b ? t : b;
  }
  if ((threadIdx.x & 31) == 0 && b < *((volatile unsigned long long*)dst)) atomicMin(dst, b);
}

// ---- DRIFTR pre-pass (src/ModRamDrift.f90:112-113,154-168): the line buffer's
// ghost cells F(NR+1:NR+2) are only rewritten on inflow lines, so an outflow
// line whose interface NR-1 is inflow reads the value left by the most recent
// inflow line in the reference loop order (K outer, L, J inner).  The inflow
// predicate (sign of CDriftR at I=NR) does not depend on F2, so the index of
// that line is precomputed once per DRIFTPARA: last[(k*NPA+l)*NT+j].
#define SCAN_TILE 1024
template <bool FAST>
__global__ void k_driftr_inflow(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                int* __restrict__ tilemax_all, int ntiles) {
  __shared__ int sm[32];
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  int* last = (int*)sp.last;
  const int n = d.NE * d.NPA * d.NT;
  const int t = blockIdx.x * SCAN_TILE + threadIdx.x;
  int v = -1;
  if (t < n) {
    const int j = t % d.NT;
    const int l = (t / d.NT) % d.NPA;
    const int k = t / (d.NT * d.NPA);
    const int i = d.NR - 1;
    const int p = j * d.NR + i;
    const double c = FAST ? fma(sp.P4[k], d.fRb[(size_t)l * d.Pp + p], d.CR[p])
                          : coef_r(d.CR[p], d.t1[(size_t)l * d.Pp + p], sp.P4[k], d.sB[p], d.RLZ[i] + 0.5 * d.MDR);
    v = (c < 0.0) ? t : -1;
    last[t] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = sm[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) tilemax_all[(size_t)(s0 + blockIdx.y) * ntiles + blockIdx.x] = v;
  }
}
// inclusive running max of last[] (tile-local scan + carry from the tile maxima)
__global__ void k_driftr_scan(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                              const int* __restrict__ tilemax_all, int ntiles) {
  __shared__ int sm[32];
  __shared__ int s_carry;
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  int* last = (int*)sp.last;
  const int* tilemax = tilemax_all + (size_t)(s0 + blockIdx.y) * ntiles;
  const int n = d.NE * d.NPA * d.NT;
  const int t = blockIdx.x * SCAN_TILE + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int c = -1;
  for (int q = threadIdx.x; q < (int)blockIdx.x; q += blockDim.x) c = max(c, tilemax[q]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
  if (lane == 0) sm[w] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = -1;
    for (int q = 0; q < 32; ++q) m = max(m, sm[q]);
    s_carry = m;
  }
  __syncthreads();
  int v = (t < n) ? last[t] : -1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = max(v, u);
  }
  __syncthreads();
  if (lane == 31) sm[w] = v;
  __syncthreads();
  int pre = s_carry;
  for (int q = 0; q < w; ++q) pre = max(pre, sm[q]);
  v = max(v, pre);
  if (t < n) {
    last[t] = v;
    // ghost cells F(NR+1), F(NR+2) this line will see (:154-168): its own boundary flux on an
    // inflow line, otherwise F(NR+1) as left by the most recent inflow line (0 if none yet)
    double g1 = 0.0, g2 = 0.0;
    if (v >= 0) {
      const int js = v % d.NT, ls = (v / d.NT) % d.NPA, ks = v / (d.NT * d.NPA);
      if (!d.outp[js * d.NR + d.NR - 1]) {
        const double fg = sp.FGEOS[((size_t)ls * d.NE + ks) * d.NT + js];
        const double fn = R3(d.FNHS, d.NR, js + 1, ls + 1);
        g1 = fg * d.CONF1 * fn;
        if (v == t) g2 = fg * d.CONF2 * fn;
      }
    }
    sp.ghost[2 * (size_t)t] = g1;
    sp.ghost[2 * (size_t)t + 1] = g2;
  }
}

// =============================================================================
// DRIFTR  (src/ModRamDrift.f90:95-198): lines lie along the contiguous plane
// index, so a thread owns one (MLT,R) position and walks KC energies of one
// pitch angle; the flux through the cell's lower face comes from the neighbour
// lane by warp shuffle (31 cells + 1 halo lane per warp).  Everything that does
// not depend on K (plane coefficients, boundary flags, index math) is hoisted.
// grid: x = tiles of 248 cells (8 warps x 31) of one plane, y = l*KG + kgroup, z = species
// =============================================================================
// warp-wide sum -> one partial per warp (no CTA barrier), fixed order => reproducible
__device__ __forceinline__ void warp_sum_to(double* part, size_t warp_index, double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) part[warp_index] = v;
}

template <bool FAST, bool MOM>
__global__ void __launch_bounds__(256) k_driftr(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, int KC,
                                                int KG, int l0) {
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int NR = d.NR, NT = d.NT, NE = d.NE, P = d.P, Pp = d.Pp;
  const int lq = blockIdx.y / KG, kg = blockIdx.y - lq * KG;
  const int l = l0 + lq;
  const int k0 = kg * KC, k1 = min(NE, k0 + KC);
  const int lane = threadIdx.x & 31;
  const int p = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 31 + lane - 1;   // 31 cells per warp + halo lane 0
  const bool inplane = (p >= 0) && (p < P);
  double cmax = 0.0, macc = 0.0;
  int i = -1, j = 0;
  double CRp = 0, gR = 0, t1 = 0, sB = 1, rl = 1;
  bool count = false;
  if (inplane) {
    j = p / NR;
    i = p - j * NR;
    CRp = d.CR[p];
    if (FAST) gR = d.fRb[(size_t)l * Pp + p];
    else { t1 = d.t1[(size_t)l * Pp + p]; sB = d.sB[p]; rl = d.RLZ[i] + 0.5 * d.MDR; }
    count = !d.outp[p];
  }
  const int I = i + 1;
  const bool edge = inplane && (i == 0 || i >= NR - 2);     // needs the line's boundary state
  const bool inmom = MOM && inplane && i >= 1 && l >= 1 && j <= NT - 2;
  const double beta = d.BetaLim;
  const double* F = sp.F + ((size_t)l * NE + k0) * Pp + p;
  double* Fo = sp.Fo + ((size_t)l * NE + k0) * Pp + p;
  for (int k = k0; k < k1; ++k, F += Pp, Fo += Pp) {
    double phi = 0.0, F0 = 0.0;
    if (inplane) {
      const double c = FAST ? fma(sp.P4[k], gR, CRp) : coef_r(CRp, t1, sp.P4[k], sB, rl);
      if (count) cmax = dmax(cmax, fabs(c));
      F0 = F[0];
      bool inflow = false;
      double g1 = 0.0, g2 = 0.0;                    // F(NR+1), F(NR+2): per line, from k_driftr_scan
      if (edge) {
        const int line = (k * d.NPA + l) * NT + j;
        inflow = (sp.last[line] == line);
        if (i >= NR - 2) { g1 = sp.ghost[2 * (size_t)line]; g2 = sp.ghost[2 * (size_t)line + 1]; }
      }
      const double Fm1 = (i >= 1) ? F[-1] : 0.0;
      const double Fp1 = (I + 1 <= NR) ? F[1] : g1;
      const double Fp2 = (I + 2 <= NR) ? F[2] : ((I + 2 == NR + 1) ? g1 : g2);
      double FB = FAST ? limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, fabs(c), beta) : limited_flux(Fm1, F0, Fp1, Fp2, c, c, beta);
      if (I == 1) FB = inflow ? Fp1 : 0.0;          // FBND(1) = F(2) | 0   (:155,:159)
      if (I == NR && !inflow) FB = F0;              // FBND(NR) = F(NR)     (:156)
      phi = c * FB;
    }
    const double phiPrev = __shfl_up_sync(0xffffffffu, phi, 1);
    if (inplane && lane >= 1) {
      double fn = F0;
      if (i >= 1) {
        fn = F0 - phi + phiPrev;                      // :186
        if (fn < 0.0) fn = 1E-15;
      }
      Fo[0] = fn;
      if (MOM && inmom && k >= 1) macc = fma(fn, d.WE[k] * d.EKEV[k], macc);
    }
  }
  warp_min_to(sp.dtw + 0, sp.aRP / dmax(cmax, 1E-10));
  if (MOM) {   // fused SUMRC (src/ModRamRun.f90:246-253) of the updated F2
    const size_t widx = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (threadIdx.x >> 5);
    warp_sum_to(sp.part, widx, inmom ? macc * d.WMU[l] : 0.0);
  }
}

// =============================================================================
// DRIFTP  (src/ModRamDrift.f90:204-279): periodic lines along MLT, segments of
// SEG cells of J=2..NT per thread.  FBND(1)=FBND(NT), F2(J=1)=F2(J=NT).
// grid: x = tiles of (k,i) pairs, y = l*nseg + seg, z = species
// =============================================================================
template <bool FAST>
__global__ void __launch_bounds__(128) k_driftp(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, int SEG,
                                                int nseg, int l0) {
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int NR = d.NR, NT = d.NT, Pp = d.Pp;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double cmax = 0.0;
  if (t < d.NE * NR) {
    const int lq = blockIdx.y / nseg, seg = blockIdx.y - lq * nseg;
    const int l = l0 + lq;
    const int k = t / NR, i = t - k * NR;
    const int plane = l * d.NE + k;
    const int ja = 2 + seg * SEG, jb = min(NT, ja + SEG - 1);
    const double* F = sp.F + (size_t)plane * Pp + i;   // F[(J-1)*NR]
    double* Fo = sp.Fo + (size_t)plane * Pp + i;
    if (i == 0) {
      for (int J = ja; J <= jb; ++J) Fo[(J - 1) * NR] = F[(J - 1) * NR];
      if (jb == NT) Fo[0] = F[0];
    } else {
      const double P2 = sp.P2[k * NR + i];
      const double w2 = sp.w2[k];
      const double beta = d.BetaLim, OMEt = sp.OMEt;
      const double f2 = F[NR], f3 = F[2 * NR];          // F(2), F(3): wrap-around values
      // coefficient at plane offset q = (J-1)*NR  (CDriftP(I,J,K,L), :236-239)
      const size_t lo = (size_t)l * Pp + i;
      auto coef = [&](int q) -> double {
        if (FAST) return fma(-w2, d.fPb[lo + q], d.fPa[q + i]);
        return coef_p(d.pT1[q + i], P2, d.G[lo + q], d.sFp[lo + q], d.pT3[q + i], d.sBp[q + i], OMEt);
      };
      auto lim = [&](double a, double b, double c_, double e, double cc) -> double {
        return FAST ? limited_flux_fast(a, b, c_, e, cc < 0.0, fabs(cc), beta) : limited_flux(a, b, c_, e, cc, cc, beta);
      };
      // interface NT-1: far-upwind difference F(2) - F(1) with the stored F(1) (limited_flux_num)
      const double numw = f2 - F[0];
      auto limw = [&](double a, double b, double c_, double cc) -> double {
        return FAST ? limited_flux_d(b, c_, b - a, c_ - b, numw, cc < 0.0, fabs(cc), beta) : limited_flux_num(a, b, c_, numw, cc, cc, beta);
      };
      // flux through the segment's lower edge: interface ja-1, or NT for ja==2 (:261-262)
      double prev;
      if (ja == 2) {
        const double c = coef((NT - 1) * NR);
        prev = c * lim(F[(NT - 2) * NR], F[(NT - 1) * NR], f2, f3, c);
      } else {
        const double c = coef((ja - 2) * NR);
        const double e = (ja + 1 <= NT) ? F[ja * NR] : f2;
        prev = c * ((ja - 1 == NT - 1) ? limw(F[(ja - 3) * NR], F[(ja - 2) * NR], F[(ja - 1) * NR], c)
                                       : lim(F[(ja - 3) * NR], F[(ja - 2) * NR], F[(ja - 1) * NR], e, c));
      }
      // rolling window F(J-1), F(J), F(J+1), F(J+2) with wrap J=NT+1->2, NT+2->3
      int q = (ja - 1) * NR;                       // offset of F(J)
      const int qNT = (NT - 1) * NR;
      double Fm1 = F[q - NR], F0 = F[q];
      double Fp1 = (ja + 1 <= NT) ? F[q + NR] : f2;
      double Fp2 = (ja + 2 <= NT) ? F[q + 2 * NR] : ((ja + 2 == NT + 1) ? f2 : f3);
      double cn = coef(q);
      double fnew = 0.0;
      for (int J = ja; J <= jb; ++J, q += NR) {
        // next step's inputs first (independent of this step's arithmetic)
        const int q3 = q + 3 * NR;
        const double Fp3 = (q3 <= qNT) ? F[q3] : ((q3 == qNT + NR) ? f2 : f3);
        const double cnn = coef(min(q + NR, qNT));
        const double c = cn;
        if (!d.outp[q + i]) cmax = dmax(cmax, fabs(c));
        const double cur = c * ((J == NT - 1) ? limw(Fm1, F0, Fp1, c) : lim(Fm1, F0, Fp1, Fp2, c));
        fnew = F0 - cur + prev;                         // :266
        if (fnew < 0.0) fnew = 1E-15;
        Fo[q] = fnew;
        prev = cur;
        Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = Fp3; cn = cnn;
      }
      if (jb == NT) Fo[0] = fnew;                       // :272
    }
  }
  warp_min_to(sp.dtw + 1, sp.aRP / dmax(cmax, 1E-10));
}

// =============================================================================
// DRIFTE  (src/ModRamDrift.f90:285-376): lines along energy, segments of SEG
// cells of K=1..NE per thread.  Ghosts F(1),F(0) from the relativistic
// extrapolation of F2(K=2) (:334-335), F(NE+1)=F(NE+2)=0 (:312-313).
// The first interface of a segment (K=ka-1, or K=1 for the first segment) only
// provides a flux: no cell is updated there, so it is peeled off the loop.
// grid: x = tiles of the plane index p, y = l*nseg + seg, z = species
// =============================================================================
template <bool FAST>
__global__ void __launch_bounds__(128) k_drifte(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, int SEG,
                                                int nseg, int l0) {
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int NE = d.NE, Pp = d.Pp;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double dtmin = 1.0e300;
  double mmax = 0.0;   // FAST: max over cells of max(|c|,1e-10)/DE(K)
  if (p < d.P) {
    const int lq = blockIdx.y / nseg, seg = blockIdx.y - lq * nseg;
    const int l = l0 + lq;
    const int ka = 1 + seg * SEG, kb = min(NE, ka + SEG - 1);
    const int i = p % d.NR;
    const double* F = sp.F + (size_t)l * NE * Pp + p;   // F[(K-1)*Pp]
    double* Fo = sp.Fo + (size_t)l * NE * Pp + p;
    if (i == 0) {
      for (int K = ka; K <= kb; ++K) Fo[(size_t)(K - 1) * Pp] = F[(size_t)(K - 1) * Pp];
    } else {
      const size_t o = (size_t)l * Pp + p;
      double FNHS = 0, Gr = 0, Gp = 0, DRD2 = 0, DPD2 = 0, dBdt1 = 0, dIdt1 = 0, DRD1 = 0, DPD1 = 0, BNES = 0, RLZI = 0, fA = 0, fB = 0;
      if (FAST) {
        fA = d.fEa[o]; fB = d.fEb[o];
      } else {
        FNHS = d.FNHSc[o]; Gr = d.Gr[o]; Gp = d.Gp[o]; DRD2 = d.DRD2[o]; DPD2 = d.DPD2[o]; dBdt1 = d.dBdt1[o]; dIdt1 = d.dIdt1[o];
        DRD1 = d.DRD1[p]; DPD1 = d.DPD1[p]; BNES = d.BNESc[p]; RLZI = d.RLZp[p];
      }
      const bool inside = !d.outp[p];
      const double QS = sp.QS, beta = d.BetaLim;
      const int K0 = max(ka - 1, 1);
      // rolling window F(K-1..K+2) + two prefetched values; pF walks ahead of the window
      double Fm1, F0, Fp1, Fp2, nxt, nx2;
      {
        double F1 = 0.0, Fz = 0.0;
        if (K0 <= 2) {
          const double f2 = F[(size_t)Pp];
          F1 = f2 * sp.GREL1 / sp.GREL2 * sp.sqrtA;
          Fz = F1 * sp.GRZERO / sp.GREL1 * sp.sqrtB;
        }
#define GETFK(K) (((K) > NE) ? 0.0 : (((K) >= 2) ? F[(size_t)((K)-1) * Pp] : (((K) == 1) ? F1 : Fz)))
        Fm1 = GETFK(K0 - 1); F0 = GETFK(K0); Fp1 = GETFK(K0 + 1); Fp2 = GETFK(K0 + 2);
        nxt = GETFK(K0 + 3); nx2 = GETFK(K0 + 4);
#undef GETFK
      }
      if (ka == 1) Fo[0] = F[0];                        // F2(K=1) is not advanced by DRIFTE
      const double* pF = F + (size_t)(K0 + 4) * Pp;     // -> F(K0+5)
      int nleft = NE - (K0 + 4);                        // loads still inside the array
      double* pO = Fo + (size_t)K0 * Pp;                // -> Fo(K0+1)
      double cprev, FBprev;
      if (FAST) {
        const double2* tab = (const double2*)sp.tabE + 2 * (K0 - 1);   // {uE,vE},{rDE,rWE} per K
        double floorr = 0.0;
        {  // peeled first interface K0: flux only
          const double2 uv = tab[0], rr = tab[1];
          const double c = fma(uv.y, fB, uv.x * fA);
          const double ac = fabs(c) * rr.x;                                        // |CDriftE|/DE(K)
          if (inside && seg == 0) mmax = ac;
          FBprev = c * limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, ac, beta);   // the flux c*FBND
          cprev = c;
          floorr = rr.x;
        }
        for (int K = K0 + 1; K <= kb; ++K) {
          tab += 2;
          const double nn = (nleft > 0) ? *pF : 0.0;
          pF += Pp; --nleft;
          Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = nxt; nxt = nx2; nx2 = nn;
          const double2 uv = tab[0], rr = tab[1];
          const double c = fma(uv.y, fB, uv.x * fA);
          const double ac = fabs(c) * rr.x;
          mmax = dmax(mmax, ac);
          const double FB = c * limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, ac, beta);
          double fn = fma(-(FB - FBprev), rr.y, F0);
          if (fn < 0.0) fn = 1E-15;
          *pO = fn;
          pO += Pp;
          FBprev = FB;
        }
        // the max(|c|,1e-10) floor of :344, hoisted: 1/DE is largest at the segment's first K
        mmax = inside ? dmax(mmax, 1E-10 * floorr) : 0.0;
      } else {
        const double* EDOT = sp.EDOT + i + (size_t)(K0 - 1) * d.NR;
        {
          const double c = coef_e(sp.eK[K0 - 1], FNHS, RLZI, BNES, QS, DRD1, DRD2, DPD1, DPD2, Gr, Gp, dBdt1, dIdt1, *EDOT);
          if (inside && seg == 0) dtmin = dmin(dtmin, sp.aE[K0 - 1] / dmax(fabs(c), 1E-10));
          FBprev = limited_flux(Fm1, F0, Fp1, Fp2, c, c / d.DE[K0 - 1], beta);
          cprev = c;
        }
        for (int K = K0 + 1; K <= kb; ++K) {
          EDOT += d.NR;
          const double nn = (nleft > 0) ? *pF : 0.0;
          pF += Pp; --nleft;
          Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = nxt; nxt = nx2; nx2 = nn;
          const double c = coef_e(sp.eK[K - 1], FNHS, RLZI, BNES, QS, DRD1, DRD2, DPD1, DPD2, Gr, Gp, dBdt1, dIdt1, *EDOT);
          if (inside) dtmin = dmin(dtmin, sp.aE[K - 1] / dmax(fabs(c), 1E-10));
          const double FB = limited_flux(Fm1, F0, Fp1, Fp2, c, c / d.DE[K - 1], beta);
          const double WEK = d.WE[K - 1];
          double fn = F0 - c / WEK * FB + cprev / WEK * FBprev;   // :364
          if (fn < 0.0) fn = 1E-15;
          *pO = fn;
          pO += Pp;
          cprev = c; FBprev = FB;
        }
      }
    }
  }
  if (FAST && mmax > 0.0) dtmin = sp.aRP / mmax;
  warp_min_to(sp.dtw + 2, dtmin);
}

// =============================================================================
// DRIFTMU  (src/ModRamDrift.f90:382-473): lines along pitch angle, segments of
// SEG cells of L=2..NPA-1 per thread; the last segment also closes the line with
// F2(NPA) = F2(NPA-1)*FNHS(NPA)*MU(NPA)/FNHS(NPA-1)/MU(NPA-1) (:466).
// grid: x = tiles of the plane index p, y = k*nseg + seg, z = species
// =============================================================================
template <bool FAST, bool MOM>
__global__ void __launch_bounds__(128) k_driftmu(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, int SEG,
                                                 int nseg, int k0) {
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int NE = d.NE, NPA = d.NPA, Pp = d.Pp;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double dtmin = 1.0e300;
  double mmax = 0.0;   // FAST: max over cells of max(|c|,1e-32)/DMU(L)
  double macc = 0.0;
  if (p < d.P) {
    const int kq = blockIdx.y / nseg, seg = blockIdx.y - kq * nseg;
    const int k = k0 + kq;
    const int la = 2 + seg * SEG, lb = min(NPA - 1, la + SEG - 1);
    const bool lastseg = (lb == NPA - 1);
    const int i = p % d.NR;
    const size_t LS = (size_t)NE * Pp;
    const double* F = sp.F + (size_t)k * Pp + p;   // F[(L-1)*LS]
    double* Fo = sp.Fo + (size_t)k * Pp + p;
    if (seg == 0) Fo[0] = F[0];                     // F2(L=1) is not advanced by DRIFTMU
    if (i == 0) {
      for (int L = la; L <= lb; ++L) Fo[(size_t)(L - 1) * LS] = F[(size_t)(L - 1) * LS];
      if (lastseg) Fo[(size_t)(NPA - 1) * LS] = F[(size_t)(NPA - 1) * LS];
    } else {
      double DRM1 = 0, DPM1 = 0, BNES = 0, RLZI = 0, dBdt2 = 0;
      if (!FAST) { DRM1 = d.DRD1[p]; DPM1 = d.DPD1[p]; BNES = d.BNESc[p]; RLZI = d.RLZp[p]; dBdt2 = d.dBdt2[p]; }
      const bool inside = !d.outp[p];
      const double QS = sp.QS, beta = d.BetaLim, epK = sp.epK[k], wM = sp.wM[k];
      // coefficient at plane offset o = (L-1)*Pp + p  (CDriftMu(I,J,K,L), :418-433)
      auto coef = [&](size_t o) -> double {
        if (FAST) return fma(wM, d.fMb[o], d.fMa[o]);
        return coef_mu(epK, d.BOUNHSc[o], RLZI, BNES, QS, DRM1, d.DRM2[o], DPM1, d.DPM2[o], d.Gmr[o], d.Gmp[o], dBdt2, d.dIbndt2[o], d.CMUDOT[o]);
      };
      // window for the first updated cell L=la: F(la-1), F(la), F(la+1), F(la+2); F(1)=F(2) (:414)
      const double* pF = F + (size_t)(la - 1) * LS;          // -> F(la)
      double Fm1 = (la >= 3) ? pF[-(ptrdiff_t)LS] : pF[0];
      double F0 = pF[0], Fp1 = pF[LS];
      double Fp2 = (la + 2 <= NPA) ? pF[2 * LS] : 0.0;
      double nxt = (la + 3 <= NPA) ? pF[3 * LS] : 0.0;
      pF += 4 * LS;                                          // -> F(la+4)
      int nleft = NPA - (la + 3);
      size_t o = (size_t)(la - 1) * Pp + p;                  // coefficient offset of L=la
      // flux through the segment's lower edge: CDriftMu(..,1)=0, FBND(1)=0 (:456-457) for la==2
      double cprev = 0.0, FBprev = 0.0;
      if (la > 2) {
        const double c = coef(o - Pp);
        const double Fm2 = (la >= 4) ? F[(size_t)(la - 3) * LS] : F[LS];   // F(la-2), F(1)=F(2)
        if (FAST) FBprev = c * limited_flux_fast(Fm2, Fm1, F0, Fp1, c < 0.0, fabs(c) * d.rDMU[la - 2], beta);
        else FBprev = limited_flux(Fm2, Fm1, F0, Fp1, c, c / d.DMU[la - 2], beta);
        cprev = c;
      }
      double cn = coef(o);
      double fnew = 0.0;
      double* pO = Fo + (size_t)(la - 1) * LS;
      for (int L = la; L <= lb; ++L) {
        o += Pp;
        const double nn = (nleft > 0) ? *pF : 0.0;
        pF += LS; --nleft;
        const double cnn = coef(o);                           // CDriftMu(..,L+1), L+1 <= NPA
        const double c = cn;
        if (FAST) {
          const double ac = fabs(c) * d.rDMU[L - 1];         // |CDriftMu|/DMU(L)
          mmax = dmax(mmax, ac);
          double FB;
          if (L <= NPA - 2) FB = c * limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, ac, beta);   // flux c*FBND
          else FB = c * Fp1;                                  // FBND(NPA-1) = F(NPA)  (:458)
          fnew = fma(-(FB - FBprev), d.rWMU[L - 1], F0);
          FBprev = FB;
        } else {
          if (inside) dtmin = dmin(dtmin, sp.aMU[L - 1] / dmax(fabs(c), 1E-32));
          double FB;
          if (L <= NPA - 2) FB = limited_flux(Fm1, F0, Fp1, Fp2, c, c / d.DMU[L - 1], beta);
          else FB = Fp1;
          const double WM = d.WMU[L - 1];
          fnew = F0 - c / WM * FB + cprev / WM * FBprev;      // :460
          cprev = c; FBprev = FB;
        }
        if (fnew < 0.0) fnew = 1E-15;
        *pO = fnew;
        pO += LS;
        if (MOM) macc = fma(fnew, d.WMU[L - 1], macc);
        Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = nxt; nxt = nn; cn = cnn;
      }
      if (FAST && !inside) mmax = 0.0;
      if (lastseg) {
        const double c = cn;                            // CDriftMu(..,NPA)
        if (FAST) { if (inside) mmax = dmax(mmax, fabs(c) * d.rDMU[NPA - 1]); }
        else if (inside) dtmin = dmin(dtmin, sp.aMU[NPA - 1] / dmax(fabs(c), 1E-32));
        const double fN = fnew * d.FNHSc[(size_t)(NPA - 1) * Pp + p] * d.MU[NPA - 1] / d.FNHSc[(size_t)(NPA - 2) * Pp + p] / d.MU[NPA - 2];
        *pO = fN;
        if (MOM) macc = fma(fN, d.WMU[NPA - 1], macc);
      }
      // fused SUMRC (src/ModRamRun.f90:246-253): I>=2 (here), K>=2, J<=NT-1
      if (MOM) macc = (k >= 1 && p < (d.NT - 1) * d.NR) ? macc * (d.WE[k] * d.EKEV[k]) : 0.0;
    }
  }
  if (MOM) {
    const size_t widx = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    warp_sum_to(sp.part, widx, macc);
  }
  if (FAST && mmax > 0.0) dtmin = sp.aRP / mmax;   // (the 1e-32 floor of :435 can never bind: DMU/1e-32 >> 1e4)
  warp_min_to(sp.dtw + 3, dtmin);
}

// =============================================================================
// moment helper: per-thread values -> per-CTA partial sums part[cta*NM+q]
// (fixed tree => run-to-run reproducible)
// =============================================================================
template <int NM>
__device__ __forceinline__ void cta_sum_to(double* part, int cta, double (&acc)[NM]) {
  __shared__ double s_sum[NM][32];
#pragma unroll
  for (int q = 0; q < NM; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_sum[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int q = 0; q < NM; ++q) {
      double v = (threadIdx.x < ((blockDim.x + 31) >> 5)) ? s_sum[q][threadIdx.x] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (threadIdx.x == 0) part[(size_t)cta * NM + q] = v;
    }
  }
}

// out[q] = sum_b part[b*nm + q]; grid = (nm, nspecies)
// (gap: result slots of q >= 1 are shifted by gap -- the fused column kernel produces slot 0 and slots 3..6)
__global__ void k_sum_final(const __grid_constant__ SpecPack pk, int s0, int nb, int nm, int slot0, int gap = 0) {
  __shared__ double sm[32];
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  const int q = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) acc += sp.part[(size_t)b * nm + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) ((double*)(sp.dt + 4))[slot0 + q + (q > 0 ? gap : 0)] = v;
  }
}

// =============================================================================
// SUMRC  (src/ModRamRun.f90:231-259): SETRC = sum_{I>=2,K>=2,L>=2,J<=NT-1}
// F2*WE(K)*WMU(L)*EKEV(K).  One CTA per (K,L) plane, then a second-stage tree
// over the plane partials: deterministic; differs from the reference's serial
// sum by summation order only (diagnostic quantity).
// grid: x = plane l*NE+k, y = species
// =============================================================================
__global__ void __launch_bounds__(256) k_sumrc_partial(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                       PlaneRange pr) {
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  const int lq = blockIdx.x / pr.nk;
  const int l = pr.l0 + lq, k = pr.k0 + (blockIdx.x - lq * pr.nk);
  const int plane = l * d.NE + k;
  double acc[1] = {0.0};
  if (k >= 1 && l >= 1) {
    const double* F = sp.F + (size_t)plane * d.Pp;
    const double w = d.WE[k], wm = d.WMU[l], e = d.EKEV[k];
    const int pend = (d.NT - 1) * d.NR;      // J <= NT-1
    for (int p = threadIdx.x; p < pend; p += blockDim.x) {
      if (p % d.NR == 0) continue;            // I >= 2
      const double WEIGHT = F[p] * w * wm;
      acc[0] += e * WEIGHT;
    }
  }
  cta_sum_to<1>(sp.part, blockIdx.x, acc);
}

// =============================================================================
// pointwise losses, one operator per launch (the per-routine ABI entry points):
// op 0 CHAREXCHANGE (src/ModRamLoss.f90:457-478, CHARGE of CEPARA :39-83 on the
// fly: the energy-only factor sv(K)=10**Y*V(S,K) is a host table), op 1 ATMOL
// (:485-507), op 2 WAVELO (src/ModRamWPI.f90:580-636, factor table from the host).
// grid: x = tiles of p, y = plane, z = species
// =============================================================================
__global__ void __launch_bounds__(256) k_loss(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, int op) {
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  const int l = plane / d.NE, k = plane - l * d.NE;
  if (p >= d.P || k < 1) return;
  const int i = p % d.NR;
  if (i < 1) return;
  double* Fp = sp.F + (size_t)plane * d.Pp + p;
  double f = *Fp;
  if (op == 0) {
    if (l < 1) return;
    const double ALPHA = sp.sv[k] * d.HDNSc[(size_t)l * d.Pp + p] * d.DTs;
    f = f * exp(-ALPHA);
  } else if (op == 1) {
    if (l + 1 < d.UPA[i]) return;
    f = f * pow(sp.ATLOS[k * d.NR + i], 1 / d.FNHSc[(size_t)l * d.Pp + p]);
  } else {
    if (l < 1) return;
    f = f * sp.wfac[(size_t)k * d.Pp + p];
  }
  *Fp = f;
}

// =============================================================================
// The palindromic middle of ram_run fused into one pass over F2
// (src/ModRamRun.f90:108-142 with the default flags):
//   [CHAREXCHANGE | WAVELO], SUMRC, ATMOL, SUMRC, ATMOL, SUMRC, [same], SUMRC
// 4 reference passes + 4 reductions -> one read and one write of F2; the four
// SETRC moments come out as per-plane partials.  doA: apply the species' first /
// last operator (CHAREX for ions; WAVELO for electrons when DoUseWPI is off);
// bit s of doA = species s.  Each factor is applied in the reference's order,
// so a cell's value is what the four separate passes would give.
// grid: x = plane, y = species (one CTA per plane, same tree as k_sumrc_partial)
// =============================================================================
template <bool FAST>
__global__ void __launch_bounds__(256) k_loss_mid(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                  int doA, PlaneRange pr) {
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  const int lq = blockIdx.x / pr.nk;
  const int l = pr.l0 + lq, k = pr.k0 + (blockIdx.x - lq * pr.nk);
  const int plane = l * d.NE + k;
  const bool ion = (sp.kind != 3);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (k >= 1) {                                           // every operator here acts on K>=2 only
    double* F = sp.F + (size_t)plane * d.Pp;
    const double w = d.WE[k], wm = d.WMU[l], e = d.EKEV[k];
    const bool useA = ((doA >> (s0 + blockIdx.y)) & 1) && (l >= 1);
    const int pmom = (d.NT - 1) * d.NR;
    for (int p = threadIdx.x; p < d.P; p += blockDim.x) {
      const int i = p % d.NR;
      if (i < 1) continue;                                // I >= 2
      double f = F[p];
      const bool mom = (l >= 1) && (p < pmom);
      double facA = 1.0;
      if (useA) {
        if (ion) {
          const double x = -(sp.sv[k] * d.HDNSc[(size_t)l * d.Pp + p] * d.DTs);
          facA = FAST ? fast_exp(x, d.exp2tab) : exp(x);
        }
        else facA = sp.wfac[(size_t)k * d.Pp + p];
        f = f * facA;
      }
      if (mom) acc[0] += e * (f * w * wm);
      if (l + 1 >= d.UPA[i]) {
        // ATLOS**(1/FNHS) (:502).  FAST: exp(log(ATLOS)/FNHS) with log(ATLOS) = -DTs/TAUB exact from the host
        const double a = FAST ? fast_exp(sp.xATL[k * d.NR + i] * d.rFNHS[(size_t)l * d.Pp + p], d.exp2tab)
                              : pow(sp.ATLOS[k * d.NR + i], 1 / d.FNHSc[(size_t)l * d.Pp + p]);
        f = f * a;
        if (mom) acc[1] += e * (f * w * wm);
        f = f * a;
        if (mom) acc[2] += e * (f * w * wm);
      } else if (mom) {
        const double term = e * (f * w * wm);
        acc[1] += term;
        acc[2] += term;
      }
      if (useA) f = f * facA;
      if (mom) acc[3] += e * (f * w * wm);
      F[p] = f;
    }
  }
  cta_sum_to<4>(sp.part, blockIdx.x, acc);
}

// =============================================================================
// WPADIF  (src/ModRamWPI.f90:643-714): implicit pitch-angle diffusion, Thomas
// recurrences along L per (J,I,K) line.  One thread per line; RK/RL live in
// shared memory [NPA][T] (conflict-free: T consecutive threads).  D = DA + DB.
// In place (a thread owns its whole line).
// =============================================================================
__global__ void k_wpadif(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, int k0, int nk) {
  extern __shared__ double smem[];
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  const int NE = d.NE, NPA = d.NPA, Pp = d.Pp, T = blockDim.x;
  double* RK = smem;             // [NPA][T]
  double* RL = smem + NPA * T;
  const long long t = (long long)blockIdx.x * T + threadIdx.x;
  const int tx = threadIdx.x;
  if (t >= (long long)nk * Pp) return;
  const int kq = (int)(t / Pp), p = (int)(t - (long long)kq * Pp);
  const int k = k0 + kq;
  const int i = p % d.NR;
  if (p >= d.P || i < 1 || k < 1) return;
  const size_t LS = (size_t)NE * Pp;
  double* F = sp.F + (size_t)k * Pp + p;
  const double* a = sp.DA + (size_t)k * Pp + p;  // [l*LS]
  const double* b = sp.DB + (size_t)k * Pp + p;
  const double DTs = d.DTs;
  unsigned long long viol = 0;
  RK[tx] = 0.;
  RL[tx] = -1.;
  double rkm = 0., rlm = -1.;
  double Dm = a[0] + b[0];   // D(L-1) for L=2
  for (int L = 2; L <= NPA - 1; ++L) {
    const double FACMU = d.FNHSc[(size_t)(L - 1) * Pp + p] * d.MU[L - 1];
    const double RP = F[(size_t)(L - 1) * LS] / FACMU;
    const double Dl = a[(size_t)(L - 1) * LS] + b[(size_t)(L - 1) * LS];
    double AN = Dl / d.DMU[L - 1];
    double GN = Dm / d.DMU[L - 2];
    AN = AN * DTs / FACMU / d.WMU[L - 1];
    GN = GN * DTs / FACMU / d.WMU[L - 1];
    const double BN = AN + GN;
    if (fabs(-1 - BN) < (fabs(AN) + fabs(GN))) ++viol;
    const double DENOM = BN + GN * rlm + 1;
    rkm = (RP + GN * rkm) / DENOM;
    rlm = -AN / DENOM;
    RK[(L - 1) * T + tx] = rkm;
    RL[(L - 1) * T + tx] = rlm;
    Dm = Dl;
  }
  double f = rkm / (1 + rlm);  // F2(NPA-1)
  {
    const double FM1 = d.FNHSc[(size_t)(NPA - 1) * Pp + p] * d.MU[NPA - 1];
    F[(size_t)(NPA - 1) * LS] = f * FM1;  // F2(NPA)=F2(NPA-1), then *FACMU(NPA)
    const double FM2 = d.FNHSc[(size_t)(NPA - 2) * Pp + p] * d.MU[NPA - 2];
    F[(size_t)(NPA - 2) * LS] = f * FM2;
  }
  for (int L = NPA - 2; L >= 1; --L) {
    f = RK[(L - 1) * T + tx] - RL[(L - 1) * T + tx] * f;
    const double FM = d.FNHSc[(size_t)(L - 1) * Pp + p] * d.MU[L - 1];
    F[(size_t)(L - 1) * LS] = f * FM;
  }
  if (viol) atomicAdd(sp.dt + 4 + 16, viol);
}

// =============================================================================
// ANISCH moments (src/ModRamRun.f90:343-415).  Stage 1: one thread per (K, p)
// does the pitch-angle sums SUME/SUMA in the reference's serial order (rows are
// coalesced across p) and the side effect F2(K,L=1)=F2(K,L=2) (:366).  Stage 2:
// one thread per p adds the energies band by band in the reference's order
// => PPERT/PPART are bit-identical to the oracle.
// =============================================================================
__global__ void __launch_bounds__(128) k_anisch_pa(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, int l0,
                                                   int nl) {
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int NE = d.NE, Pp = d.Pp;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (p >= Pp) return;
  const size_t t = (size_t)k * Pp + p;
  const int i = p % d.NR;
  if (p >= d.P || i < 1 || k < 1) { sp.tE[t] = 0; sp.tA[t] = 0; return; }
  const size_t LS = (size_t)NE * Pp;
  double* F = sp.F + (size_t)k * Pp + p;
  double f2 = 0.0;
  if (l0 == 0) {                 // the slab that owns L=1,2 (a slab always holds >= 2 pitch angles)
    f2 = F[LS];
    F[0] = f2;  // F2(S,I,J,K,1) = F2(S,I,J,K,2)
  }
  const int u = min(d.UPA[i] - 1, l0 + nl);
  double SUME = 0., SUMA = 0.;
  for (int L = l0 + 1; L <= u; ++L) {
    const double f = (L == 1) ? f2 : F[(size_t)(L - 1) * LS];
    const double ERNM = d.WMU[L - 1] / sp.FF[((size_t)(L - 1) * NE + k) * d.NR + i] / d.FNHSc[(size_t)(L - 1) * Pp + p];
    const double EPMA = ERNM * d.MU[L - 1] * d.MU[L - 1];
    const double EPME = ERNM - EPMA;
    SUME = SUME + f * EPME;
    SUMA = SUMA + f * EPMA;
  }
  sp.tE[t] = sp.EPP[k] * SUME;
  sp.tA[t] = sp.EPP[k] * SUMA;
}
// FAST variant of stage 1: the pitch-angle sum is split into NCH chunks of LCH cells (one
// thread each; threadIdx.y = chunk), combined in a fixed order through shared memory, and
// FFACTOR's separable form A(S,I,K)*MU(L) moves the division out of the loop:
//   SUME = rFFA(k,i) * sum_L f * wPE(L) / FNHS,   SUMA likewise with wPA.
// grid: x = tiles of 32 plane points, y = k, z = species; block = (32, nch)
__global__ void __launch_bounds__(512) k_anisch_pa_fast(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                        int l0, int nl, int LCH) {
  __shared__ double sE[16][32], sA[16][32];
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int NE = d.NE, Pp = d.Pp;
  const int p = blockIdx.x * 32 + threadIdx.x;
  const int k = blockIdx.y, ch = threadIdx.y, nch = blockDim.y;
  const int i = (p < d.P) ? p % d.NR : 0;
  const bool act = (p < d.P) && i >= 1 && k >= 1;
  double se = 0., sa = 0.;
  if (act) {
    const size_t LS = (size_t)NE * Pp;
    double* F = sp.F + (size_t)k * Pp + p;
    const int La = l0 + 1 + ch * LCH;                               // Fortran L range of this chunk
    const int Lb = min(min(d.UPA[i] - 1, l0 + nl), La + LCH - 1);
    int L = La;
    if (L == 1 && L <= Lb) {                                       // F2(S,I,J,K,1) = F2(S,I,J,K,2)  (:366)
      const double f = F[LS];
      F[0] = f;
      const double g = f * d.rFNHS[p];
      se = fma(g, d.wPE[0], se);
      sa = fma(g, d.wPA[0], sa);
      ++L;
    }
    // same summation order; unrolled so that a thread has four independent 8-byte loads in flight (the kernel is a pure
    // read of F2: at one load per warp it reached 2.1 TB/s, a third of HBM)
#pragma unroll 4
    for (; L <= Lb; ++L) {
      const double f = F[(size_t)(L - 1) * LS];
      const double g = f * d.rFNHS[(size_t)(L - 1) * Pp + p];
      se = fma(g, d.wPE[L - 1], se);
      sa = fma(g, d.wPA[L - 1], sa);
    }
  }
  sE[ch][threadIdx.x] = se;
  sA[ch][threadIdx.x] = sa;
  __syncthreads();
  if (ch == 0 && p < Pp) {
    for (int c = 1; c < nch; ++c) { se += sE[c][threadIdx.x]; sa += sA[c][threadIdx.x]; }
    const double c0 = act ? sp.EPP[k] * sp.rFFA[k * d.NR + i] : 0.0;
    sp.tE[(size_t)k * Pp + p] = c0 * se;
    sp.tA[(size_t)k * Pp + p] = c0 * sa;
  }
}
// kh0..kh4: 1-based inclusive upper K of the 5 energy bands (khi, :303,322)
__global__ void k_anisch_en(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, double RFAC, int kh0, int kh1,
                            int kh2, int kh3, int kh4) {
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.P) return;
  const int i = p % d.NR;
  if (i < 1) { sp.pper[p] = 0; sp.ppar[p] = 0; return; }
  const int khi[5] = {kh0, kh1, kh2, kh3, kh4};
  int klo = 2;
  double PT = 0., AT = 0.;
  for (int w = 0; w < 5; ++w) {
    double PPER = 0., PPAR = 0.;
#pragma unroll 8
    for (int K = klo; K <= khi[w]; ++K) {
      PPER = PPER + sp.tE[(size_t)(K - 1) * d.Pp + p];
      PPAR = PPAR + sp.tA[(size_t)(K - 1) * d.Pp + p];
    }
    PPAR = 2 * RFAC * PPAR;
    PPER = RFAC * PPER;
    klo = khi[w] + 1;
    PT = PT + PPER;
    AT = AT + PPAR;
  }
  sp.pper[p] = PT;
  sp.ppar[p] = AT;
}

// =============================================================================
// layout conversion between the host array F2(nS,NR,NT,NE,NPA) (species
// fastest) and F2dev; also the ram_run epilogue (src/ModRamRun.f90:186-201).
// =============================================================================
// stage: raw host image; one thread per device element of species s
// grid: x = tiles of p, y = plane
__global__ void k_f2_from_host(RamDev d, const double* __restrict__ stage, double* __restrict__ Fs, int s, int plane0 = 0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)plane0 + blockIdx.y;  // l*NE + k
  if (p >= d.Pp) return;
  double v = 0.0;
  if (p < d.P) v = stage[(plane * d.P + p) * d.nS + s];
  Fs[plane * d.Pp + p] = v;
}
__global__ void k_f2_to_host(RamDev d, double* __restrict__ stage, const double* __restrict__ Fs, int s, int plane0 = 0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)plane0 + blockIdx.y;
  if (p < d.P) stage[(plane * d.P + p) * d.nS + s] = Fs[plane * d.Pp + p];
}
// F2(:,:,NT,:,:) = F2(:,:,1,:,:), then F2 = 1e-31 where outsideMGNP == 1
// (src/ModRamRun.f90:186-201).  Only the J=1 row and the flagged columns are
// touched: thread q < NR handles the seam cell I=q+1 (it owns both its own cell
// and the J=NT copy, so there is no read/write race); thread q >= NR handles
// the (q-NR)-th flagged (I,J) column with 2 <= J <= NT-1.
// grid: x = tiles of q, y = plane, z = species
__global__ void __launch_bounds__(128) k_epilogue(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                  const int* __restrict__ outlist, int nout, PlaneRange pr) {
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int lq = blockIdx.y / pr.nk;
  const int plane = (pr.l0 + lq) * d.NE + pr.k0 + (blockIdx.y - lq * pr.nk);
  double* Fs = sp.F + (size_t)plane * d.Pp;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < d.NR) {
    const int pN = q + (d.NT - 1) * d.NR;
    const double v = Fs[q];
    Fs[pN] = d.outp[pN] ? 1.e-31 : v;
    if (d.outp[q]) Fs[q] = 1.e-31;
  } else if (q - d.NR < nout) {
    Fs[outlist[q - d.NR]] = 1.e-31;
  }
}
// FLUX = F2/FFACTOR/FNHS for I>=2,K>=2,L>=2,J<=NT-1 (src/ModRamRun.f90:210-221), host layout
__global__ void k_flux_to_host(RamDev d, SpecDev sp, double* __restrict__ stage) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = blockIdx.y;
  if (p >= d.P) return;
  const int l = (int)(plane / d.NE), k = (int)(plane - (size_t)l * d.NE);
  const int j = p / d.NR, i = p - j * d.NR;
  double v = 0.0;
  if (i >= 1 && k >= 1 && l >= 1 && j <= d.NT - 2)
    v = sp.F[plane * d.Pp + p] / sp.FF[((size_t)l * d.NE + k) * d.NR + i] / d.FNHSc[(size_t)l * d.Pp + p];
  stage[(plane * d.P + p) * d.nS + sp.S] = v;
}

// =============================================================================
// GEOSB (src/ModRamBoundary.f90:241-319, boundary 'LANL'): FGEOS of one species, in the [l][k][j] layout the DRIFTR
// kernels read, from the geosynchronous flux (NT,NE) (the file reader get_geomlt_flux stays on the host), the
// composition factor and FFACTOR at the outermost radius.  grid: x = tiles of (k, j), y = l
// =============================================================================
__global__ void k_geosb(RamDev d, const double* __restrict__ flux, double comp, const double* __restrict__ FF, double* __restrict__ FGEOS) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.NE * d.NT) return;
  const int k = t / d.NT, j = t - k * d.NT, l = blockIdx.y;
  const int u = d.UPA[d.NR - 1] - 1;                                   // Fortran L = 2 .. u
  double v = 0.0;
  if (l >= 1 && l + 1 <= u) {
    const double f = flux[(size_t)((j == 0) ? d.NT - 1 : j) + (size_t)d.NT * k] * comp;     // FluxLanl(1,:) = FluxLanl(nT,:), then * s_comp
    v = f * FF[((size_t)l * d.NE + k) * d.NR + (d.NR - 1)];
  }
  FGEOS[((size_t)l * d.NE + k) * d.NT + j] = v;
}
// get_electric_field (src/ModRamEField.f90:14-63): VT(NR+1,NT).  vols = 0: VTOL + (VTN - VTOL)*(t - TOLV)/DtEfi;
// vols = 1: Volland-Stern, AVS*(LZ*RE)**2*SIN(PHI - PHIOFS) with the sines from the host (libm).
__global__ void k_efield(RamDev d, int vols, const double* __restrict__ a, const double* __restrict__ b, double p0, double p1, double p2,
                         double* __restrict__ VT) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.NR1 * d.NT) return;
  const int j = t / d.NR1, i = t - j * d.NR1;
  if (!vols) VT[t] = a[t] + (b[t] - a[t]) * (p0 - p1) / p2;             // a = VTOL, b = VTN; p0 = t, p1 = TOLV, p2 = DtEfi
  else VT[t] = p0 * ((a[i] * p1) * (a[i] * p1)) * b[j];                 // a = LZ, b = sin(PHI - PHIOFS); p0 = AVS, p1 = RE
}
