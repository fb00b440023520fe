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
      const double LIM = fmax(fmin(beta * R, 1.0), fmin(R, beta));
      const double CORR = (-0.5 * (chat - sgn)) * X;
      FB = FUP + LIM * CORR;
    }
  }
  return FB;
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
  d.BOUNHSc[o] = R3(BOUNHS, I, J, L);
  d.HDNSc[o] = R3(d.HDNS, I, J, L);

  {  // DRIFTR :140-144 (all I, all J)
    const double CGR1 = R3(FNIS, I + 1, J1, L) + R3(FNIS, I, J1, L) - R3(FNIS, I + 1, J0, L) - R3(FNIS, I, J0, L);
    const double CGR2 = R2(BNES, I + 1, J1) + R2(BNES, I, J1) - R2(BNES, I + 1, J0) - R2(BNES, I, J0);
    const double CGR3 =
        CGR1 + (R3(FNIS, I + 1, J, L) + R3(FNIS, I, J, L) - 2 * R3(FNHS, I + 1, J, L) - 2 * R3(FNHS, I, J, L)) * CGR2 / 2. /
                   (R2(BNES, I + 1, J) + R2(BNES, I, J));
    d.t1[o] = CGR3 / (R3(FNHS, I, J, L) + R3(FNHS, I + 1, J, L));
  }
  if (I >= 2 && J >= 2) {  // DRIFTP :232-237
    const double GPA1 = R3(FNIS, I, J, L) + R3(FNIS, I, J1, L) +
                        (R3(FNIS, I + 1, J1, L) + R3(FNIS, I + 1, J, L) - R3(FNIS, I - 1, J, L) - R3(FNIS, I - 1, J1, L)) * RLZI / 2. / MDR;
    const double GPA2 = RLZI / 4. / MDR * (R3(FNIS, I, J, L) + R3(FNIS, I, J1, L) - 2 * R3(FNHS, I, J, L) - 2 * R3(FNHS, I, J1, L)) *
                        (R2(BNES, I + 1, J1) + R2(BNES, I + 1, J) - R2(BNES, I - 1, J) - R2(BNES, I - 1, J1)) /
                        (R2(BNES, I, J) + R2(BNES, I, J1));
    d.G[o] = GPA1 + GPA2;
    d.sFp[o] = R3(FNHS, I, J, L) + R3(FNHS, I, J1, L);
  } else {
    d.G[o] = 0;
    d.sFp[o] = 1;
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
    d.CMUDOT[o] = 0;
    if (l == 0) { d.CR[p] = 0; d.pT1[p] = 0; d.pT3[p] = 0; d.DRD1[p] = 0; d.DPD1[p] = 0; d.dBdt2[p] = 0; }
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
      } else {
        d.pT1[p] = 0; d.pT3[p] = 0;
      }
      d.DRD1[p] = (R2(EIP, I, J) * RLZI - (R2(VT, I, J1) - R2(VT, I, J0)) / 2. / DPHI) / R2(BNES, I, J);
      d.DPD1[p] = OME_EARTH * RLZI + ((R2(VT, I + 1, J) - R2(VT, I - 1, J)) / 2 / MDR - R2(EIR, I, J)) / R2(BNES, I, J);
      d.dBdt2[p] = R2(d.dBdt, I, J) / 2. / R2(BNES, I, J) * RLZI;
    } else {
      d.pT1[p] = 0; d.pT3[p] = 0; d.DRD1[p] = 0; d.DPD1[p] = 0; d.dBdt2[p] = 0;
    }
  }
  if (I >= 2 && L >= 2) {
    double MUDOT = 0.;
    if (L <= d.NPA - 1) {
      const double MUBOUN = d.MU[l] + 0.5 * d.WMU[l];
      MUDOT = (1. - MUBOUN * MUBOUN) * DTs / 2 / MUBOUN / RLZI;
    }
    d.CMUDOT[o] = MUDOT * R3(d.BOUNIS, I, J, L) / R3(d.BOUNHS, I, J, L);
  } else {
    d.CMUDOT[o] = 0;
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
// DRIFTR  (src/ModRamDrift.f90:95-198)
// Lines run along the contiguous device dimension, so a CTA stages KC whole
// (MLT,R) planes of one pitch angle in shared memory (coalesced loads), one
// thread walks each line with a 4-cell register window, results go back in
// place and are stored coalesced.  The energy-independent coefficient planes
// t1/CR/sB are staged once per CTA and reused by the KC energies.
// =============================================================================

// pre-pass: inflow flag of every line (sign of CDriftR at I=NR), reference
// line order (K outer, L, J inner) -> inflow[(k*NPA + l)*NT + j] = own index or -1
__global__ void k_driftr_inflow(RamDev d, SpecDev sp, int* __restrict__ last) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = d.NE * d.NPA * d.NT;
  if (t >= n) return;
  const int j = t % d.NT;
  const int l = (t / d.NT) % d.NPA;
  const int k = t / (d.NT * d.NPA);
  const int i = d.NR - 1;
  const int p = j * d.NR + i;
  const double rl = d.RLZ[i] + 0.5 * d.MDR;
  const double c = coef_r(d.CR[p], d.t1[(size_t)l * d.Pp + p], sp.P4[k], d.sB[p], rl);
  last[t] = (c < 0.0) ? t : -1;
}

// inclusive running max over `last` (single CTA; n <= a few 1e5)
__global__ void k_scan_last(int* __restrict__ last, int n) {
  __shared__ int sm[1024];
  const int T = blockDim.x, tid = threadIdx.x;
  const int chunk = (n + T - 1) / T;
  const int b = tid * chunk, e = min(n, b + chunk);
  int m = -1;
  for (int q = b; q < e; ++q) m = max(m, last[q]);
  sm[tid] = m;
  __syncthreads();
  for (int o = 1; o < T; o <<= 1) {
    int v = (tid >= o) ? sm[tid - o] : -1;
    __syncthreads();
    sm[tid] = max(sm[tid], v);
    __syncthreads();
  }
  int run = (tid > 0) ? sm[tid - 1] : -1;
  for (int q = b; q < e; ++q) {
    run = max(run, last[q]);
    last[q] = run;
  }
}

__global__ void k_driftr(RamDev d, SpecDev sp, const int* __restrict__ last, int KC) {
  extern __shared__ double smem[];
  const int NR = d.NR, NT = d.NT, P = d.P, Pp = d.Pp;
  const int NRc = NR | 1;          // odd row stride: conflict-free column walks
  const int NRf = (NR + 2) | 1;
  double* sT1 = smem;               // [NT][NRc]
  double* sCR = sT1 + NT * NRc;
  double* sSB = sCR + NT * NRc;
  double* sF = sSB + NT * NRc;      // [KC][NT][NRf]
  const int l = blockIdx.y;
  const int k0 = blockIdx.x * KC;
  const int kc = min(KC, d.NE - k0);
  const int tid = threadIdx.x, nth = blockDim.x;

  for (int p = tid; p < P; p += nth) {
    const int j = p / NR, i = p - j * NR;
    sT1[j * NRc + i] = d.t1[(size_t)l * Pp + p];
    sCR[j * NRc + i] = d.CR[p];
    sSB[j * NRc + i] = d.sB[p];
  }
  for (int q = tid; q < kc * P; q += nth) {
    const int kk = q / P, p = q - kk * P;
    const int j = p / NR, i = p - j * NR;
    sF[(kk * NT + j) * NRf + i] = sp.F[((size_t)l * d.NE + (k0 + kk)) * Pp + p];
  }
  __syncthreads();

  double cmax = 0.0;
  if (tid < kc * NT) {
    const int kk = tid / NT, j = tid - kk * NT;
    const int k = k0 + kk;
    double* F = sF + (kk * NT + j) * NRf;   // F[i] = F(I=i+1)
    const double* T1 = sT1 + j * NRc;
    const double* CRj = sCR + j * NRc;
    const double* SBj = sSB + j * NRc;
    const double P4 = sp.P4[k];
    const double hMDR = 0.5 * d.MDR;
    const double beta = d.BetaLim;
    const unsigned char* outp = d.outp + j * NR;

    const double cNR = coef_r(CRj[NR - 1], T1[NR - 1], P4, SBj[NR - 1], d.RLZ[NR - 1] + hMDR);
    const bool inflow = (cNR < 0.0);
    double g1, g2;  // F(NR+1), F(NR+2)
    if (inflow) {
      if (outp[NR - 1]) { g1 = 0.0; g2 = 0.0; }
      else {
        const double fg = sp.FGEOS[((size_t)l * d.NE + k) * NT + j];
        const double fn = R3(d.FNHS, NR, j + 1, l + 1);
        g1 = fg * d.CONF1 * fn;
        g2 = fg * d.CONF2 * fn;
      }
    } else {
      // ghost cells keep whatever the most recent inflow line (reference loop
      // order K,L,J) left in the line buffer; 0 if none yet  (:112-113,:154-168)
      g2 = 0.0;  // never read on an outflow line
      const int line = (k * d.NPA + l) * NT + j;
      const int src = last[line];
      if (src < 0) g1 = 0.0;
      else {
        const int js = src % NT, ls = (src / NT) % d.NPA, ks = src / (NT * d.NPA);
        if (d.outp[js * NR + NR - 1]) g1 = 0.0;
        else g1 = sp.FGEOS[((size_t)ls * d.NE + ks) * NT + js] * d.CONF1 * R3(d.FNHS, NR, js + 1, ls + 1);
      }
    }
    const int UR = inflow ? NR : NR - 1;
    // I = 1
    double c = coef_r(CRj[0], T1[0], P4, SBj[0], d.RLZ[0] + hMDR);
    if (!outp[0]) cmax = fmax(cmax, fabs(c));
    double Fm1 = 0.0, F0 = F[0], Fp1 = F[1], Fp2 = F[2];
    double prev = c * (inflow ? Fp1 : 0.0);  // CDriftR(1)*FBND(1)
    for (int I = 2; I <= NR; ++I) {
      Fm1 = F0; F0 = Fp1; Fp1 = Fp2;
      Fp2 = (I + 2 <= NR) ? F[I + 1] : ((I + 2 == NR + 1) ? g1 : g2);
      c = coef_r(CRj[I - 1], T1[I - 1], P4, SBj[I - 1], d.RLZ[I - 1] + hMDR);
      if (!outp[I - 1]) cmax = fmax(cmax, fabs(c));
      double FB;
      if (I <= UR) FB = limited_flux(Fm1, F0, Fp1, Fp2, c, c, beta);
      else FB = F0;  // outflow: FBND(NR) = F(NR)
      const double cur = c * FB;
      double fn = F0 - cur + prev;
      if (fn < 0.0) fn = 1E-15;
      F[I - 1] = fn;
      prev = cur;
    }
  }
  // CFL: min over cells of FracCFL*DTs/max(|c|,1e-10) == FracCFL*DTs/max over cells
  block_min_to(sp.dt + 0, sp.aRP / fmax(cmax, 1E-10));
  __syncthreads();
  for (int q = tid; q < kc * P; q += nth) {
    const int kk = q / P, p = q - kk * P;
    const int j = p / NR, i = p - j * NR;
    if (i >= 1) sp.F[((size_t)l * d.NE + (k0 + kk)) * Pp + p] = sF[(kk * NT + j) * NRf + i];
  }
}

// =============================================================================
// DRIFTP  (src/ModRamDrift.f90:204-279): periodic lines along MLT.  One thread
// per (L,K,I) line; at every J the warp reads NR-contiguous runs => coalesced
// without staging.  The wrap-around flux (FBND(1)=FBND(NT)) is computed first.
// =============================================================================
__global__ void k_driftp(RamDev d, SpecDev sp) {
  const int NR = d.NR, NT = d.NT, Pp = d.Pp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)d.NPA * d.NE * NR;
  double cmax = 0.0;
  if (t < n) {
    const int i = (int)(t % NR);
    const int plane = (int)(t / NR);
    const int k = plane % d.NE, l = plane / d.NE;
    if (i >= 1) {
      double* F = sp.F + (size_t)plane * Pp + i;  // F[j*NR] = F(J=j+1)
      const double* G = d.G + (size_t)l * Pp + i;
      const double* sFp = d.sFp + (size_t)l * Pp + i;
      const double P2 = sp.P2[k * NR + i];
      const double beta = d.BetaLim, OMEt = sp.OMEt;
      // interface J=NT first: window F(NT-1),F(NT),F(2),F(3)
      const double f2 = F[1 * NR], f3 = F[2 * NR];
      double cNT, phiNT;
      {
        const int pj = (NT - 1) * NR;
        cNT = coef_p(d.pT1[pj + i], P2, G[pj], sFp[pj], d.pT3[pj + i], d.sBp[pj + i], OMEt);
        const double FB = limited_flux(F[(NT - 2) * NR], F[(NT - 1) * NR], f2, f3, cNT, cNT, beta);
        phiNT = cNT * FB;
      }
      double Fm1 = F[0], F0 = f2, Fp1 = f3;
      double prev = phiNT;
      double fnew = 0.0;
      for (int J = 2; J <= NT; ++J) {
        // window for interface J: F(J-1)=Fm1, F(J)=F0, F(J1)=Fp1, F(J+2 wrapped)=Fp2
        double Fp2;
        if (J + 2 <= NT) Fp2 = F[(J + 1) * NR];
        else Fp2 = (J + 2 == NT + 1) ? f2 : f3;
        const int pj = (J - 1) * NR;
        double cur;
        if (J < NT) {
          const double c = coef_p(d.pT1[pj + i], P2, G[pj], sFp[pj], d.pT3[pj + i], d.sBp[pj + i], OMEt);
          if (!d.outp[pj + i]) cmax = fmax(cmax, fabs(c));
          cur = c * limited_flux(Fm1, F0, Fp1, Fp2, c, c, beta);
        } else {
          if (!d.outp[pj + i]) cmax = fmax(cmax, fabs(cNT));
          cur = phiNT;
        }
        fnew = F0 - cur + prev;
        if (fnew < 0.0) fnew = 1E-15;
        F[pj] = fnew;
        prev = cur;
        Fm1 = F0; F0 = Fp1; Fp1 = Fp2;
      }
      F[0] = fnew;  // F2(S,I,1,K,L) = F2(S,I,NT,K,L)
    }
  }
  block_min_to(sp.dt + 1, sp.aRP / fmax(cmax, 1E-10));
}

// =============================================================================
// DRIFTE  (src/ModRamDrift.f90:285-376): lines along energy.  One thread per
// (L,J,I) line; consecutive threads are consecutive in the contiguous plane
// index, so every K step is a coalesced row access.  Register window + one
// step of software prefetch.
// =============================================================================
__global__ void k_drifte(RamDev d, SpecDev sp) {
  const int NE = d.NE, Pp = d.Pp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double dtmin = 1.0e300;
  if (t < (long long)d.NPA * Pp) {
    const int l = (int)(t / Pp), p = (int)(t - (long long)l * Pp);
    const int i = p % d.NR;
    if (p < d.P && i >= 1) {
      const size_t o = (size_t)l * Pp + p;
      const double FNHS = d.FNHSc[o], Gr = d.Gr[o], Gp = d.Gp[o], DRD2 = d.DRD2[o], DPD2 = d.DPD2[o], dBdt1 = d.dBdt1[o],
                   dIdt1 = d.dIdt1[o];
      const double DRD1 = d.DRD1[p], DPD1 = d.DPD1[p], BNES = d.BNESc[p], RLZI = d.RLZp[p];
      const bool inside = !d.outp[p];
      const double QS = sp.QS, beta = d.BetaLim;
      double* F = sp.F + (size_t)l * NE * Pp + p;  // F[k*Pp] = F2(K=k+1)
      const double* EDOT = sp.EDOT + i;             // [k*NR]
      // ghost cells F(1), F(0)  (:334-335); F(NE+1)=F(NE+2)=0 (:312-313)
      const double f2 = F[(size_t)1 * Pp];
      const double F1 = f2 * sp.GREL1 / sp.GREL2 * sp.sqrtA;
      const double Fz = F1 * sp.GRZERO / sp.GREL1 * sp.sqrtB;
      double Fm1 = Fz, F0 = F1, Fp1 = f2, Fp2 = (NE >= 3) ? F[(size_t)2 * Pp] : 0.0;
      double nxt = (NE >= 4) ? F[(size_t)3 * Pp] : 0.0;  // F(K+3) prefetch
      double cprev = 0.0, FBprev = 0.0;
      for (int K = 1; K <= NE; ++K) {
        const double nn = (K + 4 <= NE) ? F[(size_t)(K + 3) * Pp] : 0.0;
        const double c = coef_e(sp.eK[K - 1], FNHS, RLZI, BNES, QS, DRD1, DRD2, DPD1, DPD2, Gr, Gp, dBdt1, dIdt1,
                                EDOT[(K - 1) * d.NR]);
        const double DEK = d.DE[K - 1];
        if (inside) dtmin = fmin(dtmin, sp.aE[K - 1] / fmax(fabs(c), 1E-10));
        const double FB = limited_flux(Fm1, F0, Fp1, Fp2, c, c / DEK, beta);
        if (K >= 2) {
          const double WEK = d.WE[K - 1];
          double fn = F0 - c / WEK * FB + cprev / WEK * FBprev;
          if (fn < 0.0) fn = 1E-15;
          F[(size_t)(K - 1) * Pp] = fn;
        }
        cprev = c; FBprev = FB;
        Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = nxt; nxt = nn;
      }
    }
  }
  block_min_to(sp.dt + 2, dtmin);
}

// =============================================================================
// DRIFTMU  (src/ModRamDrift.f90:382-473): lines along pitch angle.  One thread
// per (K,J,I) line, same coalescing argument as DRIFTE (row stride NE*Pp).
// =============================================================================
__global__ void k_driftmu(RamDev d, SpecDev sp) {
  const int NE = d.NE, NPA = d.NPA, Pp = d.Pp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double dtmin = 1.0e300;
  if (t < (long long)NE * Pp) {
    const int k = (int)(t / Pp), p = (int)(t - (long long)k * Pp);
    const int i = p % d.NR;
    if (p < d.P && i >= 1) {
      const size_t LS = (size_t)NE * Pp;  // stride between pitch angles
      double* F = sp.F + (size_t)k * Pp + p;  // F[l*LS] = F2(L=l+1)
      const double DRM1 = d.DRD1[p], DPM1 = d.DPD1[p], BNES = d.BNESc[p], RLZI = d.RLZp[p], dBdt2 = d.dBdt2[p];
      const bool inside = !d.outp[p];
      const double QS = sp.QS, beta = d.BetaLim, epK = sp.epK[k];
      // F(1) = F(2)  (:414)
      const double f2 = F[LS];
      double Fm1 = f2, F0 = f2, Fp1 = F[2 * LS], Fp2 = F[3 * LS];   // window for L=2
      double nxt = (NPA >= 5) ? F[4 * LS] : 0.0;
      double cprev = 0.0, FBprev = 0.0;  // CDriftMu(..,1)=0, FBND(1)=0 (:456-457)
      double fnew = 0.0;
      for (int L = 2; L <= NPA; ++L) {
        const double nn = (L + 4 <= NPA) ? F[(size_t)(L + 3) * LS] : 0.0;
        const size_t o = (size_t)(L - 1) * Pp + p;
        const double c = coef_mu(epK, d.BOUNHSc[o], RLZI, BNES, QS, DRM1, d.DRM2[o], DPM1, d.DPM2[o], d.Gmr[o], d.Gmp[o], dBdt2,
                                 d.dIbndt2[o], d.CMUDOT[o]);
        if (inside) dtmin = fmin(dtmin, sp.aMU[L - 1] / fmax(fabs(c), 1E-32));
        if (L <= NPA - 1) {
          double FB;
          if (L <= NPA - 2) FB = limited_flux(Fm1, F0, Fp1, Fp2, c, c / d.DMU[L - 1], beta);
          else FB = Fp1;  // FBND(NPA-1) = F(NPA)  (:458)
          const double WM = d.WMU[L - 1];
          fnew = F0 - c / WM * FB + cprev / WM * FBprev;
          if (fnew < 0.0) fnew = 1E-15;
          F[(size_t)(L - 1) * LS] = fnew;
          cprev = c; FBprev = FB;
        } else {
          // F2(NPA) = F2(NPA-1)*FNHS(NPA)*MU(NPA)/FNHS(NPA-1)/MU(NPA-1)  (:466)
          const double r = fnew * d.FNHSc[(size_t)(NPA - 1) * Pp + p] * d.MU[NPA - 1] / d.FNHSc[(size_t)(NPA - 2) * Pp + p] / d.MU[NPA - 2];
          F[(size_t)(NPA - 1) * LS] = r;
        }
        Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = nxt; nxt = nn;
      }
    }
  }
  block_min_to(sp.dt + 3, dtmin);
}

// =============================================================================
// pointwise losses.  CHAREXCHANGE (src/ModRamLoss.f90:457-478) with CEPARA's
// CHARGE evaluated on the fly (:39-83; the energy-only factor sv(K)=10**Y*V is a
// host table): F2 *= exp(-(sv*HDNS*DTs)).  ATMOL (:485-507): F2 *=
// ATLOS(I,K)**(1/FNHS) inside the loss cone.  WAVELO (src/ModRamWPI.f90:580-636):
// F2 *= exp(-DTs/TAU_LIF(I,J,K)), factor table built on the host.
// op: 0 CHAREX, 1 ATMOL, 2 WAVELO
// =============================================================================
__global__ void k_loss(RamDev d, SpecDev sp, int op, const double* __restrict__ wfac) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)d.NPA * d.NE * d.Pp;
  if (t >= n) return;
  const int p = (int)(t % d.Pp);
  const int plane = (int)(t / d.Pp);
  const int k = plane % d.NE, l = plane / d.NE;
  if (p >= d.P || k < 1) return;
  const int i = p % d.NR;
  if (i < 1) return;
  double f = sp.F[t];
  if (op == 0) {
    if (l < 1) return;
    const double ALPHA = sp.sv[k] * d.HDNSc[(size_t)l * d.Pp + p] * d.DTs;
    f = f * exp(-ALPHA);
  } else if (op == 1) {
    if (l + 1 < d.UPA[i]) return;
    f = f * pow(sp.ATLOS[k * d.NR + i], 1 / d.FNHSc[(size_t)l * d.Pp + p]);
  } else {
    if (l < 1) return;
    f = f * wfac[(size_t)k * d.Pp + p];
  }
  sp.F[t] = f;
}

// =============================================================================
// WPADIF  (src/ModRamWPI.f90:643-714): implicit pitch-angle diffusion, Thomas
// recurrences along L per (J,I,K) line.  One thread per line; RK/RL live in
// shared memory [NPA][T] (conflict-free: T consecutive threads).  D = DA + DB.
// =============================================================================
__global__ void k_wpadif(RamDev d, SpecDev sp, const double* __restrict__ DA, const double* __restrict__ DB,
                         unsigned long long* __restrict__ nviol) {
  extern __shared__ double smem[];
  const int NE = d.NE, NPA = d.NPA, Pp = d.Pp, T = blockDim.x;
  double* RK = smem;             // [NPA][T]
  double* RL = smem + NPA * T;
  const long long t = (long long)blockIdx.x * T + threadIdx.x;
  const int tx = threadIdx.x;
  if (t >= (long long)NE * Pp) return;
  const int k = (int)(t / Pp), p = (int)(t - (long long)k * Pp);
  const int i = p % d.NR;
  if (p >= d.P || i < 1 || k < 1) return;
  const size_t LS = (size_t)NE * Pp;
  double* F = sp.F + (size_t)k * Pp + p;
  const double* a = DA + (size_t)k * Pp + p;  // [l*LS]
  const double* b = DB + (size_t)k * Pp + p;
  const double DTs = d.DTs;
  unsigned long long viol = 0;
  // F(L) = F2/FACMU(L), F(1)=F(2); only F(L) is needed at step L
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
  if (viol) atomicAdd(nviol, viol);
}

// =============================================================================
// SUMRC  (src/ModRamRun.f90:231-259): SETRC = sum_{I>=2,K>=2,L>=2,J<=NT-1}
// F2*WE(K)*WMU(L)*EKEV(K).  Two-stage deterministic tree (fixed grid), so the
// result is run-to-run reproducible; it differs from the reference's serial
// sum by summation order only (~1e-15 relative, diagnostic quantity).
// =============================================================================
__global__ void k_sumrc_partial(RamDev d, SpecDev sp, double* __restrict__ part) {
  __shared__ double sm[32];
  const long long n = (long long)d.NPA * d.NE * d.Pp;
  double acc = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t % d.Pp);
    const int plane = (int)(t / d.Pp);
    const int k = plane % d.NE, l = plane / d.NE;
    if (p >= d.P || k < 1 || l < 1) continue;
    const int j = p / d.NR, i = p - j * d.NR;
    if (i < 1 || j > d.NT - 2) continue;
    const double WEIGHT = sp.F[t] * d.WE[k] * d.WMU[l];
    acc += d.EKEV[k] * WEIGHT;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
  }
}
// out[slot] = sum(part[0..n))  (single warp-multiple CTA, fixed order)
__global__ void k_sum_final(const double* __restrict__ part, int n, double* __restrict__ out) {
  __shared__ double sm[32];
  double acc = 0.0;
  for (int q = threadIdx.x; q < n; q += blockDim.x) acc += part[q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) *out = v;
  }
}

// =============================================================================
// ANISCH moments (src/ModRamRun.f90:343-415).  Stage 1: one thread per (K, p)
// does the pitch-angle sums SUME/SUMA in the reference's serial order (rows are
// coalesced across p) and the side effect F2(K,L=1)=F2(K,L=2) (:366).  Stage 2:
// one thread per p adds the energies band by band in the reference's order
// => PPERT/PPART are bit-identical to the oracle.
// =============================================================================
__global__ void k_anisch_pa(RamDev d, SpecDev sp, double* __restrict__ tE, double* __restrict__ tA) {
  const int NE = d.NE, Pp = d.Pp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)NE * Pp) return;
  const int k = (int)(t / Pp), p = (int)(t - (long long)k * Pp);
  const int i = p % d.NR;
  if (p >= d.P || i < 1 || k < 1) { tE[t] = 0; tA[t] = 0; return; }
  const size_t LS = (size_t)NE * Pp;
  double* F = sp.F + (size_t)k * Pp + p;
  const double f2 = F[LS];
  F[0] = f2;  // F2(S,I,J,K,1) = F2(S,I,J,K,2)
  const int u = d.UPA[i] - 1;
  double SUME = 0., SUMA = 0.;
  for (int L = 1; L <= u; ++L) {
    const double f = (L == 1) ? f2 : F[(size_t)(L - 1) * LS];
    const double ERNM = d.WMU[L - 1] / sp.FF[((size_t)(L - 1) * NE + k) * d.NR + i] / d.FNHSc[(size_t)(L - 1) * Pp + p];
    const double EPMA = ERNM * d.MU[L - 1] * d.MU[L - 1];
    const double EPME = ERNM - EPMA;
    SUME = SUME + f * EPME;
    SUMA = SUMA + f * EPMA;
  }
  tE[t] = sp.EPP[k] * SUME;
  tA[t] = sp.EPP[k] * SUMA;
}
// khi[5]: 1-based inclusive upper K of the 5 energy bands
__global__ void k_anisch_en(RamDev d, const double* __restrict__ tE, const double* __restrict__ tA, double RFAC, int kh0, int kh1,
                            int kh2, int kh3, int kh4, double* __restrict__ pper, double* __restrict__ ppar) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.P) return;
  const int i = p % d.NR;
  if (i < 1) { pper[p] = 0; ppar[p] = 0; return; }
  const int khi[5] = {kh0, kh1, kh2, kh3, kh4};
  int klo = 2;
  double PT = 0., AT = 0.;
  for (int w = 0; w < 5; ++w) {
    double PPER = 0., PPAR = 0.;
    for (int K = klo; K <= khi[w]; ++K) {
      PPER = PPER + tE[(size_t)(K - 1) * d.Pp + p];
      PPAR = PPAR + tA[(size_t)(K - 1) * d.Pp + p];
    }
    PPAR = 2 * RFAC * PPAR;
    PPER = RFAC * PPER;
    klo = khi[w] + 1;
    PT = PT + PPER;
    AT = AT + PPAR;
  }
  pper[p] = PT;
  ppar[p] = AT;
}

// =============================================================================
// layout conversion between the host array F2(nS,NR,NT,NE,NPA) (species
// fastest) and F2dev; also the ram_run epilogue (src/ModRamRun.f90:186-201).
// =============================================================================
// stage: raw host image; one thread per device element of species s
__global__ void k_f2_from_host(RamDev d, const double* __restrict__ stage, double* __restrict__ Fs, int s) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)d.NPA * d.NE * d.Pp;
  if (t >= n) return;
  const int p = (int)(t % d.Pp);
  const long long plane = t / d.Pp;  // l*NE + k
  double v = 0.0;
  if (p < d.P) v = stage[((size_t)plane * d.P + p) * d.nS + s];
  Fs[t] = v;
}
__global__ void k_f2_to_host(RamDev d, double* __restrict__ stage, const double* __restrict__ Fs, int s) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)d.NPA * d.NE * d.Pp;
  if (t >= n) return;
  const int p = (int)(t % d.Pp);
  const long long plane = t / d.Pp;
  if (p < d.P) stage[((size_t)plane * d.P + p) * d.nS + s] = Fs[t];
}
// F2(:,:,NT,:,:) = F2(:,:,1,:,:), then F2 = 1e-31 where outsideMGNP == 1.  The
// J=1 thread owns both its own cell and the J=NT copy (no read/write race).
__global__ void k_epilogue(RamDev d, double* __restrict__ Fs) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)d.NPA * d.NE * d.Pp;
  if (t >= n) return;
  const int p = (int)(t % d.Pp);
  if (p >= d.P) return;
  const int j = p / d.NR;
  if (j == d.NT - 1) return;
  const double v = Fs[t];
  if (j == 0) {
    const int pN = p + (d.NT - 1) * d.NR;
    Fs[t - p + pN] = d.outp[pN] ? 1.e-31 : v;
  }
  if (d.outp[p]) Fs[t] = 1.e-31;
}
// FLUX = F2/FFACTOR/FNHS for I>=2,K>=2,L>=2,J<=NT-1 (src/ModRamRun.f90:210-221), host layout
__global__ void k_flux_to_host(RamDev d, SpecDev sp, double* __restrict__ stage) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)d.NPA * d.NE * d.Pp;
  if (t >= n) return;
  const int p = (int)(t % d.Pp);
  const long long plane = t / d.Pp;
  if (p >= d.P) return;
  const int k = (int)(plane % d.NE), l = (int)(plane / d.NE);
  const int j = p / d.NR, i = p - j * d.NR;
  double v = 0.0;
  if (i >= 1 && k >= 1 && l >= 1 && j <= d.NT - 2)
    v = sp.F[t] / sp.FF[((size_t)l * d.NE + k) * d.NR + i] / d.FNHSc[(size_t)l * d.Pp + p];
  stage[((size_t)plane * d.P + p) * d.nS + sp.S] = v;
}
