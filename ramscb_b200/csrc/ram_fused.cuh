// FAST-mode fused kernels of the RAM step.
//
// The unfused sweeps (ram_kernels.cuh) stream F2 through HBM/L2 once per operator:
// 9 read+write passes per step, ~110 warp instructions per cell, issue slots half
// used because every cell waits on global loads.  Here F2 makes three round trips:
//
//   k_plane_rp<fwd>  DRIFTR, DRIFTP           on a shared-memory copy of a (K,L) plane
//   k_col_fused      DRIFTE, DRIFTMU, SUMRC, [CHAREX|WAVELO], ATMOL x2, ..., DRIFTMU, DRIFTE
//                    on a shared-memory block of PG plane positions x all (L,K)
//   k_plane_rp<rev>  DRIFTP, DRIFTR, SUMRC
//
// (the palindrome of src/ModRamRun.f90:70-175).  Per cell the arithmetic is the FAST
// arithmetic of the unfused kernels, operation for operation: F2 after a fused step is
// bit-identical to the unfused FAST path (tests/test_ram_parity_gpu.py); only the
// summation order of the SUMRC moments differs.  All updates are in place.
#pragma once
#include "ram_kernels.cuh"

// CTA-wide sum of NM per-thread values -> part[cta*NM + q] (fixed order => reproducible).
// `red` is shared scratch of NM*32 doubles.
template <int NM>
__device__ __forceinline__ void block_sum_to(double* part, size_t cta, double (&acc)[NM], double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int q = 0; q < NM; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[q * 32 + w] = v;
  }
  __syncthreads();
  if (threadIdx.x < NM) {
    double v = 0.0;
    for (int x = 0; x < nw; ++x) v += red[threadIdx.x * 32 + x];
    part[cta * NM + threadIdx.x] = v;
  }
}

// =============================================================================
// k_plane_rp: DRIFTR (src/ModRamDrift.f90:95-198) and DRIFTP (:204-279) of one
// (K,L) plane, back to back on a shared-memory copy; REV = reverse half step
// (DRIFTP first, then DRIFTR with the fused SUMRC of src/ModRamRun.f90:174).
// A CTA owns pitch angle l and KC consecutive energies; thread t owns the cells
// p = t + m*blockDim.x (m < NC) of every plane: their K-independent coefficients
// stay in registers, the next plane's values are prefetched while the current
// plane is advanced.  Both sweeps are cell-parallel: interface fluxes to shared
// memory, barrier, update.
// grid: x = energy chunk, y = l, z = species;  smem: 2*Pp + 2*NT doubles + NT ints
// =============================================================================
template <int NC, bool REV>
__global__ void __launch_bounds__(NC >= 4 ? 1024 : 256)
k_plane_rp(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0, int KC) {
  extern __shared__ double smem[];
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int NR = d.NR, NT = d.NT, NE = d.NE, P = d.P, Pp = d.Pp;
  const int T = blockDim.x, tid = threadIdx.x;
  double* sF = smem;                   // the plane
  double* sPhi = smem + Pp;            // interface fluxes
  double* sG = smem + 2 * Pp;          // ghost cells F(NR+1), F(NR+2) of the NT lines
  int* sIn = (int*)(sG + 2 * NT);      // inflow flag of the lines
  const int l = blockIdx.y;
  const int k0 = blockIdx.x * KC, k1 = min(NE, k0 + KC);
  const double beta = d.BetaLim;

  double CRc[NC], gR[NC], pa[NC], pb[NC];
  int ij[NC];                          // i | j << 12 | counted << 30; -1: no cell
#pragma unroll
  for (int m = 0; m < NC; ++m) {
    const int p = tid + m * T;
    ij[m] = -1;
    CRc[m] = gR[m] = pa[m] = pb[m] = 0.0;
    if (p < P) {
      const int j = p / NR, i = p - j * NR;
      ij[m] = i | (j << 12) | (d.outp[p] ? 0 : (1 << 30));
      CRc[m] = d.CR[p];
      gR[m] = d.fRb[(size_t)l * Pp + p];
      pa[m] = d.fPa[p];
      pb[m] = d.fPb[(size_t)l * Pp + p];
    }
  }
  double* Fg = sp.F + ((size_t)l * NE + k0) * Pp;
  double Fn[NC];
#pragma unroll
  for (int m = 0; m < NC; ++m) Fn[m] = (ij[m] >= 0) ? Fg[tid + m * T] : 0.0;
  int line = (k0 * d.NPA + l) * NT + tid;   // threads < NT carry the line state
  int inN = 0;
  double g1N = 0.0, g2N = 0.0;
  if (tid < NT) { inN = (sp.last[line] == line); g1N = sp.ghost[2 * (size_t)line]; g2N = sp.ghost[2 * (size_t)line + 1]; }
  double cmaxR = 0.0, cmaxP = 0.0, macc = 0.0;

  for (int k = k0; k < k1; ++k, Fg += Pp) {
    double Fc[NC];
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      Fc[m] = Fn[m];
      if (ij[m] >= 0) sF[tid + m * T] = Fc[m];
    }
    if (tid < NT) { sIn[tid] = inN; sG[2 * tid] = g1N; sG[2 * tid + 1] = g2N; }
    __syncthreads();
    if (k + 1 < k1) {                  // next plane: in flight during this plane's arithmetic
#pragma unroll
      for (int m = 0; m < NC; ++m) Fn[m] = (ij[m] >= 0) ? Fg[Pp + tid + m * T] : 0.0;
      if (tid < NT) {
        line += d.NPA * NT;
        inN = (sp.last[line] == line); g1N = sp.ghost[2 * (size_t)line]; g2N = sp.ghost[2 * (size_t)line + 1];
      }
    }
    const double P4k = sp.P4[k], w2k = sp.w2[k];
    const double wk = (REV && k >= 1) ? d.WE[k] * d.EKEV[k] : 0.0;

    // ---- DRIFTP: cells J=2..NT, I>=2; F(NT+1)=F(2), F(NT+2)=F(3), FBND(1)=FBND(NT) (:246-262)
    auto driftp = [&]() {
      double cur[NC];
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        const int p = tid + m * T, i = ij[m] & 4095, j = (ij[m] >> 12) & 4095;
        cur[m] = 0.0;
        if (ij[m] >= 0 && i >= 1 && j >= 1) {
          const double c = fma(-w2k, pb[m], pa[m]);
          if (REV && (ij[m] >> 30)) cmaxP = dmax(cmaxP, fabs(c));
          const double F0 = REV ? Fc[m] : sF[p];
          const double Fm1 = sF[p - NR];
          const double f2 = sF[NR + i], f3 = sF[2 * NR + i];
          const double Fp1 = (j + 1 <= NT - 1) ? sF[p + NR] : f2;
          const double Fp2 = (j + 2 <= NT - 1) ? sF[p + 2 * NR] : ((j + 2 == NT) ? f2 : f3);
          if (!REV) Fc[m] = F0;
          cur[m] = c * limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, fabs(c), beta);
          sPhi[p] = cur[m];
        }
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        const int p = tid + m * T, i = ij[m] & 4095, j = (ij[m] >> 12) & 4095;
        if (ij[m] >= 0 && i >= 1 && j >= 1) {
          const double prev = (j == 1) ? sPhi[(NT - 1) * NR + i] : sPhi[p - NR];
          double fnew = Fc[m] - cur[m] + prev;               // :266
          if (fnew < 0.0) fnew = 1E-15;
          if (REV) {
            sF[p] = fnew;
            if (j == NT - 1) sF[i] = fnew;                   // F2(J=1) = F2(J=NT)  (:272)
          } else {
            Fg[p] = fnew;
            if (j == NT - 1) Fg[i] = fnew;
          }
        }
      }
    };
    // ---- DRIFTR: all cells; FBND(1), FBND(NR) and the ghost cells from the line state (:154-168)
    auto driftr = [&]() {
      double phi[NC], F0v[NC];
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        const int p = tid + m * T, i = ij[m] & 4095, j = (ij[m] >> 12) & 4095;
        phi[m] = 0.0;
        F0v[m] = 0.0;
        if (ij[m] >= 0) {
          const double c = fma(P4k, gR[m], CRc[m]);
          if (REV && (ij[m] >> 30)) cmaxR = dmax(cmaxR, fabs(c));
          const double F0 = REV ? sF[p] : Fc[m];
          F0v[m] = F0;
          const int I = i + 1;
          const double g1 = sG[2 * j], g2 = sG[2 * j + 1];
          const bool inflow = sIn[j] != 0;
          const double Fm1 = (i >= 1) ? sF[p - 1] : 0.0;
          const double Fp1 = (I + 1 <= NR) ? sF[p + 1] : g1;
          const double Fp2 = (I + 2 <= NR) ? sF[p + 2] : ((I + 2 == NR + 1) ? g1 : g2);
          double FB = limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, fabs(c), beta);
          if (I == 1) FB = inflow ? Fp1 : 0.0;
          if (I == NR && !inflow) FB = F0;
          phi[m] = c * FB;
          sPhi[p] = phi[m];
        }
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        const int p = tid + m * T, i = ij[m] & 4095, j = (ij[m] >> 12) & 4095;
        if (ij[m] >= 0 && i >= 1) {
          double fn = F0v[m] - phi[m] + sPhi[p - 1];         // :186
          if (fn < 0.0) fn = 1E-15;
          if (REV) {
            Fg[p] = fn;
            if (l >= 1 && j <= NT - 2) macc = fma(fn, wk, macc);
          } else {
            sF[p] = fn;
          }
        }
      }
    };
    if (REV) { driftp(); __syncthreads(); driftr(); }
    else { driftr(); __syncthreads(); driftp(); }
    // the next iteration's first shared-memory writes go to sF / the line state, whose readers
    // are all behind the barrier inside the second sweep; sPhi is rewritten only after the
    // next iteration's first barrier.
  }
  if (REV) {
    warp_min_to(sp.dtw + 0, sp.aRP / dmax(cmaxR, 1E-10));
    warp_min_to(sp.dtw + 1, sp.aRP / dmax(cmaxP, 1E-10));
    __syncthreads();
    double acc[1] = {macc * d.WMU[l]};
    block_sum_to<1>(sp.part, (size_t)blockIdx.y * gridDim.x + blockIdx.x, acc, sPhi);
  }
}

// =============================================================================
// k_col_fused: everything of the step that couples only energy and pitch angle,
// on a shared-memory block of PG consecutive plane positions x all (L,K):
//   DRIFTE, DRIFTMU, SUMRC | [CHAREXCHANGE|WAVELO], SUMRC, ATMOL, SUMRC, ATMOL,
//   SUMRC, [same], SUMRC | DRIFTMU, DRIFTE              (src/ModRamRun.f90:75-173)
// Block layout sT[l][k][pp] with row stride NEs (odd: the energy walks of 4 pitch
// angles x PG positions of a half-warp then hit 16 different 8-byte banks).
// Sweeps: one thread per line segment with a 4-value register window (as the
// unfused kernels); a segment's foreign halo cells are read before the barrier
// that precedes the in-place walk.  With more lines than threads, whole lines.
// grid: x = block of PG positions, y = species
// =============================================================================
struct ColCfg {
  int NEs;            // padded energy stride of the block
  int nsegE, segE;    // DRIFTE: segments per line, cells per segment
  int nsegM, segM;    // DRIFTMU
  int doA;            // bit s: species s applies its first/last loss operator
};

template <int PG>
__global__ void __launch_bounds__(1024) k_col_fused(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                    ColCfg cfg) {
  extern __shared__ double smem[];
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  const int NR = d.NR, NT = d.NT, NE = d.NE, NPA = d.NPA, P = d.P, Pp = d.Pp;
  const int T = blockDim.x, tid = threadIdx.x;
  const int p0 = blockIdx.x * PG;
  const int NEs = cfg.NEs, RS = NEs * PG;
  double* sT = smem;                         // [NPA][NEs][PG]
  double* sEa = sT + (size_t)NPA * RS;       // [NPA][PG] each
  double* sEb = sEa + NPA * PG;
  double* sMa = sEb + NPA * PG;
  double* sMb = sMa + NPA * PG;
  double* sH = sMb + NPA * PG;               // HDNS (charge exchange)
  double* sRF = sH + NPA * PG;               // 1/FNHS (ATMOL)
  double* sTab = sRF + NPA * PG;             // [NE][4] uE, vE, 1/DE, 1/WE
  double* sWM = sTab + 4 * NE;               // [NE] each
  double* sSV = sWM + NE;
  double* sWE = sSV + NE;
  double* sEK = sWE + NE;
  double* sX = sEK + NE;                     // [NE][PG] log(ATLOS) at this block's radii
  double* sW = sX + NE * PG;                 // [NE][PG] WAVELO factor
  double* sRD = sW + NE * PG;                // [NPA] each: 1/DMU, 1/WMU, WMU
  double* sRW = sRD + NPA;
  double* sWMU = sRW + NPA;
  double* sRed = sWMU + NPA;                 // [5][32]

  // ---- stage the block and its coefficient tables ----------------------------
  {
    constexpr int H = PG / 2;                // 16-byte items per (l,k) row
    const int rowItems = NE * H;
    int l = tid / rowItems, r = tid - l * rowItems;
    const int dl = T / rowItems, dr = T - dl * rowItems;
    while (l < NPA) {
      const int k = r / H, hh = r - k * H;
      const double2 v = *(const double2*)(sp.F + ((size_t)l * NE + k) * Pp + p0 + 2 * hh);
      *(double2*)(sT + (size_t)l * RS + k * PG + 2 * hh) = v;
      r += dr; l += dl;
      if (r >= rowItems) { r -= rowItems; ++l; }
    }
    for (int t = tid; t < NPA * PG; t += T) {
      const int l2 = t / PG, pp = t - l2 * PG;
      const size_t o = (size_t)l2 * Pp + p0 + pp;
      sEa[t] = d.fEa[o]; sEb[t] = d.fEb[o]; sMa[t] = d.fMa[o]; sMb[t] = d.fMb[o];
      sH[t] = d.HDNSc[o]; sRF[t] = d.rFNHS[o];
    }
    for (int t = tid; t < 4 * NE; t += T) sTab[t] = sp.tabE[t];
    for (int t = tid; t < NE; t += T) { sWM[t] = sp.wM[t]; sSV[t] = sp.sv[t]; sWE[t] = d.WE[t]; sEK[t] = d.EKEV[t]; }
    for (int t = tid; t < NE * PG; t += T) {
      const int k = t / PG, pp = t - k * PG;
      const int p = min(p0 + pp, P - 1);
      sX[t] = sp.xATL[k * NR + p % NR];
      sW[t] = sp.wfac[(size_t)k * Pp + p0 + pp];
    }
    for (int t = tid; t < NPA; t += T) { sRD[t] = d.rDMU[t]; sRW[t] = d.rWMU[t]; sWMU[t] = d.WMU[t]; }
  }
  __syncthreads();

  const double beta = d.BetaLim;
  double mmaxE = 0.0, mmaxM = 0.0;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // SUMRC after DRIFTMU, and the four of the loss block

  // ---- DRIFTE (src/ModRamDrift.f90:285-376) ------------------------------------
  auto drifte = [&](const bool cfl) {
    const int ntask = NPA * PG * cfg.nsegE;
    const int nround = (ntask + T - 1) / T;
    for (int rd = 0; rd < nround; ++rd) {
      const int e = tid + rd * T;
      const int pp = e % PG, lq = e / PG;
      const int l = lq % NPA, seg = lq / NPA;
      const int p = p0 + pp;
      const bool act = (e < ntask) && (p < P) && (p % NR != 0);
      const int ka = 1 + seg * cfg.segE, kb = min(NE, ka + cfg.segE - 1);
      const int K0 = max(ka - 1, 1);
      double* col = sT + (size_t)l * RS + pp;           // F(K) at col[(K-1)*PG]
      double Fm1 = 0, F0 = 0, Fp1 = 0, Fp2 = 0, hi1 = 0, hi2 = 0;
      if (act) {
        double F1 = 0.0, Fz = 0.0;
        if (K0 <= 2) {
          const double f2 = col[PG];
          F1 = f2 * sp.GREL1 / sp.GREL2 * sp.sqrtA;
          Fz = F1 * sp.GRZERO / sp.GREL1 * sp.sqrtB;
        }
#define GETFK(K) (((K) > NE) ? 0.0 : (((K) >= 2) ? col[((K)-1) * PG] : (((K) == 1) ? F1 : Fz)))
        Fm1 = GETFK(K0 - 1); F0 = GETFK(K0); Fp1 = GETFK(K0 + 1); Fp2 = GETFK(K0 + 2);
        hi1 = GETFK(kb + 1); hi2 = GETFK(kb + 2);
#undef GETFK
      }
      if (cfg.nsegE > 1) __syncthreads();               // halo reads before anybody's in-place walk
      if (act) {
        const bool inside = !d.outp[p];
        const double fA = sEa[l * PG + pp], fB = sEb[l * PG + pp];
        const double* tab = sTab + 4 * (K0 - 1);
        double FBprev, mmax = 0.0;
        const double floorr = tab[2];
        {                                               // peeled first interface K0: flux only
          const double c = fma(tab[1], fB, tab[0] * fA);
          const double ac = fabs(c) * tab[2];
          if (inside && seg == 0) mmax = ac;
          FBprev = c * limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, ac, beta);
        }
        double* pO = col + (size_t)K0 * PG;             // -> F(K0+1)
        auto step = [&](const double nn) {
          tab += 4;
          Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = nn;
          const double c = fma(tab[1], fB, tab[0] * fA);
          const double ac = fabs(c) * tab[2];
          if (cfl) mmax = dmax(mmax, ac);
          const double FB = c * limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, ac, beta);
          double fn = fma(-(FB - FBprev), tab[3], F0);
          if (fn < 0.0) fn = 1E-15;
          *pO = fn;
          pO += PG;
          FBprev = FB;
        };
        int K = K0 + 1;
        for (; K + 2 <= kb; ++K) step(pO[2 * PG]);      // F(K+2): own cell, not yet rewritten
        if (K + 1 <= kb) { step(hi1); ++K; }            // K = kb-1: F(kb+1)
        if (K <= kb) step(hi2);                         // K = kb:   F(kb+2)
        if (cfl) mmaxE = dmax(mmaxE, inside ? dmax(mmax, 1E-10 * floorr) : 0.0);
      }
    }
  };

  // ---- DRIFTMU (src/ModRamDrift.f90:382-473), optionally with the SUMRC moment ----
  auto driftmu = [&](const bool mom, const bool cfl) {
    const int ntask = NE * PG * cfg.nsegM;
    const int nround = (ntask + T - 1) / T;
    for (int rd = 0; rd < nround; ++rd) {
      const int e = tid + rd * T;
      const int pp = e % PG, kq = e / PG;
      const int k = kq % NE, seg = kq / NE;
      const int p = p0 + pp;
      const bool act = (e < ntask) && (p < P) && (p % NR != 0);
      const int la = 2 + seg * cfg.segM, lb = min(NPA - 1, la + cfg.segM - 1);
      const bool lastseg = (lb == NPA - 1);
      double* col = sT + k * PG + pp;                   // F(L) at col[(L-1)*RS]
      double Fm2 = 0, Fm1 = 0, F0 = 0, Fp1 = 0, Fp2 = 0, hi1 = 0, hi2 = 0;
      if (act) {
#define GETFL(L) (((L) > NPA) ? 0.0 : col[(size_t)(max((L), 2) - 1) * RS])   /* F(1) = F(2)  (:414) */
        Fm2 = GETFL(la - 2); Fm1 = GETFL(la - 1); F0 = GETFL(la); Fp1 = GETFL(la + 1); Fp2 = GETFL(la + 2);
        hi1 = GETFL(lb + 1); hi2 = GETFL(lb + 2);
#undef GETFL
      }
      if (cfg.nsegM > 1) __syncthreads();
      if (act) {
        const bool inside = !d.outp[p];
        const double wM = sWM[k];
        const double* ca = sMa + (la - 1) * PG + pp;    // coefficient pieces of L = la
        const double* cb = sMb + (la - 1) * PG + pp;
        double FBprev = 0.0, mmax = 0.0, macc = 0.0, fnew = 0.0;
        if (la > 2) {                                   // flux through the segment's lower edge
          const double c = fma(wM, cb[-PG], ca[-PG]);
          FBprev = c * limited_flux_fast(Fm2, Fm1, F0, Fp1, c < 0.0, fabs(c) * sRD[la - 2], beta);
        }
        double* pO = col + (size_t)(la - 1) * RS;
        int L = la;
        auto step = [&](const double nn) {
          const double c = fma(wM, *cb, *ca);
          ca += PG; cb += PG;
          const double ac = fabs(c) * sRD[L - 1];
          if (cfl) mmax = dmax(mmax, ac);
          double FB;
          if (L <= NPA - 2) FB = c * limited_flux_fast(Fm1, F0, Fp1, Fp2, c < 0.0, ac, beta);
          else FB = c * Fp1;                            // FBND(NPA-1) = F(NPA)  (:458)
          fnew = fma(-(FB - FBprev), sRW[L - 1], F0);
          FBprev = FB;
          if (fnew < 0.0) fnew = 1E-15;
          *pO = fnew;
          pO += RS;
          if (mom) macc = fma(fnew, sWMU[L - 1], macc);
          Fm1 = F0; F0 = Fp1; Fp1 = Fp2; Fp2 = nn;
          ++L;
        };
        // the value entering the window after cell L is F(L+3)
        for (; L + 3 <= lb;) step(pO[3 * (size_t)RS]);
        if (L + 2 <= lb) step(hi1);                     // L = lb-2: F(lb+1)
        if (L + 1 <= lb) step(hi2);                     // L = lb-1: F(lb+2)
        if (L <= lb) step(0.0);
        if (!inside) mmax = 0.0;
        if (lastseg) {
          const double c = fma(wM, *cb, *ca);           // CDriftMu(..,NPA)
          if (cfl && inside) mmax = dmax(mmax, fabs(c) * sRD[NPA - 1]);
          const double fN = fnew * d.FNHSc[(size_t)(NPA - 1) * Pp + p] * d.MU[NPA - 1] / d.FNHSc[(size_t)(NPA - 2) * Pp + p] / d.MU[NPA - 2];
          *pO = fN;                                     // :466
          if (mom) macc = fma(fN, sWMU[NPA - 1], macc);
        }
        if (mom && k >= 1 && p < (NT - 1) * NR) acc[0] += macc * (sWE[k] * sEK[k]);
        if (cfl) mmaxM = dmax(mmaxM, mmax);
      }
    }
  };

  // ---- the loss block (k_loss_mid): pointwise, four SUMRC moments ------------------
  auto losses = [&]() {
    const bool ion = (sp.kind != 3);
    const bool doA = (cfg.doA >> (s0 + blockIdx.y)) & 1;
    const int rowItems = NE * PG;
    int l = tid / rowItems, r = tid - l * rowItems;
    const int dl = T / rowItems, dr = T - dl * rowItems;
    const int pmom = (NT - 1) * NR;
    while (l < NPA) {
      const int k = r / PG, pp = r - k * PG;
      const int p = p0 + pp;
      const int i = p % NR;
      if (k >= 1 && p < P && i >= 1) {
        double* c = sT + (size_t)l * RS + r;
        double f = *c;
        const double w = sWE[k], wm = sWMU[l], e = sEK[k];
        const bool mom = (l >= 1) && (p < pmom);
        const bool useA = doA && (l >= 1);
        double facA = 1.0;
        if (useA) {
          if (ion) facA = exp(-(sSV[k] * sH[l * PG + pp] * d.DTs));
          else facA = sW[r];
          f = f * facA;
        }
        if (mom) acc[1] += e * (f * w * wm);
        if (l + 1 >= d.UPA[i]) {
          const double a = exp(sX[r] * sRF[l * PG + pp]);
          f = f * a;
          if (mom) acc[2] += e * (f * w * wm);
          f = f * a;
          if (mom) acc[3] += e * (f * w * wm);
        } else if (mom) {
          const double term = e * (f * w * wm);
          acc[2] += term;
          acc[3] += term;
        }
        if (useA) f = f * facA;
        if (mom) acc[4] += e * (f * w * wm);
        *c = f;
      }
      r += dr; l += dl;
      if (r >= rowItems) { r -= rowItems; ++l; }
    }
  };

  drifte(false);
  __syncthreads();
  driftmu(true, false);
  __syncthreads();
  losses();
  __syncthreads();
  driftmu(false, true);
  __syncthreads();
  drifte(true);
  __syncthreads();

  // ---- write the block back, reductions -------------------------------------------
  {
    constexpr int H = PG / 2;
    const int rowItems = NE * H;
    int l = tid / rowItems, r = tid - l * rowItems;
    const int dl = T / rowItems, dr = T - dl * rowItems;
    while (l < NPA) {
      const int k = r / H, hh = r - k * H;
      *(double2*)(sp.F + ((size_t)l * NE + k) * Pp + p0 + 2 * hh) = *(const double2*)(sT + (size_t)l * RS + k * PG + 2 * hh);
      r += dr; l += dl;
      if (r >= rowItems) { r -= rowItems; ++l; }
    }
  }
  double dtE = 1.0e300, dtM = 1.0e300;
  if (mmaxE > 0.0) dtE = sp.aRP / mmaxE;
  if (mmaxM > 0.0) dtM = sp.aRP / mmaxM;
  warp_min_to(sp.dtw + 2, dtE);
  warp_min_to(sp.dtw + 3, dtM);
  block_sum_to<5>(sp.part, blockIdx.x, acc, sRed);
}
