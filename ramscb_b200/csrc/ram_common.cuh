// Shared declarations of the RAM device library (sm_100a).
//
// Device data layout:
//   F2dev[S][l][k][Pp]      species-major; the (MLT, R) plane (p = j*NR + i, R
//                           fastest) is contiguous and padded to Pp (multiple of
//                           16 doubles = 128 B); K (energy) next, L (pitch angle)
//                           slowest.  All indices 0-based: I = i+1 etc.
//   2-D fields  [j][i]      row length NR+1 (a raw copy of the Fortran array)
//   3-D fields  [l][j][i]   row length NR+1 (raw copy)
//   "plane" arrays [p] and [l][Pp]: coefficient pieces precomputed by the prep
//                           kernels in the same p indexing as F2dev.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define RSG_MAX_SPECIES 8
#define RSG_NMOM 4   // moments produced by the fused loss block

// everything a sweep kernel needs that is shared by all species
struct RamDev {
  int nS, NR, NT, NE, NPA, NR1, P, Pp;
  double MDR, DPHI, CONF1, CONF2, BetaLim, FracCFL, DTs;
  // 1-D grids
  const double *RLZ, *EKEV, *WE, *DE, *MU, *WMU, *DMU;
  const int* UPA;  // [NR] 1-based L index of the loss-cone edge
  // raw fields
  const double *BNES, *dBdt, *VT, *EIR, *EIP;                      // (NR+1,NT)
  const double *FNHS, *FNIS, *BOUNHS, *BOUNIS, *HDNS, *dIdt, *dIbndt;  // (NR+1,NT,NPA)
  const int* outside;                                              // (NR,NT)
  // prep, 2-D planes [Pp]
  double *CR, *sB, *pT1, *pT3, *sBp, *DRD1, *DPD1, *BNESc, *dBdt2, *RLZp;
  unsigned char* outp;  // outsideMGNP per plane point
  // prep, 3-D planes [NPA][Pp]
  double *t1, *G, *sFp, *Gr, *Gp, *DRD2, *DPD2, *dBdt1, *dIdt1, *FNHSc;
  double *CMUDOT, *Gmr, *Gmp, *DRM2, *DPM2, *dIbndt2, *BOUNHSc, *HDNSc;
  // FAST-mode separable coefficient planes: CDrift* = a + w(K)*b
  //   R: CR[p] + P4[k]*fRb     P: fPa[p] - w2[k]*fPb
  //   E: uE[k]*fEa + vE[k]*fEb MU: fMa + wM[k]*fMb          (3-D ones are [NPA][Pp])
  double *fRb, *fPa, *fPb, *fEa, *fEb, *fMa, *fMb;
  double *CRt, *fRbt;          // CR, fRb transposed to [i][j] / [l][i][j]: coalesced for the radial walks (lane = MLT line)
  const double *rDMU, *rWMU;   // 1/DMU(L), 1/WMU(L)
  double* rFNHS;               // FAST: 1/FNHS plane [NPA][Pp]
  const double* exp2tab;       // FAST: 2^(j/64), j < 64 (table of fast_exp)
  const double *wPE, *wPA;     // FAST ANISCH pitch-angle weights [NPA]: WMU/MU*(1-MU^2), WMU*MU (MU(1):=MU(2) as FFACTOR(..,1)=FFACTOR(..,2))
};

// per-species device tables (pointers into one buffer) + scalars
struct SpecDev {
  int S;            // 0-based species
  int kind;
  double QS;
  double* F;        // species block of F2dev (current buffer)
  double* Fo;       // the other ping-pong buffer (sweeps read F, write Fo)
  const int* last;  // DRIFTR: index of the most recent inflow line (reference loop order)
  double* ghost;    // DRIFTR: F(NR+1), F(NR+2) of every line [(k*NPA+l)*NT+j][2]
  double* part;     // per-CTA partial sums for the moment reductions
  const double* FGEOS;  // [l][k][j]
  const double *P4, *eK, *epK, *aE, *sv;  // [NE]
  const double *P2, *EDOT, *ATLOS;        // [NE][NR]
  const double* aMU;                      // [NPA]
  const double *w2, *wM;                  // FAST-mode energy tables [NE]
  const double* tabE;                     // FAST DRIFTE: {uE, vE, 1/DE, 1/WE} per K, 32-byte records
  const double* FF;                       // FFACTOR [l][k][i]
  const double* rFFA;                     // FAST ANISCH: MU(2)/FFACTOR(S,I,K,2) = 1/(LZ^2*GREL/sqrt(GREL^2-1)) [k][i]
  const double* xATL;                     // FAST: log(ATLOS(S,I,K)) = -DTs*V/(2*RLZ) [k][i]
  const double* EPP;                      // [NE]
  const double* wfac;                     // WAVELO factor exp(-DTs/TAU_LIF) [NE][Pp]
  const double *DA, *DB;                  // WPADIF coefficient pair [l][k][Pp]
  const double *CA, *CB;                  // fused COULMU: elimination factors (cA,cB) pairs and RL, like the fused WPADIF's (k_coulmu_tables)
  const double* cK;                       // fused COULEN: COULE + COULI per energy [NE] (the same for every L < NPA, src/ModRamCoul.f90:98-101)
  double cg1, cg0;                        // COULEN ghost-cell ratios sqrt((GREL1^2-1)/(GREL2^2-1)), sqrt((GRZERO^2-1)/(GREL1^2-1)) (:182-183)
  double *tE, *tA;                        // ANISCH scratch [NE][Pp]
  double *aE2, *aA2;                      // fused step: ANISCH rows written by k_plane_rp<REV> [NPA * energy chunks][Pp]
  double *pper, *ppar;                    // ANISCH results [Pp]
  double GREL1, GREL2, sqrtA, GRZERO, sqrtB;  // DRIFTE ghost cells
  double aRP;                             // FracCFL*DTs
  double OMEt;                            // OME*DTs/DPHI
  unsigned long long* dt;                 // result block: [4] CFL minima (ordered bit patterns), moments, counters
  unsigned long long* dtw;                // where the sweeps of this launch accumulate their CFL minima [4]
};

// Sub-range of the (L,K) planes a launch works on (multi-GPU slabs): planes
// (l0 + a, k0 + b), a < nl, b < nk.  A single GPU uses {0, NPA, 0, NE}.
struct PlaneRange {
  int l0, nl, k0, nk;
};

// all species of one launch (blockIdx.y selects the species)
struct SpecPack {
  SpecDev s[RSG_MAX_SPECIES];
};
