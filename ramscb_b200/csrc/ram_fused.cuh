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
#include <type_traits>
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
// k_cfl_fast: the four CFL limits DtDriftR/P/E/Mu (src/ModRamDrift.f90:147, :240,
// :344, :435).  They are functions of the drift coefficients only, not of F2, so
// the fused path evaluates them here -- once per coefficient set (DRIFTPARA /
// field / mode change), cached by the host -- instead of inside every sweep.
// Same expressions and therefore the same bits as the sweeps' own tracking.
// grid: x = l, y = energy chunk of KCH, z = species
// =============================================================================
__global__ void __launch_bounds__(256) k_cfl_fast(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                  unsigned long long* __restrict__ cfl_all, int KCH) {
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  unsigned long long* out = cfl_all + 4 * (size_t)(s0 + blockIdx.z);
  const int NR = d.NR, NE = d.NE, P = d.P, Pp = d.Pp;
  const int l = blockIdx.x;
  const int ka = blockIdx.y * KCH, kb = min(NE, ka + KCH);
  double mR = 0.0, mP = 0.0, mE = 0.0, mM = 0.0;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    if (d.outp[p]) continue;
    const int j = p / NR, i = p - j * NR;
    const size_t o = (size_t)l * Pp + p;
    const double CRp = d.CR[p], gR = d.fRb[o], pa = d.fPa[p], pb = d.fPb[o];
    const double fA = d.fEa[o], fB = d.fEb[o], mA = d.fMa[o], mB = d.fMb[o];
    const double rdmu = d.rDMU[l];
    for (int k = ka; k < kb; ++k) {
      mR = dmax(mR, fabs(fma(sp.P4[k], gR, CRp)));
      if (i >= 1) {
        if (j >= 1) mP = dmax(mP, fabs(fma(-sp.w2[k], pb, pa)));
        const double* tb = sp.tabE + 4 * k;
        mE = dmax(mE, fabs(fma(tb[1], fB, tb[0] * fA)) * tb[2]);
        if (l >= 1) mM = dmax(mM, fabs(fma(sp.wM[k], mB, mA)) * rdmu);
      }
    }
    if (i >= 1 && ka == 0) mE = dmax(mE, 1E-10 * sp.tabE[2]);   // the max(|c|,1e-10) floor of :344 (1/DE largest at K=1)
  }
  warp_min_to(out + 0, sp.aRP / dmax(mR, 1E-10));
  warp_min_to(out + 1, sp.aRP / dmax(mP, 1E-10));
  warp_min_to(out + 2, mE > 0.0 ? sp.aRP / mE : 1.0e300);
  warp_min_to(out + 3, mM > 0.0 ? sp.aRP / mM : 1.0e300);
}

// CR, fRb in [i][j] order for the radial walks of k_plane_rp (grid: x = tiles of the plane, y = l)
__global__ void k_transpose_rcoef(RamDev d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;   // t = i*NT + j
  if (t >= d.P) return;
  const int i = t / d.NT, j = t - i * d.NT;
  const int l = blockIdx.y;
  if (l == 0) d.CRt[t] = d.CR[j * d.NR + i];
  d.fRbt[(size_t)l * d.Pp + t] = d.fRb[(size_t)l * d.Pp + j * d.NR + i];
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

// ---- TMA bulk copies (cp.async.bulk + mbarrier): one instruction moves a whole row of a plane between global and
// shared memory; the copy engine does the addressing, the threads only wait on the barrier ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_fence() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// =============================================================================
// k_plane_rp: DRIFTR (src/ModRamDrift.f90:95-198) and DRIFTP (:204-279) of KC
// consecutive (K,L) planes, back to back on a shared-memory copy; REV = reverse half
// step (DRIFTP first, then DRIFTR with the fused SUMRC of src/ModRamRun.f90:174).
// Planes are stored with an odd row stride NRp so that the radial walks of
// consecutive MLT lines hit different banks.  Both sweeps walk line segments in
// place with a rolling register window (one thread per segment); the foreign halo
// cells of a segment are read before the barrier that precedes the walk.
// grid: x = energy chunk, y = l, z = species
// =============================================================================
// Where the ranks that SHARE a species hold their copies of F2 (multi-GPU, ram_shard.inl): the plane kernels
// own pitch-angle slabs [lcut[g], lcut[g+1]), the column kernel ranges of plane positions [ccut[g], ccut[g+1])
// (multiples of the column block).  F[g] is rank g's F2 buffer mapped into this process (CUDA IPC peer memory over
// NVLink; F[gidx] is the local one), all with the same layout, so the re-sharding of SURVEY 8(e) is the write-back of
// the producing kernel: every finished value is stored once, straight into the buffer of the rank that reads it next.
#define RSG_MAX_PEERS 8
struct PeerView {
  int G, gidx;
  double* F[RSG_MAX_PEERS];
  int lcut[RSG_MAX_PEERS + 1];
  int ccut[RSG_MAX_PEERS + 1];
};
__device__ __forceinline__ int peer_owner(const int* cut, int G, int x) {
  int g = 0;
  while (g + 1 < G && x >= cut[g + 1]) ++g;
  return g;
}

struct PlaneCfg {
  int KC;             // planes (energies) per CTA
  int NRp, PS;        // padded row stride and plane stride of the shared copy (doubles)
  int nsegR, segR;    // DRIFTR: segments per line, cells per segment (cells I=2..NR)
  int nsegP, segP;    // DRIFTP (cells J=2..NT)
  int part_off;       // REV: offset of this kernel's SUMRC partials in SpecDev::part
  int l0;             // first pitch angle of the launch (slab-sharded ranks); blockIdx.y counts from it
  int anisch;         // REV: also write this CTA's share of the ANISCH pitch-angle / energy sums (SpecDev::aE2 / aA2)
  int tma;            // stage rows with TMA bulk copies when the layout allows (NR even)
};

template <bool REV, bool PEER = false>
__global__ void __launch_bounds__(512) k_plane_rp(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                  PlaneCfg cfg, const __grid_constant__ PeerView pv) {
  extern __shared__ double smem[];
  const SpecDev& sp = pk.s[s0 + blockIdx.z];
  const int NR = d.NR, NT = d.NT, NE = d.NE, P = d.P, Pp = d.Pp;
  const int T = blockDim.x, tid = threadIdx.x;
  const int NRp = cfg.NRp, PS = cfg.PS;
  const int l = cfg.l0 + blockIdx.y;
  const int k0 = blockIdx.x * cfg.KC, KCa = min(NE, k0 + cfg.KC) - k0;
  double* sP = smem;                                   // [KC][PS]
  double* sG = smem + (size_t)cfg.KC * PS;             // [KC][NT][2] ghost cells F(NR+1), F(NR+2)
  int* sIn = (int*)(sG + 2 * (size_t)cfg.KC * NT);     // [KC][NT] inflow flag
  double* sRed = (double*)(sIn + ((cfg.KC * NT + 1) & ~1));   // [32]
  const double beta = d.BetaLim;
  double* Fg = sp.F + ((size_t)l * NE + k0) * Pp;

  // ---- stage the planes and the line state.  A plane is contiguous in global memory; element
  // e = q*P + j*NR + i goes to q*PS + j*NRp + i.  16-byte chunks when NR is even (a chunk never
  // straddles a row; NRp is even then), else 8-byte; the offsets advance incrementally.
  const int E = (NRp & 1) ? 1 : 2;
  auto for_chunks = [&](auto&& body) {
    const int e0 = tid * E;
    int q = e0 / P, p = e0 - q * P;
    int j = p / NR, i = p - j * NR;
    const int stp = T * E, dj = stp / NR, di = stp - dj * NR;
    int so = q * PS + j * NRp + i, go = q * Pp + j * NR + i;
    const int dso = dj * NRp + di, dgo = dj * NR + di;
    while (q < KCa) {
      body(so, go, p);
      i += di; j += dj; so += dso; go += dgo; p += dgo;
      if (i >= NR) { i -= NR; ++j; so += NRp - NR; }
      while (j >= NT) { j -= NT; ++q; so += PS - NT * NRp; go += Pp - P; p -= P; }
    }
  };
  // NR even: every row of a plane is a 16-byte aligned run of NR*8 bytes on both sides -- one TMA bulk copy per row
  // (cp.async.bulk, completion counted on an mbarrier); else 8-byte cp.async chunks.
  __shared__ unsigned long long mbar;
  const bool tma = (E == 2) && cfg.tma;
  const int nrows = KCa * NT;
  {
    if (tma) {
      if (tid == 0) mbar_init(&mbar, 1);
      __syncthreads();
      if (tid == 0) mbar_expect_tx(&mbar, (unsigned)(nrows * NR * 8));
      for (int t = tid; t < nrows; t += T) {
        const int q = t / NT, j = t - q * NT;
        tma_load(sP + (size_t)q * PS + j * NRp, Fg + (size_t)q * Pp + j * NR, (unsigned)(NR * 8), &mbar);
      }
    } else if (E == 2) for_chunks([&](int so, int go, int) { cp_async16(sP + so, Fg + go); });
    else for_chunks([&](int so, int go, int) { cp_async8(sP + so, Fg + go); });
    asm volatile("cp.async.commit_group;");
    for (int t = tid; t < KCa * NT; t += T) {
      const int q2 = t / NT, j2 = t - q2 * NT;
      const int line = ((k0 + q2) * d.NPA + l) * NT + j2;
      sIn[t] = (sp.last[line] == line);
      sG[2 * t] = sp.ghost[2 * (size_t)line];
      sG[2 * t + 1] = sp.ghost[2 * (size_t)line + 1];
    }
    asm volatile("cp.async.wait_group 0;");
    if (tma) mbar_wait(&mbar, 0);
  }
  __syncthreads();
  double macc = 0.0;

  // ---- DRIFTR: cells I=2..NR of every line (j, plane q) ------------------------------
  auto driftr = [&]() {
    const int ntask = NT * KCa * cfg.nsegR;
    const int nround = (ntask + T - 1) / T;
    for (int rd = 0; rd < nround; ++rd) {
      const int e = tid + rd * T;
      const int j = e % NT, qq = e / NT;
      const int q = qq % KCa, seg = qq / KCa;
      const bool act = e < ntask;
      const int ia = 2 + seg * cfg.segR, ib = min(NR, ia + cfg.segR - 1);
      double* row = sP + (size_t)q * PS + j * NRp;     // F(I) at row[I-1]
      double Fm1 = 0, F0 = 0, Fp1 = 0, Fp2 = 0, hi1 = 0, hi2 = 0, g1 = 0, g2 = 0;
      bool inflow = false;
      if (act) {
        g1 = sG[2 * (q * NT + j)]; g2 = sG[2 * (q * NT + j) + 1];
        inflow = sIn[q * NT + j] != 0;
#define GETR(I) (((I) < 1) ? 0.0 : (((I) <= NR) ? row[(I)-1] : (((I) == NR + 1) ? g1 : g2)))
        Fm1 = GETR(ia - 2); F0 = GETR(ia - 1); Fp1 = GETR(ia); Fp2 = GETR(ia + 1);
        hi1 = GETR(ib + 1); hi2 = GETR(ib + 2);
#undef GETR
      }
      if (cfg.nsegR > 1) __syncthreads();
      if (act) {
        const int k = k0 + q;
        const double P4k = sp.P4[k];
        const double* cr = d.CRt + (ia - 2) * NT + j;                    // coefficient pieces of I = ia-1, [i][j] order
        const double* gr = d.fRbt + (size_t)l * Pp + (ia - 2) * NT + j;
        double dm1 = F0 - Fm1, d0 = Fp1 - F0, dp1 = Fp2 - Fp1;
        double phiPrev;
        {                                               // interface ia-1: flux only
          const double c = fma(P4k, *gr, *cr);
          double FB = limited_flux_d(F0, Fp1, dm1, d0, dp1, c < 0.0, fabs(c), beta);
          if (ia == 2) FB = inflow ? Fp1 : 0.0;         // FBND(1) = F(2) | 0   (:155,:159)
          phiPrev = c * FB;
        }
        double cnext = fma(P4k, gr[NT], cr[NT]);        // CDriftR of I = ia
        cr += 2 * NT; gr += 2 * NT;
        double* pO = row + (ia - 1);
        const bool inmom = REV && l >= 1 && j <= NT - 2 && k >= 1;
        const double wk = inmom ? d.WE[k] * d.EKEV[k] : 0.0;
        auto step = [&](const double nn, const bool lastI) {
          F0 = Fp1; Fp1 = Fp2; Fp2 = nn;
          dm1 = d0; d0 = dp1; dp1 = Fp2 - Fp1;
          const double c = cnext;
          if (!lastI) cnext = fma(P4k, *gr, *cr);       // next cell's coefficient, off the critical path
          cr += NT; gr += NT;
          double FB = limited_flux_d(F0, Fp1, dm1, d0, dp1, c < 0.0, fabs(c), beta);
          if (lastI && !inflow) FB = F0;                // FBND(NR) = F(NR)     (:156)
          const double phi = c * FB;
          double fn = F0 - phi + phiPrev;               // :186
          if (fn < 0.0) fn = 1E-15;
          *pO++ = fn;
          phiPrev = phi;
          if (REV) macc = fma(fn, wk, macc);
        };
        int I = ia;
#pragma unroll 4
        for (; I + 2 <= ib; ++I) step(pO[2], false);
        if (I + 1 <= ib) { step(hi1, false); ++I; }
        if (I <= ib) step(hi2, ib == NR);
      }
    }
  };

  // ---- DRIFTP: cells J=2..NT of every line (i >= 1, plane q); periodic in J -----------
  auto driftp = [&]() {
    const int ntask = NR * KCa * cfg.nsegP;
    const int nround = (ntask + T - 1) / T;
    for (int rd = 0; rd < nround; ++rd) {
      const int e = tid + rd * T;
      const int i = e % NR, qq = e / NR;
      const int q = qq % KCa, seg = qq / KCa;
      const bool act = (e < ntask) && (i >= 1);
      const int ja = 2 + seg * cfg.segP, jb = min(NT, ja + cfg.segP - 1);
      double* colp = sP + (size_t)q * PS + i;           // F(J) at colp[(J-1)*NRp]
      double Fm1 = 0, F0 = 0, Fp1 = 0, Fp2 = 0, hi1 = 0, hi2 = 0, F1row = 0;
      if (act) {
#define GETP(J) colp[(size_t)(((J) > NT ? (J) - NT + 1 : (J)) - 1) * NRp]   /* F(NT+1)=F(2), F(NT+2)=F(3) */
        const int Jpe = (ja == 2) ? NT : ja - 1;        // interface below the segment (:261-262)
        Fm1 = GETP(Jpe - 1); F0 = GETP(Jpe); Fp1 = GETP(Jpe + 1); Fp2 = GETP(Jpe + 2);
        hi1 = GETP(jb + 1); hi2 = GETP(jb + 2);
        F1row = colp[0];
#undef GETP
      }
      if (cfg.nsegP > 1) __syncthreads();
      if (act) {
        const int k = k0 + q;
        const double w2k = sp.w2[k];
        const int Jpe = (ja == 2) ? NT : ja - 1;
        const double* pa = d.fPa + (Jpe - 1) * NR + i;
        const double* pb = d.fPb + (size_t)l * Pp + (Jpe - 1) * NR + i;
        double dm1 = F0 - Fm1, d0 = Fp1 - F0, dp1 = Fp2 - Fp1;
        double prev;
        {
          const double c = fma(-w2k, *pb, *pa);
          // interface NT-1 (a segment that starts at J = NT): far-upwind difference F(2) - F(1), see limited_flux_num
          prev = c * limited_flux_d(F0, Fp1, dm1, d0, (Jpe == NT - 1) ? Fp2 - F1row : dp1, c < 0.0, fabs(c), beta);
        }
        if (ja == 2) d0 = Fp1 - F1row;                  // cell J=2 sees the stored F(1), not F(NT)
        pa = d.fPa + (ja - 1) * NR + i;
        pb = d.fPb + (size_t)l * Pp + (ja - 1) * NR + i;
        double cnext = fma(-w2k, *pb, *pa);
        double* pO = colp + (size_t)(ja - 1) * NRp;
        double fnew = 0.0;
        // wrapfix: this step's interface is J = NT-1, whose far-upwind difference is F(2) - F(1) (limited_flux_num)
        auto step = [&](const double nn, const bool more, const bool wrapfix) {
          F0 = Fp1; Fp1 = Fp2; Fp2 = nn;
          dm1 = d0; d0 = dp1; dp1 = Fp2 - Fp1;
          const double c = cnext;
          pa += NR; pb += NR;
          if (more) cnext = fma(-w2k, *pb, *pa);
          // (colp[0] = the stored F(1): only this thread rewrites it, after its last step; re-read so that it is not live in the loop)
          const double cur = c * limited_flux_d(F0, Fp1, dm1, d0, wrapfix ? Fp2 - colp[0] : dp1, c < 0.0, fabs(c), beta);
          fnew = F0 - cur + prev;                       // :266
          if (fnew < 0.0) fnew = 1E-15;
          *pO = fnew;
          pO += NRp;
          prev = cur;
        };
        int J = ja;
#pragma unroll 4
        for (; J + 2 <= jb; ++J) step(pO[2 * (size_t)NRp], true, false);     // J <= NT-2 here
        if (J + 1 <= jb) { step(hi1, true, J == NT - 1); ++J; }
        if (J <= jb) step(hi2, false, J == NT - 1);
        if (jb == NT) colp[0] = fnew;                   // F2(J=1) = F2(J=NT)  (:272)
      }
    }
  };

  if (REV) { driftp(); __syncthreads(); driftr(); }
  else { driftr(); __syncthreads(); driftp(); }
  __syncthreads();

  // ---- write the planes back (the never-advanced I=1 cells are rewritten with their own value).
  // REV: the step ends here, so the epilogue of ram_run (src/ModRamRun.f90:186-201) rides along:
  // F2(J=NT) = F2(J=1), then F2 = 1e-31 outside the magnetopause.
  if (REV) {
    for (int t = tid; t < KCa * NR; t += T) {
      const int q = t / NR, i = t - q * NR;
      sP[(size_t)q * PS + (NT - 1) * NRp + i] = sP[(size_t)q * PS + i];
    }
    __syncthreads();
    // ---- ANISCH (src/ModRamRun.f90:360-400) rides along too: the step's final F2 of this pitch angle and these
    // energies is in shared memory, so the CTA writes its share of  sum_K EPP(K)/A(S,I,K) * sum_L F2 * w(L)/FNHS  for every
    // plane position (one row of [slab pitch angles x energy chunks][Pp] per CTA; k_finalize adds the rows in a fixed
    // order).  L = 1 carries the value of L = 2 (the side effect F2(L=1) = F2(L=2), :366, is applied by k_finalize).
    if (cfg.anisch) {
      const size_t row = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      double* oE = sp.aE2 + row * Pp;
      double* oA = sp.aA2 + row * Pp;
      const double wE = (l >= 1) ? d.wPE[l] : 0.0, wA = (l >= 1) ? d.wPA[l] : 0.0;
      const double wE1 = (l == 1) ? d.wPE[0] : 0.0, wA1 = (l == 1) ? d.wPA[0] : 0.0;
      for (int p = tid; p < Pp; p += T) {
        double se = 0.0, sa = 0.0;
        if (p < P) {
          const int j = p / NR, i = p - j * NR;
          if (i >= 1 && l + 1 <= d.UPA[i] - 1) {              // Fortran L = l+1 inside the loss-cone edge
            double acc = 0.0;
            const bool outm = d.outp[p] != 0;
            for (int q = 0; q < KCa; ++q) {
              const int k = k0 + q;
              if (k < 1) continue;
              const double f = outm ? 1.e-31 : sP[(size_t)q * PS + j * NRp + i];
              acc = fma(f, sp.EPP[k] * sp.rFFA[k * NR + i], acc);
            }
            const double r0 = d.rFNHS[(size_t)l * Pp + p];
            se = acc * (r0 * wE);
            sa = acc * (r0 * wA);
            if (l == 1) {                                      // the L = 1 term, with F2(L=1) := F2(L=2)
              const double r1 = d.rFNHS[p];
              se = fma(acc, r1 * wE1, se);
              sa = fma(acc, r1 * wA1, sa);
            }
          }
        }
        oE[p] = se;
        oA[p] = sa;
      }
    }
    if (tma) {
      // the magnetopause mask on the shared copy, then one bulk store per row
      for (int p = tid; p < P; p += T)
        if (d.outp[p]) {
          const int j = p / NR, i = p - j * NR;
          for (int q = 0; q < KCa; ++q) sP[(size_t)q * PS + j * NRp + i] = 1.e-31;
        }
      tma_store_fence();
      __syncthreads();
      for (int t = tid; t < nrows; t += T) {
        const int q = t / NT, j = t - q * NT;
        tma_store(Fg + (size_t)q * Pp + j * NR, sP + (size_t)q * PS + j * NRp, (unsigned)(NR * 8));
      }
      tma_store_commit_wait();
    } else if (E == 2)
      for_chunks([&](int so, int go, int pl) {
        double2 v = *(const double2*)(sP + so);
        const unsigned short o2 = *(const unsigned short*)(d.outp + pl);
        if (o2 & 0xff) v.x = 1.e-31;
        if (o2 >> 8) v.y = 1.e-31;
        *(double2*)(Fg + go) = v;
      });
    else
      for_chunks([&](int so, int go, int pl) { Fg[go] = d.outp[pl] ? 1.e-31 : sP[so]; });
  } else if (PEER) {
    // the column kernel runs next, on the rank that owns the plane position: store there (NVLink peer memory)
    const ptrdiff_t rel = Fg - pv.F[pv.gidx];
    if (E == 2)
      for_chunks([&](int so, int go, int pl) {
        *(double2*)(pv.F[peer_owner(pv.ccut, pv.G, pl)] + rel + go) = *(const double2*)(sP + so);
      });
    else for_chunks([&](int so, int go, int pl) { pv.F[peer_owner(pv.ccut, pv.G, pl)][rel + go] = sP[so]; });
  } else if (tma) {
    tma_store_fence();
    __syncthreads();
    for (int t = tid; t < nrows; t += T) {
      const int q = t / NT, j = t - q * NT;
      tma_store(Fg + (size_t)q * Pp + j * NR, sP + (size_t)q * PS + j * NRp, (unsigned)(NR * 8));
    }
    tma_store_commit_wait();
  } else {
    if (E == 2) for_chunks([&](int so, int go, int) { *(double2*)(Fg + go) = *(const double2*)(sP + so); });
    else for_chunks([&](int so, int go, int) { Fg[go] = sP[so]; });
  }
  if (REV) {
    double acc[1] = {macc * d.WMU[l]};
    block_sum_to<1>(sp.part + cfg.part_off, (size_t)blockIdx.y * gridDim.x + blockIdx.x, acc, sRed);
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
  int nsegL, segL;    // loss block: pitch-angle segments per (k, position) column
  int doA;            // bit s: species s applies its first/last loss operator
  int b0;             // first block of plane positions of the launch (column-sharded ranks)
  int pb0;            // partial-sum slot of the launch's first block (chunked launches of one rank's block range)
  int doW;            // bit s: species s runs WPADIF after the first / before the second DRIFTMU (WPI instantiation only)
  int wpart_off;      // where the two WPADIF moments of a block go in sp.part
  int doC;            // Coulomb operators (COULEN, COULMU | COULMU, COULEN around the loss block) for every species (WPI instantiation only)
  int tpos;           // TimeRamElapsed > 0: COULMU clamps negatives (src/ModRamCoul.f90:289)
  int cpart_off;      // where the four Coulomb moments of a block go in sp.part
  const double* NECR; // plasmaspheric density [j][i] (NR,NT)
};

// =============================================================================
// k_wpadif_tables: WPADIF (src/ModRamWPI.f90:643-714) is a tridiagonal solve per (K, position)
// line whose matrix depends on the diffusion coefficients, the field factors and DTs only -- not
// on F2.  The elimination factors of the Thomas recurrences are therefore tabulated once per
// (coefficient set, fields, DTs) and the fused column kernel applies them with two FMAs per cell:
//   forward   RK(L) = F2(L) * cA(L) + RK(L-1) * cB(L),   cA = 1/(FACMU*DENOM), cB = GN/DENOM
//   closing   f(NPA-1) = RK(NPA-1) * cA(NPA)             cA(NPA) := 1/(1 + RL(NPA-1))
//   backward  f(L) = RK(L) - RL(L) * f(L+1),  F2(L) = f(L) * FACMU(L)
// with AN, GN, DENOM, RL exactly as in the reference (same operation order as k_wpadif).
// AB holds (cA, cB) pairs, RL the third factor, both [l][k][Pp] like F2.  The count of rows that
// are not diagonally dominant (the reference's warning, :689) is added to *viol.
// One thread per line; grid: x = lines of nk*Pp / T
// =============================================================================
__global__ void __launch_bounds__(128) k_wpadif_tables(const __grid_constant__ RamDev d, const double* __restrict__ DA,
                                                       const double* __restrict__ DB, double2* __restrict__ AB,
                                                       double* __restrict__ RLt, unsigned long long* __restrict__ viol_out) {
  const int NE = d.NE, NPA = d.NPA, Pp = d.Pp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)NE * Pp) return;
  const int k = (int)(t / Pp), p = (int)(t - (long long)k * Pp);
  const int i = p % d.NR;
  if (p >= d.P || i < 1 || k < 1) return;
  const size_t LS = (size_t)NE * Pp;
  const size_t o = (size_t)k * Pp + p;
  const double DTs = d.DTs;
  unsigned long long viol = 0;
  double rlm = -1.;
  double Dm = DA[o] + DB[o];   // D(L-1) for L=2
  for (int L = 2; L <= NPA - 1; ++L) {
    const double FACMU = d.FNHSc[(size_t)(L - 1) * Pp + p] * d.MU[L - 1];
    const double Dl = DA[o + (size_t)(L - 1) * LS] + DB[o + (size_t)(L - 1) * LS];
    double AN = Dl / d.DMU[L - 1];
    double GN = Dm / d.DMU[L - 2];
    AN = AN * DTs / FACMU / d.WMU[L - 1];
    GN = GN * DTs / FACMU / d.WMU[L - 1];
    const double BN = AN + GN;
    if (fabs(-1 - BN) < (fabs(AN) + fabs(GN))) ++viol;
    const double DENOM = BN + GN * rlm + 1;
    rlm = -AN / DENOM;
    AB[o + (size_t)(L - 1) * LS] = make_double2(1.0 / (FACMU * DENOM), GN / DENOM);
    RLt[o + (size_t)(L - 1) * LS] = rlm;
    Dm = Dl;
  }
  AB[o + (size_t)(NPA - 1) * LS] = make_double2(1.0 / (1 + rlm), 0.0);
  if (viol) atomicAdd(viol_out, viol);
}

// =============================================================================
// k_coulmu_tables: the same idea for COULMU (src/ModRamCoul.f90:229-296): its tridiagonal matrix depends on the rate
// tables ATA / GTA (COULPARA), the plasmaspheric density and the field factors only -- AN, GN, DENOM in the reference's
// operation order (k_coulmu), factors laid out like k_wpadif_tables' so the fused column kernel applies them with the
// same stage.  One thread per (K, position) line.
// =============================================================================
__global__ void __launch_bounds__(128) k_coulmu_tables(const __grid_constant__ RamDev d, const double* __restrict__ ATA,
                                                       const double* __restrict__ GTA, const double* __restrict__ NECR,
                                                       double2* __restrict__ AB, double* __restrict__ RLt) {
  const int NE = d.NE, NPA = d.NPA, Pp = d.Pp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)NE * Pp) return;
  const int k = (int)(t / Pp), p = (int)(t - (long long)k * Pp);
  const int j = p / d.NR, i = p - j * d.NR;
  if (p >= d.P || i < 1 || k < 1) return;
  const int I = i + 1, J = j + 1;
  const size_t LS = (size_t)NE * Pp;
  const size_t o = (size_t)k * Pp + p;
  const double XNE = NECR[p];
  double rlm = -1.;
  double BASm = XNE * R3(d.BOUNIS, I, J, 1) / 2. / R3(d.BOUNHS, I, J, 1);
  for (int L = 2; L <= NPA - 1; ++L) {
    const double BOUNHSl = R3(d.BOUNHS, I, J, L), FNHSl = R3(d.FNHS, I, J, L);
    const double BAS = XNE * R3(d.BOUNIS, I, J, L) / 2. / BOUNHSl;
    const double AN = ATA[(size_t)k * NPA + (L - 1)] * BAS / FNHSl * BOUNHSl;
    const double GN = GTA[(size_t)k * NPA + (L - 1)] * BASm / FNHSl * R3(d.BOUNHS, I, J, L - 1);
    const double BN = AN + GN;
    const double DENOM = BN + GN * rlm + 1;
    rlm = -AN / DENOM;
    AB[o + (size_t)(L - 1) * LS] = make_double2(1.0 / (FNHSl * d.MU[L - 1] * DENOM), GN / DENOM);
    RLt[o + (size_t)(L - 1) * LS] = rlm;
    BASm = BAS;
  }
  AB[o + (size_t)(NPA - 1) * LS] = make_double2(1.0 / (1 + rlm), 0.0);
}

// EXT: 0 = drift + loss stages only, 1 = + WPADIF, 2 = + WPADIF and the Coulomb operators (own instantiation: the extra
// stages cost registers the flags-5 step should not pay for)
template <int PG, int MAXT, int EXT, bool PEER = false>
__global__ void __launch_bounds__(MAXT) k_col_fused(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                    ColCfg cfg, const __grid_constant__ PeerView pv) {
  extern __shared__ double smem[];
  constexpr bool WPI = EXT >= 1, CO = EXT >= 2;
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  const int NR = d.NR, NT = d.NT, NE = d.NE, NPA = d.NPA, P = d.P, Pp = d.Pp;
  const int T = blockDim.x, tid = threadIdx.x;
  const int p0 = (cfg.b0 + blockIdx.x) * PG;
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
  double* sRD = sW + NE * PG;                // [NPA] each: 1/DMU, 1/WMU, WMU, WMU with L=1 zeroed
  double* sRW = sRD + NPA;
  double* sWMU = sRW + NPA;
  double* sWMZ = sWMU + NPA;
  double* sE2 = sWMZ + NPA;                  // [64] 2^(j/64)
  double* sRed = sE2 + 64;                   // [5][32]
  double* sFM = sRed + 5 * 32;               // WPI only: [NPA][PG] FNHS*MU (FACMU of WPADIF / COULMU)
  double* sXc = sFM + NPA * PG;              // WPI only: [NPA][PG] NECR*BANE(L) of COULEN (src/ModRamCoul.f90:170-176)
  double* sTabC = sXc + NPA * PG;            // WPI only: [NE][4] COULE+COULI, 0, 1/DE, 1/WE
  const bool doW = WPI && ((cfg.doW >> (s0 + blockIdx.y)) & 1);
  const bool doC = CO && cfg.doC;

  // ---- stage the block (asynchronous 16-byte copies, all in flight) and its tables -----
  {
    // item t = 16 bytes: (l,k) row t/H, half hh = t%H.  Offsets advance incrementally (32-bit: a
    // species block is < 2^31 doubles); the shared offset skips the pad of the energy stride at
    // every l wrap.
    constexpr int H = PG / 2;
    const int rowItems = NE * H;
    int l = tid / rowItems, r = tid - l * rowItems;
    const int dl = T / rowItems, dr = T - dl * rowItems;
    const int pad = (NEs - NE) * PG;
    int so = l * RS + (r / H) * PG + 2 * (r % H);
    int go = (l * NE + r / H) * Pp + p0 + 2 * (r % H);
    // one step of T items = dl whole l's + dr items; dr items = (dr/H) rows + (dr%H) halves
    const int so_step = dl * RS + (dr / H) * PG + 2 * (dr % H);
    const int go_step = (dl * NE + dr / H) * Pp + 2 * (dr % H);
    int hh = r % H;
    auto advance = [&]() {
      r += dr; l += dl; so += so_step; go += go_step; hh += dr % H;
      if (hh >= H) { hh -= H; so += PG - 2 * H; go += Pp - 2 * H; }    // carry of the half index into the row
      if (r >= rowItems) { r -= rowItems; ++l; so += pad; }
    };
    while (l < NPA) {
      cp_async16(sT + so, sp.F + go);
      advance();
    }
    asm volatile("cp.async.commit_group;");
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
    for (int t = tid; t < NPA; t += T) {
      sRD[t] = d.rDMU[t]; sRW[t] = d.rWMU[t]; sWMU[t] = d.WMU[t];
      sWMZ[t] = (t >= 1) ? d.WMU[t] : 0.0;
    }
    for (int t = tid; t < 64; t += T) sE2[t] = d.exp2tab[t];
    if (WPI && (doW || doC))
      for (int t = tid; t < NPA * PG; t += T) {
        const int l2 = t / PG, pp = t - l2 * PG;
        sFM[t] = d.FNHSc[(size_t)l2 * Pp + p0 + pp] * d.MU[l2];
      }
    if (CO && doC) {
      for (int t = tid; t < NPA * PG; t += T) {
        const int l2 = t / PG, pp = t - l2 * PG;
        const int p = min(p0 + pp, P - 1);
        const int j = p / NR, i = p - j * NR;
        // BANE(L) = (1 - FNIS/2/FNHS)/(1 - MU^2), L >= NPA-10 repeat L = NPA-11; L = 1 is not advanced by COULEN
        const int Lb = min(l2 + 1, NPA - 11);
        const double bane = (1. - R3(d.FNIS, i + 1, j + 1, Lb) / 2. / R3(d.FNHS, i + 1, j + 1, Lb)) / (1. - d.MU[Lb - 1] * d.MU[Lb - 1]);
        sXc[t] = (l2 == NPA - 1) ? 0.0 : cfg.NECR[p] * bane;   // COULE, COULI are never assigned at L = NPA (:98-101): c = 0 there
      }
      for (int t = tid; t < NE; t += T) {
        sTabC[4 * t] = sp.cK[t];
        sTabC[4 * t + 1] = 0.0;
        sTabC[4 * t + 2] = sp.tabE[4 * t + 2];
        sTabC[4 * t + 3] = sp.tabE[4 * t + 3];
      }
    }
    asm volatile("cp.async.wait_group 0;");
  }
  __syncthreads();

  const double beta = d.BetaLim;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // SUMRC after DRIFTMU, and the four of the loss block
  double accW[2] = {0.0, 0.0};                 // WPI: SUMRC after the first and after the second WPADIF
  double accC[4] = {0.0, 0.0, 0.0, 0.0};       // Coulomb: SUMRC after COULEN, COULMU | COULMU, COULEN (slots 10..13)

  // ---- WPADIF (src/ModRamWPI.f90:643-714) with the tabulated elimination factors of
  // k_wpadif_tables (sp.DA = (cA,cB) pairs, sp.DB = RL): a thread solves the lines of one
  // (energy, position); RK overwrites F2 on the way up, F2 is rebuilt on the way down -------
  auto thomas = [&](const double2* abBase, const double* rlBase, double& accum, const bool clamp) {
    const int ntask = NE * PG;
    for (int e = tid; e < ntask; e += T) {
      const int pp = e % PG, k = e / PG;
      const int p = p0 + pp;
      if (k < 1 || p >= P || p % NR < 1) continue;
      const size_t LS = (size_t)NE * Pp;
      const double2* ab = abBase + (size_t)k * Pp + p;
      const double* rl = rlBase + (size_t)k * Pp + p;
      double* col = sT + k * PG + pp;                   // F(L) at col[(L-1)*RS]
      const double* fm = sFM + pp;
      // The factors come from global memory (three F2-sized tables: L2 / HBM latency per access) and the
      // recurrences are serial in L: the loads of the next WB cells are issued before the current WB cells
      // are eliminated, so a line always has 2*WB table reads in flight instead of waiting on each one.
      constexpr int WB = 8;
      double rk = 0.0;
      {
        double2 nx[WB];
#pragma unroll
        for (int b = 0; b < WB; ++b) nx[b] = (1 + b <= NPA - 2) ? ab[(size_t)(1 + b) * LS] : make_double2(0.0, 0.0);
        for (int l0 = 1; l0 <= NPA - 2; l0 += WB) {
          double2 cu[WB];
#pragma unroll
          for (int b = 0; b < WB; ++b) cu[b] = nx[b];
#pragma unroll
          for (int b = 0; b < WB; ++b) if (l0 + WB + b <= NPA - 2) nx[b] = ab[(size_t)(l0 + WB + b) * LS];
#pragma unroll
          for (int b = 0; b < WB; ++b)
            if (l0 + b <= NPA - 2) {
              rk = fma(col[(size_t)(l0 + b) * RS], cu[b].x, rk * cu[b].y);
              col[(size_t)(l0 + b) * RS] = rk;
            }
        }
      }
      double f = rk * ab[(size_t)(NPA - 1) * LS].x;      // f(NPA-1) = RK/(1+RL)
      double fN = f * fm[(NPA - 1) * PG];               // F2(NPA) = f(NPA-1)*FACMU(NPA)
      if (clamp && fN < 0.0) fN = 1E-15;                // COULMU with TimeRamElapsed > 0 (src/ModRamCoul.f90:289)
      double macc = fN * sWMU[NPA - 1];
      col[(size_t)(NPA - 1) * RS] = fN;
      fN = f * fm[(NPA - 2) * PG];
      if (clamp && fN < 0.0) fN = 1E-15;
      macc = fma(fN, sWMU[NPA - 2], macc);
      col[(size_t)(NPA - 2) * RS] = fN;
      {
        double nx[WB];
#pragma unroll
        for (int b = 0; b < WB; ++b) nx[b] = (NPA - 3 - b >= 1) ? rl[(size_t)(NPA - 3 - b) * LS] : 0.0;
        for (int l0 = NPA - 3; l0 >= 1; l0 -= WB) {
          double cu[WB];
#pragma unroll
          for (int b = 0; b < WB; ++b) cu[b] = nx[b];
#pragma unroll
          for (int b = 0; b < WB; ++b) if (l0 - WB - b >= 1) nx[b] = rl[(size_t)(l0 - WB - b) * LS];
#pragma unroll
          for (int b = 0; b < WB; ++b)
            if (l0 - b >= 1) {
              const int l = l0 - b;
              f = fma(-cu[b], f, col[(size_t)l * RS]);
              fN = f * fm[l * PG];
              if (clamp && fN < 0.0) fN = 1E-15;
              macc = fma(fN, sWMU[l], macc);
              col[(size_t)l * RS] = fN;
            }
        }
      }
      {
        double f1 = f * fm[0];                          // RK(1) = 0, RL(1) = -1: f(1) = f(2)
        if (clamp && f1 < 0.0) f1 = 1E-15;
        col[0] = f1;
      }
      if (p < (NT - 1) * NR) accum += macc * (sWE[k] * sEK[k]);
    }
  };
  auto wpadif = [&](const int which) { thomas((const double2*)sp.DA, sp.DB, accW[which], false); };
  auto coulmu = [&](const int which) { thomas((const double2*)sp.CA, sp.CB, accC[which], cfg.tpos != 0); };

  // ---- DRIFTE (src/ModRamDrift.f90:285-376) ------------------------------------
  // COUL: the same walk is COULEN (src/ModRamCoul.f90:133-221) -- coefficient (COULE+COULI)(K) * NECR*BANE(L), ghost cells
  // with COULEN's own ratios, pitch angles L >= 2 only, optionally with the SUMRC moment of the result
  auto ewalk = [&](auto coulTag, double* mom) {
    constexpr bool COUL = decltype(coulTag)::value;
    const int ntask = NPA * PG * cfg.nsegE;
    const int nround = (ntask + T - 1) / T;
    for (int rd = 0; rd < nround; ++rd) {
      const int e = tid + rd * T;
      const int pp = e % PG, lq = e / PG;
      const int l = lq % NPA, seg = lq / NPA;
      const int p = p0 + pp;
      const bool act = (e < ntask) && (p < P) && (p % NR != 0) && (!COUL || l >= 1);
      const int ka = 1 + seg * cfg.segE, kb = min(NE, ka + cfg.segE - 1);
      const int K0 = max(ka - 1, 1);
      double* col = sT + (size_t)l * RS + pp;           // F(K) at col[(K-1)*PG]
      double Fm1 = 0, F0 = 0, Fp1 = 0, Fp2 = 0, hi1 = 0, hi2 = 0;
      if (act) {
        double F1 = 0.0, Fz = 0.0;
        if (K0 <= 2) {
          const double f2 = col[PG];
          F1 = f2 * sp.GREL1 / sp.GREL2 * (COUL ? sp.cg1 : sp.sqrtA);
          Fz = F1 * sp.GRZERO / sp.GREL1 * (COUL ? sp.cg0 : sp.sqrtB);
        }
#define GETFK(K) (((K) > NE) ? 0.0 : (((K) >= 2) ? col[((K)-1) * PG] : (((K) == 1) ? F1 : Fz)))
        Fm1 = GETFK(K0 - 1); F0 = GETFK(K0); Fp1 = GETFK(K0 + 1); Fp2 = GETFK(K0 + 2);
        hi1 = GETFK(kb + 1); hi2 = GETFK(kb + 2);
#undef GETFK
      }
      if (cfg.nsegE > 1) __syncthreads();               // halo reads before anybody's in-place walk
      if (act) {
        const double fA = COUL ? sXc[l * PG + pp] : sEa[l * PG + pp], fB = COUL ? 0.0 : sEb[l * PG + pp];
        const double4* tab = (const double4*)(COUL ? sTabC : sTab) + (K0 - 1);
        double macc = 0.0;
        const double wl = sWMU[l];
        int Kc = K0 + 1;                                // 1-based energy of the cell the next step() writes
        double FBprev;
        double dm1 = F0 - Fm1, d0 = Fp1 - F0, dp1 = Fp2 - Fp1;
        {                                               // peeled first interface K0: flux only
          const double4 tb = *tab;
          const double c = fma(tb.y, fB, tb.x * fA);
          const double ac = fabs(c) * tb.z;
          FBprev = c * limited_flux_d(F0, Fp1, dm1, d0, dp1, c < 0.0, ac, beta);
        }
        double* pO = col + (size_t)K0 * PG;             // -> F(K0+1)
        auto step = [&](const double nn) {
          ++tab;
          F0 = Fp1; Fp1 = Fp2; Fp2 = nn;
          dm1 = d0; d0 = dp1; dp1 = Fp2 - Fp1;
          const double4 tb = *tab;
          const double c = fma(tb.y, fB, tb.x * fA);
          const double ac = fabs(c) * tb.z;
          const double FB = c * limited_flux_d(F0, Fp1, dm1, d0, dp1, c < 0.0, ac, beta);
          double fn = fma(-(FB - FBprev), tb.w, F0);
          if (fn < 0.0) fn = 1E-15;
          *pO = fn;
          pO += PG;
          FBprev = FB;
          if (mom) macc = fma(fn, sWE[Kc - 1] * sEK[Kc - 1], macc);
          ++Kc;
        };
        int K = K0 + 1;
#pragma unroll 4
        for (; K + 2 <= kb; ++K) step(pO[2 * PG]);      // F(K+2): own cell, not yet rewritten
        if (K + 1 <= kb) { step(hi1); ++K; }            // K = kb-1: F(kb+1)
        if (K <= kb) step(hi2);                         // K = kb:   F(kb+2)
        if (mom && p < (NT - 1) * NR) *mom += macc * wl;
      }
    }
  };
  auto drifte = [&]() { ewalk(std::false_type{}, nullptr); };
  auto coulen = [&](const int which) { ewalk(std::true_type{}, &accC[which]); };

  // ---- DRIFTMU (src/ModRamDrift.f90:382-473), optionally with the SUMRC moment ----
  auto driftmu = [&](const bool mom) {
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
        const double wM = sWM[k];
        const double* ca = sMa + (la - 1) * PG + pp;    // coefficient pieces of L = la
        const double* cb = sMb + (la - 1) * PG + pp;
        double FBprev = 0.0, macc = 0.0, fnew = 0.0;
        double dm1 = F0 - Fm1, d0 = Fp1 - F0, dp1 = Fp2 - Fp1;
        if (la > 2) {                                   // flux through the segment's lower edge
          const double c = fma(wM, cb[-PG], ca[-PG]);
          FBprev = c * limited_flux_d(Fm1, F0, Fm1 - Fm2, dm1, d0, c < 0.0, fabs(c) * sRD[la - 2], beta);
        }
        double* pO = col + (size_t)(la - 1) * RS;
        const double* rd_ = sRD + (la - 1);
        const double* rw_ = sRW + (la - 1);
        const double* wm_ = sWMU + (la - 1);
        // cells L <= NPA-2 (limited flux); the value entering the window after cell L is F(L+3)
        auto step = [&](const double nn, const bool limited) {
          const double c = fma(wM, *cb, *ca);
          ca += PG; cb += PG;
          const double ac = fabs(c) * (*rd_++);
          double FB;
          if (limited) FB = c * limited_flux_d(F0, Fp1, dm1, d0, dp1, c < 0.0, ac, beta);
          else FB = c * Fp1;                            // FBND(NPA-1) = F(NPA)  (:458)
          fnew = fma(-(FB - FBprev), *rw_++, F0);
          FBprev = FB;
          if (fnew < 0.0) fnew = 1E-15;
          *pO = fnew;
          pO += RS;
          if (mom) macc = fma(fnew, *wm_, macc);
          ++wm_;
          F0 = Fp1; Fp1 = Fp2; Fp2 = nn;
          dm1 = d0; d0 = dp1; dp1 = Fp2 - Fp1;
        };
        int L = la;
        const int lim_end = min(lb, NPA - 2);           // last cell with a limited upper flux
#pragma unroll 4
        for (; L + 3 <= lim_end; ++L) step(pO[3 * (size_t)RS], true);
        for (; L <= lb; ++L) {                          // the (at most 3) cells that read the halo
          const double nn = (L + 3 <= lb) ? pO[3 * (size_t)RS] : ((L + 3 == lb + 1) ? hi1 : ((L + 3 == lb + 2) ? hi2 : 0.0));
          step(nn, L <= NPA - 2);
        }
        if (lastseg) {
          const double fN = fnew * d.FNHSc[(size_t)(NPA - 1) * Pp + p] * d.MU[NPA - 1] / d.FNHSc[(size_t)(NPA - 2) * Pp + p] / d.MU[NPA - 2];
          *pO = fN;                                     // :466
          if (mom) macc = fma(fN, sWMU[NPA - 1], macc);
        }
        if (mom && k >= 1 && p < (NT - 1) * NR) acc[0] += macc * (sWE[k] * sEK[k]);
      }
    }
  };

  // ---- the loss block (k_loss_mid): pointwise; a thread walks the pitch angles of one
  // (energy, position) so everything but the pitch-angle factors is hoisted ------------
  auto losses = [&]() {
    const bool ion = (sp.kind != 3);
    const bool doA = (cfg.doA >> (s0 + blockIdx.y)) & 1;
    const int ntask = NE * PG * cfg.nsegL;
    const int nround = (ntask + T - 1) / T;
    const double DTs = d.DTs;
    for (int rd = 0; rd < nround; ++rd) {
      const int e = tid + rd * T;
      const int pp = e % PG, kq = e / PG;
      const int k = kq % NE, seg = kq / NE;
      const int p = p0 + pp;
      const int i = p % NR;
      if (e >= ntask || k < 1 || p >= P || i < 1) continue;
      const int la = seg * cfg.segL, lb = min(NPA, la + cfg.segL);     // 0-based [la, lb)
      const int lcone = max(la, min(lb, d.UPA[i] - 1));                 // ATMOL acts on l >= UPA-1
      const double svk = sSV[k], xk = sX[k * PG + pp], wfk = sW[k * PG + pp];
      double* col = sT + k * PG + pp;
      double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
      auto cell = [&](const int l, const bool useA, const bool cone) {
        double f = col[(size_t)l * RS];
        const double wm = sWMZ[l];
        double facA = 1.0;
        if (useA) {
          facA = ion ? fast_exp(-(svk * sH[l * PG + pp] * DTs), sE2) : wfk;
          f = f * facA;
        }
        s1 = fma(wm, f, s1);
        if (cone) {
          const double a = fast_exp(xk * sRF[l * PG + pp], sE2);
          f = f * a;
          s2 = fma(wm, f, s2);
          f = f * a;
          s3 = fma(wm, f, s3);
        } else {
          s2 = fma(wm, f, s2);
          s3 = fma(wm, f, s3);
        }
        if (useA) f = f * facA;
        s4 = fma(wm, f, s4);
        col[(size_t)l * RS] = f;
      };
      int l = la;
      if (l == 0 && l < lb) { cell(0, false, lcone == 0); ++l; }      // L=1: the first/last operator skips it
      if (doA) {
        for (; l < lcone; ++l) cell(l, true, false);
        for (; l < lb; ++l) cell(l, true, true);
      } else {
        for (; l < lcone; ++l) cell(l, false, false);
        for (; l < lb; ++l) cell(l, false, true);
      }
      if (p < (NT - 1) * NR) {
        const double we = sEK[k] * sWE[k];
        acc[1] += we * s1; acc[2] += we * s2; acc[3] += we * s3; acc[4] += we * s4;
      }
    }
  };

  // (the CFL limits DtDriftE/Mu come from k_cfl_fast, not from the sweeps)
  drifte();
  __syncthreads();
  driftmu(true);
  __syncthreads();
  if constexpr (CO) if (doC) { coulen(0); __syncthreads(); coulmu(1); __syncthreads(); }   // src/ModRamRun.f90:79-88
  if (WPI && doW) { wpadif(0); __syncthreads(); }       // :91-104
  losses();
  __syncthreads();
  if (WPI && doW) { wpadif(1); __syncthreads(); }       // :140-154
  if constexpr (CO) if (doC) { coulmu(2); __syncthreads(); coulen(3); __syncthreads(); }   // :156-165
  driftmu(false);
  __syncthreads();
  drifte();
  __syncthreads();

  // ---- write the block back, reductions -------------------------------------------
  {
    constexpr int H = PG / 2;
    const int rowItems = NE * H;
    int l = tid / rowItems, r = tid - l * rowItems;
    const int dl = T / rowItems, dr = T - dl * rowItems;
    const int pad = (NEs - NE) * PG;
    int so = l * RS + (r / H) * PG + 2 * (r % H);
    int go = (l * NE + r / H) * Pp + p0 + 2 * (r % H);
    const int so_step = dl * RS + (dr / H) * PG + 2 * (dr % H);
    const int go_step = (dl * NE + dr / H) * Pp + 2 * (dr % H);
    int hh = r % H;
    const ptrdiff_t rel = PEER ? sp.F - pv.F[pv.gidx] : 0;
    while (l < NPA) {
      // PEER: the reverse plane kernel runs next, on the rank that owns the pitch angle: store there
      if (PEER) *(double2*)(pv.F[peer_owner(pv.lcut, pv.G, l)] + rel + go) = *(const double2*)(sT + so);
      else *(double2*)(sp.F + go) = *(const double2*)(sT + so);
      r += dr; l += dl; so += so_step; go += go_step; hh += dr % H;
      if (hh >= H) { hh -= H; so += PG - 2 * H; go += Pp - 2 * H; }
      if (r >= rowItems) { r -= rowItems; ++l; so += pad; }
    }
  }
  block_sum_to<5>(sp.part, cfg.pb0 + blockIdx.x, acc, sRed);
  if (WPI) {
    __syncthreads();
    block_sum_to<2>(sp.part + cfg.wpart_off, cfg.pb0 + blockIdx.x, accW, sRed);
    if (doC) {
      __syncthreads();
      block_sum_to<4>(sp.part + cfg.cpart_off, cfg.pb0 + blockIdx.x, accC, sRed);
    }
  }
}

// =============================================================================
// k_finalize_wpi: the two WPADIF moments of the fused step (slots 1 and 8 for the electrons'
// WPI diffusion, 2 and 7 for the EMIC diffusion of H+; src/ModRamRun.f90:91-104, :140-154) and
// the violation count, after k_finalize.  grid: x = 2, y = species; block = 256
// =============================================================================
__global__ void __launch_bounds__(256) k_finalize_wpi(const __grid_constant__ SpecPack pk, int s0, int doW, int nb_col, int wpart_off,
                                                      int res_n, int nsum, const unsigned long long* __restrict__ viol,
                                                      unsigned long long* __restrict__ host_res) {
  __shared__ double sm[32];
  const int s = s0 + blockIdx.y;
  if (!((doW >> s) & 1)) return;
  const SpecDev& sp = pk.s[s];
  const int q = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < nb_col; b += blockDim.x) acc += sp.part[wpart_off + (size_t)b * 2 + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      unsigned long long* hr = host_res + (size_t)s * res_n;
      const bool el = (sp.kind == 3);
      const int slot = (q == 0) ? (el ? 1 : 2) : (el ? 8 : 7);
      sp.dt[4 + slot] = dbl_bits(v);
      hr[4 + slot] = dbl_bits(v);
      if (q == 0) { sp.dt[4 + nsum] = viol[s]; hr[4 + nsum] = viol[s]; }
    }
  }
}

// k_finalize_coul: the four Coulomb moments of the fused step (slots 10..13: COULEN, COULMU | COULMU, COULEN;
// src/ModRamRun.f90:79-88, :156-165).  grid: x = 4, y = species; block = 256
__global__ void __launch_bounds__(256) k_finalize_coul(const __grid_constant__ SpecPack pk, int s0, int nb_col, int cpart_off, int res_n,
                                                       unsigned long long* __restrict__ host_res) {
  __shared__ double sm[32];
  const int s = s0 + blockIdx.y;
  const SpecDev& sp = pk.s[s];
  const int q = blockIdx.x;
  double acc = 0.0;
  for (int b = threadIdx.x; b < nb_col; b += blockDim.x) acc += sp.part[cpart_off + (size_t)b * 4 + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      sp.dt[4 + 10 + q] = dbl_bits(v);
      host_res[(size_t)s * res_n + 4 + 10 + q] = dbl_bits(v);
    }
  }
}

// =============================================================================
// k_finalize: second stage of the SUMRC reductions of the fused step and the result
// block of every species: DtDrift* from the cached CFL limits, moment slots 0, 3..6
// (column kernel partials) and 9 (reverse plane kernel partials), zero elsewhere.
// Blocks x >= 6 do the energy sums of ANISCH (src/ModRamRun.f90:379-400) for a tile of plane
// positions: PPERT/PPART = RFAC * sum_K tE|tA (the reference groups K into five bands with the
// same factor; only the summation order differs).
// Everything is written to the device blocks and to the host-mapped copies: no memcpy nodes.
// grid: x = 6 moments + tiles of 32 positions, y = species; block = 256
// =============================================================================
__global__ void __launch_bounds__(256) k_finalize(const __grid_constant__ RamDev d, const __grid_constant__ SpecPack pk, int s0,
                                                  int nb_col, int off_rev, int nb_rev,
                                                  const unsigned long long* __restrict__ cfl_all, int res_n, int nsum,
                                                  unsigned long long* __restrict__ host_res, double RFAC, double* __restrict__ host_pp,
                                                  int anisch_rows, int anisch_l0) {
  __shared__ double sm[32];
  const int s = s0 + blockIdx.y;
  const SpecDev& sp = pk.s[s];
  if (blockIdx.x >= 6) {
    // 32 positions x 8 lanes over the rows; the 8 partial sums are combined in a fixed order
    __shared__ double sE[8][32], sA[8][32];
    const int pl = threadIdx.x & 31, kl = threadIdx.x >> 5;
    const int p = (blockIdx.x - 6) * 32 + pl;
    double pe = 0.0, pa = 0.0;
    const bool act = (p < d.P) && (p % d.NR >= 1);
    if (anisch_rows > 0) {
      // rows written by k_plane_rp<REV>: one per (slab pitch angle, energy chunk)
      if (act)
        for (int r = kl; r < anisch_rows; r += 8) { pe += sp.aE2[(size_t)r * d.Pp + p]; pa += sp.aA2[(size_t)r * d.Pp + p]; }
      // F2(S,I,J,K,1) = F2(S,I,J,K,2) for I, K >= 2 (:366); the slab that holds L = 1, 2 does it
      if (anisch_l0 == 0 && act) {
        const size_t LS = (size_t)d.NE * d.Pp;
        for (int k = 1 + kl; k < d.NE; k += 8) sp.F[(size_t)k * d.Pp + p] = sp.F[LS + (size_t)k * d.Pp + p];
      }
    } else if (act) {
      for (int k = 1 + kl; k < d.NE; k += 8) { pe += sp.tE[(size_t)k * d.Pp + p]; pa += sp.tA[(size_t)k * d.Pp + p]; }
    }
    sE[kl][pl] = pe;
    sA[kl][pl] = pa;
    __syncthreads();
    if (kl == 0 && p < d.P) {
      for (int c = 1; c < 8; ++c) { pe += sE[c][pl]; pa += sA[c][pl]; }
      pe = RFAC * pe;
      pa = 2 * RFAC * pa;
      sp.pper[p] = pe;
      sp.ppar[p] = pa;
      double* hb = host_pp + (size_t)s * 2 * d.Pp;
      hb[p] = pe;
      hb[d.Pp + p] = pa;
    }
    return;
  }
  const int q = blockIdx.x;                              // 0..4: column moments, 5: reverse DRIFTR
  double acc = 0.0;
  if (q < 5) for (int b = threadIdx.x; b < nb_col; b += blockDim.x) acc += sp.part[(size_t)b * 5 + q];
  else for (int b = threadIdx.x; b < nb_rev; b += blockDim.x) acc += sp.part[off_rev + b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  unsigned long long* hr = host_res + (size_t)s * res_n;
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      const int slot = (q == 0) ? 0 : ((q < 5) ? q + 2 : 9);
      sp.dt[4 + slot] = dbl_bits(v);
      hr[4 + slot] = dbl_bits(v);
    }
  }
  if (q == 0) {
    // the rest of the block: CFL limits, unused moment slots, counters
    for (int t = threadIdx.x; t < res_n; t += blockDim.x) {
      const int slot = t - 4;
      if (t >= 4 && t < 4 + nsum && (slot == 0 || (slot >= 3 && slot <= 6) || slot == 9)) continue;
      const unsigned long long v = (t < 4) ? cfl_all[4 * (size_t)s + t] : 0ull;
      sp.dt[t] = v;
      hr[t] = v;
    }
  }
}
