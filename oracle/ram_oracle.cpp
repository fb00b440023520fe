// =============================================================================
// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU oracle for the RAM hot path: a restatement, in plain C++ (FP64, no
// fast-math, no FMA contraction, the reference's loop nests, operation order
// and S-fastest array layout), of the Fortran operators of lanl/RAM-SCB:
//
//   DRIFTPARA/DRIFTR/DRIFTP/DRIFTE/DRIFTMU   src/ModRamDrift.f90:36-473
//   CEPARA/CHAREXCHANGE/ATMOL                src/ModRamLoss.f90:19-170,457-507
//   WAVELO/WPADIF                            src/ModRamWPI.f90:580-714
//   COULPARA/COULEN/COULMU                   src/ModRamCoul.f90:17-296
//   FLCscatter                               src/ModRamLoss.f90:513-575
//   SUMRC/ANISCH(moments)/ram_run            src/ModRamRun.f90:16-415
//   Gcoul/FUNT/FUNI                          src/ModRamFunctions.f90:72-143
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library, and only as the checker / CPU baseline.
//
// PARITY PINNING: the reference cannot be compiled in this environment (no
// Fortran compiler, no GSL/NetCDF; its input blobs are missing), and its own
// test-suite holds no per-operator vectors for these routines.  What *is*
// pinned: Gcoul against the reference's known-answer test
// (src/ModRamFunctions.f90:481-482) and the energy ladder feeding every
// operator against output/test1/dsbnd.ref (tests/golden/).  The operators
// themselves are "parity unpinned": validated by line-by-line review against
// the cited Fortran, by conservation / positivity / symmetry properties, and by
// a second, independent numpy restatement of the drifts, losses, SUMRC, ANISCH
// and WPADIF (tests/independent_ram.py) that agrees with this file bit for bit.
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

struct SpeciesWork {   // ModRamDrift.f90:15-19 THREADPRIVATE work arrays
  int QS = 1;
  std::vector<double> VR, P1, P2, MUDOT, EDOT, CDriftR, CDriftP, CDriftE, CDriftMu;
};

struct Orc {
  int nS, NR, NT, NE, NPA;
  std::map<std::string, double*> d;
  std::map<std::string, int*> i;
  std::map<std::string, double> s;
  std::vector<SpeciesWork> w;
  double* D(const char* n) {
    auto it = d.find(n);
    if (it == d.end() || !it->second) { std::fprintf(stderr, "oracle: array %s not set\n", n); std::abort(); }
    return it->second;
  }
  int* I(const char* n) {
    auto it = i.find(n);
    if (it == i.end() || !it->second) { std::fprintf(stderr, "oracle: int array %s not set\n", n); std::abort(); }
    return it->second;
  }
  double S(const char* n) {
    auto it = s.find(n);
    if (it == s.end()) { std::fprintf(stderr, "oracle: scalar %s not set\n", n); std::abort(); }
    return it->second;
  }
};

// 1-based, column-major accessors ------------------------------------------------
#define A1(a, i) (a)[(size_t)((i)-1)]
#define A2(a, n1, i, j) (a)[(size_t)((i)-1) + (size_t)(n1) * (size_t)((j)-1)]
#define A3(a, n1, n2, i, j, k) (a)[(size_t)((i)-1) + (size_t)(n1) * ((size_t)((j)-1) + (size_t)(n2) * (size_t)((k)-1))]
#define A4(a, n1, n2, n3, i, j, k, l) \
  (a)[(size_t)((i)-1) + (size_t)(n1) * ((size_t)((j)-1) + (size_t)(n2) * ((size_t)((k)-1) + (size_t)(n3) * (size_t)((l)-1)))]
#define A5(a, n1, n2, n3, n4, i, j, k, l, m)                                                             \
  (a)[(size_t)((i)-1) +                                                                                  \
      (size_t)(n1) * ((size_t)((j)-1) +                                                                  \
                      (size_t)(n2) * ((size_t)((k)-1) + (size_t)(n3) * ((size_t)((l)-1) + (size_t)(n4) * (size_t)((m)-1))))]

// commonly used views (macros rely on local nS,NR,NT,NE,NPA)
#define F2_(S, I, J, K, L) A5(F2, nS, NR, NT, NE, S, I, J, K, L)
#define BNES_(I, J) A2(BNES, NR + 1, I, J)
#define VT_(I, J) A2(VT, NR + 1, I, J)
#define EIR_(I, J) A2(EIR, NR + 1, I, J)
#define EIP_(I, J) A2(EIP, NR + 1, I, J)
#define DBDT_(I, J) A2(dBdt, NR + 1, I, J)
#define FNHS_(I, J, L) A3(FNHS, NR + 1, NT, I, J, L)
#define FNIS_(I, J, L) A3(FNIS, NR + 1, NT, I, J, L)
#define BOUNHS_(I, J, L) A3(BOUNHS, NR + 1, NT, I, J, L)
#define BOUNIS_(I, J, L) A3(BOUNIS, NR + 1, NT, I, J, L)
#define HDNS_(I, J, L) A3(HDNS, NR + 1, NT, I, J, L)
#define DIDT_(I, J, L) A3(dIdt, NR + 1, NT, I, J, L)
#define DIBNDT_(I, J, L) A3(dIbndt, NR + 1, NT, I, J, L)
#define OUT_(I, J) A2(outsideMGNP, NR, I, J)
#define GREL_(S, K) A2(GREL, nS, S, K)
#define GRBND_(S, K) A2(GRBND, nS, S, K)
#define V_(S, K) A2(V, nS, S, K)
#define VBND_(S, K) A2(VBND, nS, S, K)
#define CD4(a, I, J, K, L) A4(a, NR, NT, NE, I, J, K, L)

#define DIMS                                                      \
  const int nS = o->nS, NR = o->NR, NT = o->NT, NE = o->NE, NPA = o->NPA; \
  (void)nS; (void)NR; (void)NT; (void)NE; (void)NPA;

inline double sq(double x) { return x * x; }

// -----------------------------------------------------------------------------
// src/ModRamFunctions.f90:72-143
double Gcoul(double x) {
  const double PI = 3.1415926535897932384626433832795;
  double G1 = std::erf(x) - 2. * x / std::sqrt(PI) * std::exp(-x * x);
  return G1 / 2. / x / x;
}
double FUNT(double x) {
  const double PI = 3.1415926535897932384626433832795;
  double Y = std::sqrt(1 - x * x);
  double ALPHA = 1. + std::log(2. + std::sqrt(3.)) / 2. / std::sqrt(3.);
  double BETA = ALPHA / 2. - PI * std::sqrt(2.) / 12.;
  double a1 = 0.055, a2 = -0.037, a3 = -0.074, a4 = 0.056;
  return ALPHA - BETA * (Y + std::sqrt(Y)) + a1 * std::pow(Y, 1. / 3.) + a2 * std::pow(Y, 2. / 3.) + a3 * Y +
         a4 * std::pow(Y, 4. / 3.);
}
double FUNI(double x) {
  const double PI = 3.1415926535897932384626433832795;
  double ylog = 0.0;
  double Y = std::sqrt(1 - x * x);
  if (Y > 0) ylog = std::log(Y);
  double ALPHA = 1. + std::log(2. + std::sqrt(3.)) / 2. / std::sqrt(3.);
  double BETA = ALPHA / 2. - PI * std::sqrt(2.) / 12.;
  double a1 = 0.055, a2 = -0.037, a3 = -0.074, a4 = 0.056;
  return 2. * ALPHA * (1. - Y) + 2. * BETA * Y * ylog + 4. * BETA * (Y - std::sqrt(Y)) +
         3. * a1 * (std::pow(Y, 1. / 3.) - Y) + 6. * a2 * (std::pow(Y, 2. / 3.) - Y) +
         6. * a4 * (Y - std::pow(Y, 4. / 3.)) - 2. * a3 * Y * ylog;
}

// -----------------------------------------------------------------------------
// DRIFTPARA  src/ModRamDrift.f90:36-88
void driftpara(Orc* o, int S) {
  DIMS
  SpeciesWork& w = o->w[S - 1];
  const double DTs = o->S("DTs"), MDR = o->S("MDR"), DPHI = o->S("DPHI");
  const double *RLZ = o->D("RLZ"), *EKEV = o->D("EKEV"), *GREL = o->D("GREL"), *WMU = o->D("WMU"),
               *EBND = o->D("EBND"), *GRBND = o->D("GRBND"), *MU = o->D("MU");
  if (w.VR.empty()) {
    w.VR.assign(NR, 0.0); w.P1.assign(NR, 0.0); w.P2.assign((size_t)NR * NE, 0.0);
    w.EDOT.assign((size_t)NR * NE, 0.0); w.MUDOT.assign((size_t)NR * NPA, 0.0);
    size_t n4 = (size_t)NR * NT * NE * NPA;
    w.CDriftR.assign(n4, 0.0); w.CDriftP.assign(n4, 0.0); w.CDriftE.assign(n4, 0.0); w.CDriftMu.assign(n4, 0.0);
  }
  w.QS = A1(o->I("QS"), S);
  const double QS = (double)w.QS;
  for (int I = 1; I <= NR; ++I) {
    A1(w.VR, I) = DTs / MDR / (A1(RLZ, I) + 0.5 * MDR) / 2 / DPHI;
    A1(w.P1, I) = DTs / DPHI / 2 / MDR / A1(RLZ, I);
    for (int K = 1; K <= NE; ++K)
      A2(w.P2, NR, I, K) = DTs * A1(EKEV, K) * 1000 * (GREL_(S, K) + 1) / GREL_(S, K) / (A1(RLZ, I) * A1(RLZ, I)) / DPHI / QS;
  }
  for (int I = 1; I <= NR; ++I) {
    for (int L = 1; L <= NPA - 1; ++L) {
      double MUBOUN = A1(MU, L) + 0.5 * A1(WMU, L);
      A2(w.MUDOT, NR, I, L) = (1. - MUBOUN * MUBOUN) * DTs / 2 / MUBOUN / A1(RLZ, I);
    }
    A2(w.MUDOT, NR, I, NPA) = 0.;
    for (int K = 1; K <= NE; ++K)
      A2(w.EDOT, NR, I, K) = A1(EBND, K) * DTs / A1(RLZ, I) * (GRBND_(S, K) + 1) / GRBND_(S, K) / 2.;
  }
}

// common flux limiter, SURVEY appendix A.1 (ModRamDrift.f90:170-182 etc.)
#define LIMITED_FLUX(FBNDm, Fm, Fm1, FN, FNm1, sgn, chat, BetaLim)            \
  do {                                                                         \
    double X_ = (Fm1) - (Fm);                                                  \
    double FUP_ = 0.5 * ((Fm) + (Fm1) - (sgn)*X_);                             \
    if (std::fabs(X_) <= 1.E-27) (FBNDm) = FUP_;                               \
    if (std::fabs(X_) > 1.E-27) {                                              \
      double R_ = ((FN) - (FNm1)) / X_;                                        \
      if (R_ <= 0) (FBNDm) = FUP_;                                             \
      if (R_ > 0) {                                                            \
        double LIM_ = std::max(std::min((BetaLim)*R_, 1.), std::min(R_, (BetaLim))); \
        double CORR_ = -0.5 * ((chat) - (sgn)) * X_;                           \
        (FBNDm) = FUP_ + LIM_ * CORR_;                                         \
      }                                                                        \
    }                                                                          \
  } while (0)

// -----------------------------------------------------------------------------
// DRIFTR  src/ModRamDrift.f90:95-198
void driftr(Orc* o, int S) {
  DIMS
  SpeciesWork& w = o->w[S - 1];
  const double DTs = o->S("DTs"), MDR = o->S("MDR"), DPHI = o->S("DPHI"), BetaLim = o->S("BetaLim"),
               FracCFL = o->S("FracCFL"), CONF1 = o->S("CONF1"), CONF2 = o->S("CONF2");
  double* F2 = o->D("F2");
  const double *BNES = o->D("BNES"), *FNIS = o->D("FNIS"), *FNHS = o->D("FNHS"), *EKEV = o->D("EKEV"),
               *GREL = o->D("GREL"), *RLZ = o->D("RLZ"), *FGEOS = o->D("FGEOS"), *VT = o->D("VT"), *EIP = o->D("EIP");
  const int* outsideMGNP = o->I("outsideMGNP");
  double* DtDriftR = o->D("DtDriftR");
  const double QS = (double)w.QS;
  std::vector<int> sgn((size_t)NR * NT, 1);
  std::vector<double> CR((size_t)NR * NT, 0.0), F(NR + 2, 0.0), FBND(NR, 0.0);
  double* CDriftR = w.CDriftR.data();

  A1(DtDriftR, S) = 100000.0;
  for (int I = 1; I <= NR; ++I)
    for (int J = 1; J <= NT; ++J) {
      int J0 = J - 1; if (J == 1) J0 = NT - 1;
      int J1 = J + 1; if (J == NT) J1 = 2;
      A2(CR, NR, I, J) = A1(w.VR, I) * (VT_(I, J0) + VT_(I + 1, J0) - VT_(I, J1) - VT_(I + 1, J1)) / (BNES_(I, J) + BNES_(I + 1, J)) +
                         (EIP_(I, J) + EIP_(I + 1, J)) / (BNES_(I, J) + BNES_(I + 1, J)) * DTs / MDR;
    }
  for (int K = 1; K <= NE; ++K) {
    double P4 = DTs * A1(EKEV, K) * 1000.0 * (GREL_(S, K) + 1) / GREL_(S, K) / DPHI / MDR / QS;
    for (int L = 1; L <= NPA; ++L)
      for (int J = 1; J <= NT; ++J) {
        for (int I = 1; I <= NR; ++I) A1(F, I) = F2_(S, I, J, K, L);
        int J0 = J - 1; if (J == 1) J0 = NT - 1;
        int J1 = J + 1; if (J == NT) J1 = 2;
        for (int I = 1; I <= NR; ++I) {
          double CGR1 = FNIS_(I + 1, J1, L) + FNIS_(I, J1, L) - FNIS_(I + 1, J0, L) - FNIS_(I, J0, L);
          double CGR2 = BNES_(I + 1, J1) + BNES_(I, J1) - BNES_(I + 1, J0) - BNES_(I, J0);
          double CGR3 = CGR1 + (FNIS_(I + 1, J, L) + FNIS_(I, J, L) - 2 * FNHS_(I + 1, J, L) - 2 * FNHS_(I, J, L)) * CGR2 / 2. /
                                   (BNES_(I + 1, J) + BNES_(I, J));
          double CGR = CGR3 / (FNHS_(I, J, L) + FNHS_(I + 1, J, L)) * P4 / 2. / (BNES_(I, J) + BNES_(I + 1, J)) / (A1(RLZ, I) + 0.5 * MDR);
          CD4(CDriftR, I, J, K, L) = A2(CR, NR, I, J) + CGR;
          if (OUT_(I, J) == 0) {
            double ctemp = std::max(std::fabs(CD4(CDriftR, I, J, K, L)), 1E-10);
            A1(DtDriftR, S) = std::min(A1(DtDriftR, S), FracCFL * DTs / ctemp);
          }
          A2(sgn, NR, I, J) = 1;
          if (CD4(CDriftR, I, J, K, L) < 0) A2(sgn, NR, I, J) = -1;
        }
        int UR;
        if (A2(sgn, NR, NR, J) == 1) {
          A1(FBND, 1) = 0.;
          A1(FBND, NR) = A1(F, NR);
          UR = NR - 1;
        } else {
          A1(FBND, 1) = A1(F, 2);
          UR = NR;
          if (OUT_(NR, J) == 1) {
            A1(F, NR + 1) = 0.0;
            A1(F, NR + 2) = 0.0;
          } else {
            A1(F, NR + 1) = A4(FGEOS, nS, NT, NE, S, J, K, L) * CONF1 * FNHS_(NR, J, L);
            A1(F, NR + 2) = A4(FGEOS, nS, NT, NE, S, J, K, L) * CONF2 * FNHS_(NR, J, L);
          }
        }
        for (int I = 2; I <= UR; ++I) {
          const int sg = A2(sgn, NR, I, J);
          const int N = I + 1 - sg;
          LIMITED_FLUX(A1(FBND, I), A1(F, I), A1(F, I + 1), A1(F, N), A1(F, N - 1), sg, CD4(CDriftR, I, J, K, L), BetaLim);
        }
        for (int I = 2; I <= NR; ++I) {
          F2_(S, I, J, K, L) = F2_(S, I, J, K, L) - CD4(CDriftR, I, J, K, L) * A1(FBND, I) + CD4(CDriftR, I - 1, J, K, L) * A1(FBND, I - 1);
          if (F2_(S, I, J, K, L) < 0) F2_(S, I, J, K, L) = 1E-15;
        }
      }
  }
}

// -----------------------------------------------------------------------------
// DRIFTP  src/ModRamDrift.f90:204-279
void driftp(Orc* o, int S) {
  DIMS
  SpeciesWork& w = o->w[S - 1];
  const double DTs = o->S("DTs"), MDR = o->S("MDR"), DPHI = o->S("DPHI"), BetaLim = o->S("BetaLim"), FracCFL = o->S("FracCFL");
  double* F2 = o->D("F2");
  const double *BNES = o->D("BNES"), *FNIS = o->D("FNIS"), *FNHS = o->D("FNHS"), *RLZ = o->D("RLZ"), *VT = o->D("VT"),
               *EIR = o->D("EIR");
  const int* outsideMGNP = o->I("outsideMGNP");
  double* DtDriftP = o->D("DtDriftP");
  std::vector<double> FBND(NT, 0.0), F(NT, 0.0);
  double* CDriftP = w.CDriftP.data();

  A1(DtDriftP, S) = 100000.0;
  const double OME = 7.3E-5;
  for (int L = 1; L <= NPA; ++L)
    for (int K = 1; K <= NE; ++K)
      for (int I = 2; I <= NR; ++I) {
        for (int J = 1; J <= NT; ++J) A1(F, J) = F2_(S, I, J, K, L);
        for (int J = 2; J <= NT; ++J) {
          int J1 = J + 1; if (J == NT) J1 = 2;
          double GPA1 = FNIS_(I, J, L) + FNIS_(I, J1, L) +
                        (FNIS_(I + 1, J1, L) + FNIS_(I + 1, J, L) - FNIS_(I - 1, J, L) - FNIS_(I - 1, J1, L)) * A1(RLZ, I) / 2. / MDR;
          double GPA2 = A1(RLZ, I) / 4. / MDR * (FNIS_(I, J, L) + FNIS_(I, J1, L) - 2 * FNHS_(I, J, L) - 2 * FNHS_(I, J1, L)) *
                        (BNES_(I + 1, J1) + BNES_(I + 1, J) - BNES_(I - 1, J) - BNES_(I - 1, J1)) / (BNES_(I, J) + BNES_(I, J1));
          CD4(CDriftP, I, J, K, L) = ((VT_(I + 1, J) + VT_(I + 1, J1) - VT_(I - 1, J) - VT_(I - 1, J1)) * A1(w.P1, I) -
                                      A2(w.P2, NR, I, K) * (GPA1 + GPA2) / (FNHS_(I, J, L) + FNHS_(I, J1, L)) -
                                      (EIR_(I, J1) + EIR_(I, J)) / A1(RLZ, I) * DTs / DPHI) /
                                         (BNES_(I, J) + BNES_(I, J1)) +
                                     OME * DTs / DPHI;
          if (OUT_(I, J) == 0) {
            double ctemp = std::max(std::fabs(CD4(CDriftP, I, J, K, L)), 1E-10);
            A1(DtDriftP, S) = std::min(A1(DtDriftP, S), FracCFL * DTs / ctemp);
          }
          int sg = 1;
          if (CD4(CDriftP, I, J, K, L) < 0) sg = -1;
          int N = J + 1 - sg;
          if (N > NT) N = N - NT + 1;
          LIMITED_FLUX(A1(FBND, J), A1(F, J), A1(F, J1), A1(F, N), A1(F, N - 1), sg, CD4(CDriftP, I, J, K, L), BetaLim);
        }
        CD4(CDriftP, I, 1, K, L) = CD4(CDriftP, I, NT, K, L);
        A1(FBND, 1) = A1(FBND, NT);
        for (int J = 2; J <= NT; ++J) {
          F2_(S, I, J, K, L) = F2_(S, I, J, K, L) - CD4(CDriftP, I, J, K, L) * A1(FBND, J) + CD4(CDriftP, I, J - 1, K, L) * A1(FBND, J - 1);
          if (F2_(S, I, J, K, L) < 0) F2_(S, I, J, K, L) = 1E-15;
        }
        F2_(S, I, 1, K, L) = F2_(S, I, NT, K, L);
      }
}

// -----------------------------------------------------------------------------
// DRIFTE  src/ModRamDrift.f90:285-376
void drifte(Orc* o, int S) {
  DIMS
  SpeciesWork& w = o->w[S - 1];
  const double DTs = o->S("DTs"), MDR = o->S("MDR"), DPHI = o->S("DPHI"), BetaLim = o->S("BetaLim"), FracCFL = o->S("FracCFL");
  const double CS = 2.998E8, Q = 1.602E-19;
  double* F2 = o->D("F2");
  const double *BNES = o->D("BNES"), *FNIS = o->D("FNIS"), *FNHS = o->D("FNHS"), *dBdt = o->D("dBdt"), *dIdt = o->D("dIdt"),
               *EKEV = o->D("EKEV"), *WE = o->D("WE"), *RMAS = o->D("RMAS"), *RLZ = o->D("RLZ"), *EBND = o->D("EBND"),
               *GREL = o->D("GREL"), *GRBND = o->D("GRBND"), *DE = o->D("DE"), *VT = o->D("VT"), *EIR = o->D("EIR"),
               *EIP = o->D("EIP");
  const int* outsideMGNP = o->I("outsideMGNP");
  double* DtDriftE = o->D("DtDriftE");
  const double QS = (double)w.QS;
  std::vector<double> FBNDv(NE, 0.0), Fv(NE + 3, 0.0);
  double* FBND = FBNDv.data();
  double* F0 = Fv.data();  // F(0:NE+2) -> F0[k]
  double* CDriftE = w.CDriftE.data();

  A1(DtDriftE, S) = 10000.0;
  const double OME = 7.3E-5;
  const double EZERO = A1(EKEV, 1) - A1(WE, 1);
  const double GRZERO = 1. + EZERO * 1000. * Q / A1(RMAS, S) / CS / CS;
  F0[NE + 1] = 0.;
  F0[NE + 2] = 0.;
  for (int J = 1; J <= NT; ++J) {
    int J0 = J - 1; if (J == 1) J0 = NT - 1;
    int J2 = J + 1; if (J == NT) J2 = 2;
    for (int I = 2; I <= NR; ++I) {
      double DRD1 = (EIP_(I, J) * A1(RLZ, I) - (VT_(I, J2) - VT_(I, J0)) / 2. / DPHI) / BNES_(I, J);
      double DPD1 = OME * A1(RLZ, I) + ((VT_(I + 1, J) - VT_(I - 1, J)) / 2 / MDR - EIR_(I, J)) / BNES_(I, J);
      for (int L = 1; L <= NPA; ++L) {
        double GPA = (1. - FNIS_(I, J, L) / 2. / FNHS_(I, J, L)) / BNES_(I, J);
        double GPR1 = GPA * (BNES_(I + 1, J) - BNES_(I - 1, J)) / 2. / MDR;
        double GPR2 = -FNIS_(I, J, L) / FNHS_(I, J, L) / A1(RLZ, I);
        double GPR3 = -(FNIS_(I + 1, J, L) - FNIS_(I - 1, J, L)) / 2. / MDR / FNHS_(I, J, L);
        double GPP1 = GPA * (BNES_(I, J2) - BNES_(I, J0)) / 2. / DPHI;
        double GPP2 = -(FNIS_(I, J2, L) - FNIS_(I, J0, L)) / 2. / DPHI / FNHS_(I, J, L);
        double DRD2 = (FNIS_(I, J2, L) - FNIS_(I, J0, L)) / 2. / DPHI +
                      (FNIS_(I, J, L) - 2 * FNHS_(I, J, L)) * (BNES_(I, J2) - BNES_(I, J0)) / 4 / BNES_(I, J) / DPHI;
        double DPD2 = FNIS_(I, J, L) + (FNIS_(I + 1, J, L) - FNIS_(I - 1, J, L)) * A1(RLZ, I) / 2 / MDR +
                      A1(RLZ, I) * (FNIS_(I, J, L) - 2 * FNHS_(I, J, L)) / 4 / MDR * (BNES_(I + 1, J) - BNES_(I - 1, J)) / BNES_(I, J);
        for (int K = 1; K <= NE; ++K) F0[K] = F2_(S, I, J, K, L);
        F0[1] = F0[2] * GREL_(S, 1) / GREL_(S, 2) * std::sqrt((sq(GREL_(S, 2)) - 1) / (sq(GREL_(S, 1)) - 1));
        F0[0] = F0[1] * GRZERO / GREL_(S, 1) * std::sqrt((sq(GREL_(S, 1)) - 1) / (sq(GRZERO) - 1));
        for (int K = 1; K <= NE; ++K) {
          double EDT1 = A1(EBND, K) * 1e3 * (GRBND_(S, K) + 1) / 2 / GRBND_(S, K) / FNHS_(I, J, L) / A1(RLZ, I) / BNES_(I, J) / QS;
          double DRDT = DRD1 + EDT1 * DRD2 * A1(RLZ, I);
          double DPDT = DPD1 - EDT1 * DPD2;
          double dBdt1 = DBDT_(I, J) * (1. - FNIS_(I, J, L) / 2. / FNHS_(I, J, L)) * A1(RLZ, I) / BNES_(I, J);
          double dIdt1 = -DIDT_(I, J, L) * A1(RLZ, I) / FNHS_(I, J, L);
          CD4(CDriftE, I, J, K, L) = A2(w.EDOT, NR, I, K) * ((GPR1 + GPR2 + GPR3) * DRDT + (GPP1 + GPP2) * DPDT + dBdt1 + dIdt1);
          if (OUT_(I, J) == 0) {
            double ctemp = std::max(std::fabs(CD4(CDriftE, I, J, K, L)), 1E-10);
            A1(DtDriftE, S) = std::min(A1(DtDriftE, S), FracCFL * DTs * A1(DE, K) / ctemp);
          }
          int sg = 1;
          if (CD4(CDriftE, I, J, K, L) < 0) sg = -1;
          int N = K + 1 - sg;
          LIMITED_FLUX(A1(FBND, K), F0[K], F0[K + 1], F0[N], F0[N - 1], sg, CD4(CDriftE, I, J, K, L) / A1(DE, K), BetaLim);
        }
        for (int K = 2; K <= NE; ++K) {
          F2_(S, I, J, K, L) = F2_(S, I, J, K, L) - CD4(CDriftE, I, J, K, L) / A1(WE, K) * A1(FBND, K) +
                               CD4(CDriftE, I, J, K - 1, L) / A1(WE, K) * A1(FBND, K - 1);
          if (F2_(S, I, J, K, L) < 0) F2_(S, I, J, K, L) = 1E-15;
        }
      }
    }
  }
}

// -----------------------------------------------------------------------------
// DRIFTMU  src/ModRamDrift.f90:382-473
void driftmu(Orc* o, int S) {
  DIMS
  SpeciesWork& w = o->w[S - 1];
  const double DTs = o->S("DTs"), MDR = o->S("MDR"), DPHI = o->S("DPHI"), BetaLim = o->S("BetaLim"), FracCFL = o->S("FracCFL");
  double* F2 = o->D("F2");
  const double *BNES = o->D("BNES"), *BOUNIS = o->D("BOUNIS"), *BOUNHS = o->D("BOUNHS"), *FNHS = o->D("FNHS"), *dBdt = o->D("dBdt"),
               *dIbndt = o->D("dIbndt"), *RLZ = o->D("RLZ"), *GREL = o->D("GREL"), *EKEV = o->D("EKEV"), *DMU = o->D("DMU"),
               *WMU = o->D("WMU"), *MU = o->D("MU"), *VT = o->D("VT"), *EIP = o->D("EIP"), *EIR = o->D("EIR");
  const int* outsideMGNP = o->I("outsideMGNP");
  double* DtDriftMu = o->D("DtDriftMu");
  const double QS = (double)w.QS;
  std::vector<double> FBND(NPA, 0.0), F(NPA, 0.0);
  double* CDriftMu = w.CDriftMu.data();

  A1(DtDriftMu, S) = 10000.0;
  const double OME = 7.3E-5;
  for (int K = 1; K <= NE; ++K)
    for (int J = 1; J <= NT; ++J) {
      int J0 = J - 1; if (J == 1) J0 = NT - 1;
      int J1 = J + 1; if (J == NT) J1 = 2;
      for (int I = 2; I <= NR; ++I) {
        for (int L = 1; L <= NPA; ++L) A1(F, L) = F2_(S, I, J, K, L);
        A1(F, 1) = A1(F, 2);
        double DRM1 = (EIP_(I, J) * A1(RLZ, I) - (VT_(I, J1) - VT_(I, J0)) / 2 / DPHI) / BNES_(I, J);
        double DPM1 = OME * A1(RLZ, I) + ((VT_(I + 1, J) - VT_(I - 1, J)) / 2 / MDR - EIR_(I, J)) / BNES_(I, J);
        for (int L = 2; L <= NPA; ++L) {
          double CMUDOT = A2(w.MUDOT, NR, I, L) * BOUNIS_(I, J, L) / BOUNHS_(I, J, L);
          double GMR1 = (BNES_(I + 1, J) - BNES_(I - 1, J)) / 4 / MDR / BNES_(I, J);
          double GMR2 = 1 / A1(RLZ, I);
          double GMR3 = (BOUNIS_(I + 1, J, L) - BOUNIS_(I - 1, J, L)) / 2 / MDR / BOUNIS_(I, J, L);
          double GMP1 = (BNES_(I, J1) - BNES_(I, J0)) / 4 / DPHI / BNES_(I, J);
          double GMP2 = (BOUNIS_(I, J1, L) - BOUNIS_(I, J0, L)) / 2 / DPHI / BOUNIS_(I, J, L);
          double EDT = A1(EKEV, K) * 1e3 * (GREL_(S, K) + 1) / 2 / GREL_(S, K) / BOUNHS_(I, J, L) / A1(RLZ, I) / BNES_(I, J) / QS;
          double DRM2 = (BOUNIS_(I, J1, L) - BOUNIS_(I, J0, L)) / 2 / DPHI +
                        (BOUNIS_(I, J, L) - 2 * BOUNHS_(I, J, L)) * (BNES_(I, J1) - BNES_(I, J0)) / 4 / BNES_(I, J) / DPHI;
          double DPM2 = BOUNIS_(I, J, L) + (BOUNIS_(I + 1, J, L) - BOUNIS_(I - 1, J, L)) * A1(RLZ, I) / 2 / MDR +
                        (BOUNIS_(I, J, L) - 2 * BOUNHS_(I, J, L)) * A1(RLZ, I) / 4 / MDR * (BNES_(I + 1, J) - BNES_(I - 1, J)) / BNES_(I, J);
          double DRDM = DRM1 + EDT * DRM2 * A1(RLZ, I);
          double DPDM = DPM1 - EDT * DPM2;
          double dBdt2 = DBDT_(I, J) / 2. / BNES_(I, J) * A1(RLZ, I);
          double dIbndt2 = DIBNDT_(I, J, L) * A1(RLZ, I) / BOUNIS_(I, J, L);
          CD4(CDriftMu, I, J, K, L) = -CMUDOT * ((GMR1 + GMR2 + GMR3) * DRDM + (GMP1 + GMP2) * DPDM + dBdt2 + dIbndt2);
          if (OUT_(I, J) == 0) {
            double ctemp = std::max(std::fabs(CD4(CDriftMu, I, J, K, L)), 1E-32);
            A1(DtDriftMu, S) = std::min(A1(DtDriftMu, S), FracCFL * DTs * A1(DMU, L) / ctemp);
          }
          int ISGM = 1;
          if (CD4(CDriftMu, I, J, K, L) < 0.0) ISGM = -1;
          if (L <= NPA - 2) {
            int N = L + 1 - ISGM;
            LIMITED_FLUX(A1(FBND, L), A1(F, L), A1(F, L + 1), A1(F, N), A1(F, N - 1), ISGM, CD4(CDriftMu, I, J, K, L) / A1(DMU, L), BetaLim);
          }
        }
        CD4(CDriftMu, I, J, K, 1) = 0.;
        A1(FBND, 1) = 0.;
        A1(FBND, NPA - 1) = A1(F, NPA);
        for (int L = 2; L <= NPA - 1; ++L) {
          F2_(S, I, J, K, L) = F2_(S, I, J, K, L) - CD4(CDriftMu, I, J, K, L) / A1(WMU, L) * A1(FBND, L) +
                               CD4(CDriftMu, I, J, K, L - 1) / A1(WMU, L) * A1(FBND, L - 1);
          if (F2_(S, I, J, K, L) < 0) F2_(S, I, J, K, L) = 1E-15;
        }
        F2_(S, I, J, K, NPA) = F2_(S, I, J, K, NPA - 1) * FNHS_(I, J, NPA) * A1(MU, NPA) / FNHS_(I, J, NPA - 1) / A1(MU, NPA - 1);
      }
    }
}

// -----------------------------------------------------------------------------
// CEPARA  src/ModRamLoss.f90:19-170 (the three built-in species; file-driven
// cross-sections of the 'default' branch are out of scope: needs GSL + data)
void cepara(Orc* o, int S) {
  DIMS
  const double DTs = o->S("DTs");
  const double *EKEV = o->D("EKEV"), *V = o->D("V"), *RLZ = o->D("RLZ"), *HDNS = o->D("HDNS");
  double *CHARGE = o->D("CHARGE"), *ATLOS = o->D("ATLOS");
  const int kind = A1(o->I("kind"), S);
  for (int L = 1; L <= NPA; ++L)
    for (int K = 1; K <= NE; ++K)
      for (int J = 1; J <= NT; ++J)
        for (int I = 1; I <= NR; ++I) A5(CHARGE, nS, NR, NT, NE, S, I, J, K, L) = 1.0;
  if (kind == 0 || kind == 1 || kind == 2) {
    for (int L = 2; L <= NPA; ++L)
      for (int K = 2; K <= NE; ++K)
        for (int I = 2; I <= NR; ++I)
          for (int J = 1; J <= NT; ++J) {
            double X = std::log10(A1(EKEV, K));
            if (X < -2.) X = -2.;
            double Y;
            if (kind == 0)  // Hydrogen :46
              Y = -18.767 - 0.11017 * X - 3.8173e-2 * (X * X) - 0.1232 * (X * X * X) - 5.0488e-2 * ((X * X) * (X * X));
            else if (kind == 2)  // HeliumP1 :61
              Y = -20.789 + 0.92316 * X - 0.68017 * (X * X) + 0.66153 * (X * X * X) - 0.20998 * ((X * X) * (X * X));
            else  // OxygenP1 :76
              Y = -18.987 - 0.10613 * X - 5.4841E-3 * (X * X) - 1.6262E-2 * (X * X * X) - 7.0554E-3 * ((X * X) * (X * X));
            double ALPHA = std::pow(10., Y) * V_(S, K) * HDNS_(I, J, L) * DTs;
            A5(CHARGE, nS, NR, NT, NE, S, I, J, K, L) = std::exp(-ALPHA);
          }
  }
  for (int K = 2; K <= NE; ++K)
    for (int I = 2; I <= NR; ++I) {
      double TAUB = 2 * A1(RLZ, I) / V_(S, K);
      A3(ATLOS, nS, NR, S, I, K) = std::exp(-DTs / TAUB);
    }
}

// CHAREXCHANGE  src/ModRamLoss.f90:457-478
void charexchange(Orc* o, int S) {
  DIMS
  double* F2 = o->D("F2");
  const double* CHARGE = o->D("CHARGE");
  for (int K = 2; K <= NE; ++K)
    for (int J = 1; J <= NT; ++J)
      for (int I = 2; I <= NR; ++I)
        for (int L = 2; L <= NPA; ++L) F2_(S, I, J, K, L) = F2_(S, I, J, K, L) * A5(CHARGE, nS, NR, NT, NE, S, I, J, K, L);
}

// ATMOL  src/ModRamLoss.f90:485-507
void atmol(Orc* o, int S) {
  DIMS
  double* F2 = o->D("F2");
  const double *FNHS = o->D("FNHS"), *UPA = o->D("UPA"), *ATLOS = o->D("ATLOS");
  for (int K = 2; K <= NE; ++K)
    for (int J = 1; J <= NT; ++J)
      for (int I = 2; I <= NR; ++I) {
        int u = (int)A1(UPA, I);
        for (int L = u; L <= NPA; ++L) F2_(S, I, J, K, L) = F2_(S, I, J, K, L) * std::pow(A3(ATLOS, nS, NR, S, I, K), 1 / FNHS_(I, J, L));
      }
}

// WAVELO  src/ModRamWPI.f90:580-636 (DoUsePlasmasphere = .false.)
void wavelo(Orc* o, int S) {
  DIMS
  const double DTs = o->S("DTs"), KP = o->S("Kp"), Kpmax12 = o->S("Kpmax12");
  double* F2 = o->D("F2");
  const double *LZ = o->D("LZ"), *EKEV = o->D("EKEV"), *WALOS1 = o->D("WALOS1"), *WALOS2 = o->D("WALOS2"), *WALOS3 = o->D("WALOS3");
  std::vector<double> RLpp(NT, 0.0);
  double Bw = 30.;
  if (KP >= 4.0) Bw = 100.;
  for (int J = 1; J <= NT; ++J) A1(RLpp, J) = 5.39 - 0.382 * Kpmax12;
  double TAU_LIF = 0.0;
  for (int K = 2; K <= NE; ++K)
    for (int I = 2; I <= NR; ++I)
      for (int J = 1; J <= NT; ++J)
        for (int L = 2; L <= NPA; ++L) {
          if (A1(LZ, I) <= A1(RLpp, J)) {
            TAU_LIF = A2(WALOS1, NR, I, K) * (sq(10. / Bw));
          } else if (A1(LZ, I) > A1(RLpp, J)) {
            if (A1(EKEV, K) <= 1000.) {
              TAU_LIF = A2(WALOS2, NR, I, K) * (1 + A2(WALOS3, NR, I, K) / A2(WALOS2, NR, I, K));
              if (A1(EKEV, K) <= 1.1)
                TAU_LIF = TAU_LIF * 37.5813 * std::exp(-1.81255 * A1(EKEV, K));
              else if (A1(EKEV, K) > 1.1 && A1(EKEV, K) <= 5.)
                TAU_LIF = TAU_LIF * (7.5 - 1.15 * A1(EKEV, K));
            } else if (A1(EKEV, K) > 1000.) {
              TAU_LIF = 5. * 3600 * 24 / KP;
            }
          }
          F2_(S, I, J, K, L) = F2_(S, I, J, K, L) * std::exp(-DTs / TAU_LIF);
        }
}

// WPADIF  src/ModRamWPI.f90:643-714.  mode 0: electrons (ATAW+ATAC); mode 1:
// ions with EMIC (ATAW_emic_h + ATAW_emic_he).  Returns the number of lines on
// which the reference would have appended to diffcf_e.dat (:688-694).
long wpadif(Orc* o, int S) {
  DIMS
  const double DTs = o->S("DTs");
  double* F2 = o->D("F2");
  const double *FNHS = o->D("FNHS"), *MU = o->D("MU"), *DMU = o->D("DMU"), *WMU = o->D("WMU");
  const int kind = A1(o->I("kind"), S);
  const double *DA, *DB;
  if (kind == 3) { DA = o->D("ATAW"); DB = o->D("ATAC"); }
  else { DA = o->D("ATAW_emic_h"); DB = o->D("ATAW_emic_he"); }
  std::vector<double> F(NPA, 0.0), RK(NPA, 0.0), RL(NPA, 0.0), FACMU(NPA, 0.0);
  long nviol = 0;
  for (int J = 1; J <= NT; ++J)
    for (int I = 2; I <= NR; ++I)
      for (int K = 2; K <= NE; ++K) {
        for (int L = 2; L <= NPA; ++L) {
          A1(FACMU, L) = FNHS_(I, J, L) * A1(MU, L);
          A1(F, L) = F2_(S, I, J, K, L) / A1(FACMU, L);
        }
        A1(FACMU, 1) = FNHS_(I, J, 1) * A1(MU, 1);
        A1(F, 1) = A1(F, 2);
        A1(RK, 1) = 0.;
        A1(RL, 1) = -1.;
        for (int L = 2; L <= NPA - 1; ++L) {
          double AN = (A4(DA, NR, NT, NE, I, J, K, L) + A4(DB, NR, NT, NE, I, J, K, L)) / A1(DMU, L);
          double GN = (A4(DA, NR, NT, NE, I, J, K, L - 1) + A4(DB, NR, NT, NE, I, J, K, L - 1)) / A1(DMU, L - 1);
          AN = AN * DTs / A1(FACMU, L) / A1(WMU, L);
          GN = GN * DTs / A1(FACMU, L) / A1(WMU, L);
          double BN = AN + GN;
          if (std::fabs(-1 - BN) < (std::fabs(AN) + std::fabs(GN))) ++nviol;
          double RP = A1(F, L);
          double DENOM = BN + GN * A1(RL, L - 1) + 1;
          A1(RK, L) = (RP + GN * A1(RK, L - 1)) / DENOM;
          A1(RL, L) = -AN / DENOM;
        }
        F2_(S, I, J, K, NPA - 1) = A1(RK, NPA - 1) / (1 + A1(RL, NPA - 1));
        for (int L = NPA - 2; L >= 1; --L) F2_(S, I, J, K, L) = A1(RK, L) - A1(RL, L) * F2_(S, I, J, K, L + 1);
        F2_(S, I, J, K, NPA) = F2_(S, I, J, K, NPA - 1);
        for (int L = 1; L <= NPA; ++L) F2_(S, I, J, K, L) = F2_(S, I, J, K, L) * A1(FACMU, L);
      }
  return nviol;
}

// FLCscatter  src/ModRamLoss.f90:513-575: field-line-curvature scattering, the same
// implicit pitch-angle diffusion with the single coefficient array FLC_coef(S,I,J,K,L)
// (here the (I,J,K,L) slab of species S, "FLC_coef").  Skipped during the first
// boundary cycle (TimeRamElapsed < Dt_bc, :523).  Returns the number of lines the
// reference would have logged to flc_cf_ion.dat (:548-555).
long flcscatter(Orc* o, int S) {
  DIMS
  const double DTs = o->S("DTs");
  if (o->S("T") < o->S("Dt_bc")) return 0;
  double* F2 = o->D("F2");
  const double *FNHS = o->D("FNHS"), *MU = o->D("MU"), *DMU = o->D("DMU"), *WMU = o->D("WMU"), *D = o->D("FLC_coef");
  std::vector<double> F(NPA, 0.0), RK(NPA, 0.0), RL(NPA, 0.0), FACMU(NPA, 0.0);
  long nviol = 0;
  for (int J = 1; J <= NT; ++J)
    for (int I = 2; I <= NR; ++I)
      for (int K = 2; K <= NE; ++K) {
        for (int L = 2; L <= NPA; ++L) {
          A1(FACMU, L) = FNHS_(I, J, L) * A1(MU, L);
          A1(F, L) = F2_(S, I, J, K, L) / A1(FACMU, L);
        }
        A1(FACMU, 1) = FNHS_(I, J, 1) * A1(MU, 1);
        A1(F, 1) = A1(F, 2);
        A1(RK, 1) = 0.;
        A1(RL, 1) = -1.;
        for (int L = 2; L <= NPA - 1; ++L) {
          double AN = A4(D, NR, NT, NE, I, J, K, L) / A1(DMU, L);
          double GN = A4(D, NR, NT, NE, I, J, K, L - 1) / A1(DMU, L - 1);
          AN = AN * DTs / A1(FACMU, L) / A1(WMU, L);
          GN = GN * DTs / A1(FACMU, L) / A1(WMU, L);
          double BN = AN + GN;
          if (std::fabs(-1 - BN) < (std::fabs(AN) + std::fabs(GN))) ++nviol;
          double RP = A1(F, L);
          double DENOM = BN + GN * A1(RL, L - 1) + 1;
          A1(RK, L) = (RP + GN * A1(RK, L - 1)) / DENOM;
          A1(RL, L) = -AN / DENOM;
        }
        F2_(S, I, J, K, NPA - 1) = A1(RK, NPA - 1) / (1 + A1(RL, NPA - 1));
        for (int L = NPA - 2; L >= 1; --L) F2_(S, I, J, K, L) = A1(RK, L) - A1(RL, L) * F2_(S, I, J, K, L + 1);
        F2_(S, I, J, K, NPA) = F2_(S, I, J, K, NPA - 1);
        for (int L = 1; L <= NPA; ++L) F2_(S, I, J, K, L) = F2_(S, I, J, K, L) * A1(FACMU, L);
      }
  return nviol;
}

// -----------------------------------------------------------------------------
// PARA_FLC  src/ModRamLoss.f90:342-455: the field-line-curvature pitch-angle diffusion
// coefficient FLC_coef(S,I,J,K,L) (Young 2002/2008, Ebihara 2011 fits) from the equatorial
// curvature radius and its zeta parameters r_curvEq, zeta1Eq, zeta2Eq (NR,NT) -- the
// output of FLC_Radius (:176-340) -- and BNES, BOUNHS.  Integer powers as the compiler
// evaluates them (x**(-n) = 1/(x*...*x)); everything else left to right.
void para_flc(Orc* o, int S) {
  DIMS
  const double Q = 1.602E-19, REarth = 6.4 * 1.E6;     // ModRamConst.f90:19, ModScbMain.f90:15
  const double *MU = o->D("MU"), *WMU = o->D("WMU"), *BOUNHS = o->D("BOUNHS"), *BNES = o->D("BNES"), *RMAS = o->D("RMAS"),
               *LZ = o->D("LZ"), *V = o->D("V"), *rc = o->D("r_curvEq"), *z1 = o->D("zeta1Eq"), *z2 = o->D("zeta2Eq");
  double* FLC = o->D("FLC_coef");
  for (size_t q = 0; q < (size_t)NR * NT * NE * NPA; ++q) FLC[q] = 0.0;
  std::vector<double> Nfactor(NPA), tau_bounce(NPA), D(NPA), Daa(NPA);
  for (int I = 1; I <= NR; ++I)
    for (int J = 1; J <= NT; ++J)
      for (int K = 1; K <= NE; ++K) {
        double Nmin = 1.0e20;
        int lmin = 1;
        for (int L = 1; L <= NPA; ++L) A1(Nfactor, L) = A1(tau_bounce, L) = A1(D, L) = A1(Daa, L) = 0.0;
        const double Vk = A2(V, nS, S, K);
        const double r_gyro = A1(RMAS, S) * Vk / std::fabs(BNES_(I, J) * Q);
        double epsl = r_gyro / A2(rc, NR, I, J);
        if (epsl > 0.584) epsl = 0.584;
        if (epsl >= 0.1) {
          const double e1 = 1.0 / epsl, e2 = 1.0 / (epsl * epsl), e3 = 1.0 / (epsl * epsl * epsl);
          const double a1 = -0.35533865 + 0.12800347 * e1 + 0.0017113113 * e2;
          const double a2 = 0.23156321 + 0.15561211 * e1 - 0.001860433 * e2;
          const double ba = -0.51057275 + 0.93651781 * e1 - 0.031690658 * e2;
          const double ca = 1.0663037 - 1.0944973 * e1 + 0.016679378 * e2 - 0.000499 * e3;
          const double da = -0.49667826 - 0.0081941799 * e1 + 0.0013621659 * e2;
          const double omegaa = 1.0513540 + 0.1351358 * epsl - 0.50787555 * (epsl * epsl);
          const double Am = std::exp(ca) * (std::pow(A2(z1, NR, I, J), a1) * std::pow(A2(z2, NR, I, J), a2) + da);
          for (int L = 1; L <= NPA - 1; ++L) {
            const double MUBOUN = A1(MU, L) + 0.5 * A1(WMU, L);
            const double alph = std::acos(MUBOUN);
            A1(Nfactor, L) = 1.0 / (std::sin(omegaa * alph) * std::pow(MUBOUN, ba));
            A1(tau_bounce, L) = 4 * A1(LZ, I) * REarth * BOUNHS_(I, J, L) / Vk;
            A1(D, L) = (Am * Am) / (2 * A1(tau_bounce, L));
            if (A1(Nfactor, L) <= Nmin) {
              Nmin = A1(Nfactor, L);
              lmin = L;
            }
          }
          for (int L = 1; L <= NPA - 1; ++L) {
            const double MUBOUN = A1(MU, L) + 0.5 * A1(WMU, L);
            const double alph = std::acos(MUBOUN);
            const double sn = std::sin(omegaa * alph);
            A1(Daa, L) = A1(D, L) * (A1(Nfactor, lmin) * A1(Nfactor, lmin)) * (sn * sn) * std::pow(MUBOUN, 2 * ba) /
                         ((1 - MUBOUN * MUBOUN) * (MUBOUN * MUBOUN));
            A4(FLC, NR, NT, NE, I, J, K, L) = A1(Daa, L) * (1 - MUBOUN * MUBOUN) * MUBOUN * BOUNHS_(I, J, L);
          }
        }
      }
}

// -----------------------------------------------------------------------------
// COULPARA  src/ModRamCoul.f90:17-125.  The plasmasphere species table is the
// reference's RAMSpecies(1:6) (src/ModRamSpecies.f90:42-133): mass, charge,
// plasmasphereRatio.  NOTE the reference never resets CCE/CDE/EDRE/CCI/CDI/EDRI
// inside the K loop: they accumulate across energies.  Reproduced.
void coulpara(Orc* o, int S) {
  DIMS
  const double Q = 1.602E-19, PI = 3.1415926535897932384626433832795, CS = 2.998E8, RE = 6.371E6, MP = 1.673E-27;
  const double DTs = o->S("DTs");
  const double *RMAS = o->D("RMAS"), *VBND = o->D("VBND"), *V = o->D("V"), *GREL = o->D("GREL"), *MU = o->D("MU"), *WMU = o->D("WMU"),
               *DMU = o->D("DMU"), *EKEV = o->D("EKEV"), *GRBND = o->D("GRBND");
  double *COULE = o->D("COULE"), *COULI = o->D("COULI"), *ATA = o->D("ATA"), *GTA = o->D("GTA"), *CEDR = o->D("CEDR"), *CIDR = o->D("CIDR");
  static const double ps_mass[6] = {5.4462E-4, 1.0, 4.0, 16.0, 14.0, 87.62};
  static const double ps_charge[6] = {-1, 1, 1, 1, 1, 1};
  static const double ps_ratio[6] = {1.0, 0.77, 0.2, 0.03, 0.0, 0.0};
  std::vector<double> COULDE(NPA, 0.0), COULDI(NPA, 0.0);
  double CCE = 0, CDE = 0, EDRE = 0, CCI = 0, CDI = 0, EDRI = 0;
  const double EPS = 8.854E-12, DLN = 21.5;
  const double Zt = (double)A1(o->I("QS"), S);
  const double QE = (Q * Q / EPS);
  const double GAMA = Zt * Zt * DLN / 4. / PI * QE * 1E6 * QE;
  const double CCO = GAMA / Q * DTs / Q / 1E3;
  const double CCD = GAMA * DTs / (A1(RMAS, S) * A1(RMAS, S)) / (CS * CS * CS);
  const double EDRCO = DLN * QE * RE / A1(RMAS, S) * QE * 1E9 / Q;
#define C3(a, S, K, L) A3(a, nS, NE, S, K, L)
  for (int k = 1; k <= NE; ++k) {
    for (int iS = 0; iS < 6; ++iS) {
      double RA = ps_ratio[iS];
      if (RA < 1e-9) continue;
      double VF = std::sqrt(2. * Q / (MP * ps_mass[iS]));
      double Zb = ps_charge[iS];
      double X = VBND_(S, k) / VF;
      double XD = V_(S, k) / VF;
      if (Zb < 0.0) {
        CCE = CCE + RA * Gcoul(X);
        CDE = CDE + RA * (std::erf(XD) - Gcoul(XD));
        EDRE = EDRE + RA * Gcoul(XD);
      } else {
        CCI = CCI + RA * (Zb * Zb) * Gcoul(X);
        CDI = CDI + RA * (Zb * Zb) * (std::erf(XD) - Gcoul(XD));
        EDRI = EDRI + RA * Gcoul(XD);
      }
    }
    C3(COULE, S, k, 1) = -CCE * VBND_(S, k) * CCO * (GRBND_(S, k) * GRBND_(S, k));
    C3(COULI, S, k, 1) = -CCI * VBND_(S, k) * CCO * (GRBND_(S, k) * GRBND_(S, k));
    C3(CEDR, S, k, 1) = EDRCO * A1(EKEV, k) * EDRE * (GREL_(S, k) + 1) / (GREL_(S, k) * GREL_(S, k)) / (V_(S, k) * V_(S, k));
    C3(CIDR, S, k, 1) = EDRCO * A1(EKEV, k) * EDRI * (GREL_(S, k) + 1) / (GREL_(S, k) * GREL_(S, k)) / (V_(S, k) * V_(S, k));
    double CCDE = CCD * CDE * GREL_(S, k) / std::pow(GREL_(S, k) * GREL_(S, k) - 1, 1.5);
    double CCDI = CCD * CDI * GREL_(S, k) / std::pow(GREL_(S, k) * GREL_(S, k) - 1, 1.5);
    for (int L = 2; L <= NPA - 1; ++L) {
      double BANE = (1. - FUNI(A1(MU, L)) / 2. / FUNT(A1(MU, L))) / (1. - A1(MU, L) * A1(MU, L));
      C3(COULE, S, k, L) = C3(COULE, S, k, 1);
      C3(COULI, S, k, L) = C3(COULI, S, k, 1);
      C3(CEDR, S, k, L) = C3(CEDR, S, k, 1) * BANE * FUNT(A1(MU, L)) * A1(MU, L) * A1(WMU, L);
      C3(CIDR, S, k, L) = C3(CIDR, S, k, 1) * BANE * FUNT(A1(MU, L)) * A1(MU, L) * A1(WMU, L);
      double MUBOUN = A1(MU, L) + 0.5 * A1(WMU, L);
      double BADIF = (1. - MUBOUN * MUBOUN) / MUBOUN / 2.;
      A1(COULDE, L) = CCDE * BADIF;
      double AFER = A1(COULDE, L) / A1(MU, L) / A1(DMU, L) / A1(WMU, L);
      double ASEC = A1(COULDE, L - 1) / A1(MU, L) / A1(DMU, L - 1) / A1(WMU, L);
      A1(COULDI, L) = CCDI * BADIF;
      double AFIR = A1(COULDI, L) / A1(MU, L) / A1(DMU, L) / A1(WMU, L);
      double ASIC = A1(COULDI, L - 1) / A1(MU, L) / A1(DMU, L - 1) / A1(WMU, L);
      C3(ATA, S, k, L) = AFIR + AFER;
      C3(GTA, S, k, L) = ASIC + ASEC;
    }
    C3(CEDR, S, k, NPA) = C3(CEDR, S, k, NPA - 1);
    C3(CIDR, S, k, NPA) = C3(CIDR, S, k, NPA - 1);
    C3(ATA, S, k, NPA) = 0;
  }
}

// COULEN  src/ModRamCoul.f90:133-221
void coulen(Orc* o, int S) {
  DIMS
  const double CS = 2.998E8, Q = 1.602E-19;
  const double BetaLim = o->S("BetaLim");
  double* F2 = o->D("F2");
  const double *EKEV = o->D("EKEV"), *WE = o->D("WE"), *DE = o->D("DE"), *RMAS = o->D("RMAS"), *NECR = o->D("NECR"), *GREL = o->D("GREL"),
               *COULE = o->D("COULE"), *COULI = o->D("COULI"), *MU = o->D("MU"), *FNHS = o->D("FNHS"), *FNIS = o->D("FNIS");
  std::vector<double> FBNDv(NE, 0.0), Fv(NE + 3, 0.0), CccolE(NE, 0.0), BANE(NPA, 0.0);
  double* FBND = FBNDv.data();
  double* F0 = Fv.data();
  const double EZERO = A1(EKEV, 1) - A1(WE, 1);
  const double GRZERO = 1. + EZERO * 1000. * Q / A1(RMAS, S) / CS / CS;
  F0[NE + 1] = 0.;
  F0[NE + 2] = 0.;
  for (int J = 1; J <= NT; ++J)
    for (int I = 2; I <= NR; ++I) {
      for (int L = 2; L <= NPA - 1; ++L) A1(BANE, L) = (1. - FNIS_(I, J, L) / 2. / FNHS_(I, J, L)) / (1. - A1(MU, L) * A1(MU, L));
      for (int L = NPA - 10; L <= NPA; ++L) A1(BANE, L) = A1(BANE, L - 1);
      for (int L = 2; L <= NPA; ++L) {
        double XNE = A2(NECR, NR, I, J) * A1(BANE, L);
        for (int K = 2; K <= NE; ++K) F0[K] = F2_(S, I, J, K, L);
        F0[1] = F0[2] * GREL_(S, 1) / GREL_(S, 2) * std::sqrt((sq(GREL_(S, 1)) - 1) / (sq(GREL_(S, 2)) - 1));
        F0[0] = F0[1] * GRZERO / GREL_(S, 1) * std::sqrt((sq(GRZERO) - 1) / (sq(GREL_(S, 1)) - 1));
        for (int K = 1; K <= NE; ++K) {
          A1(CccolE, K) = (C3(COULE, S, K, L) + C3(COULI, S, K, L)) * XNE;
          int ISIGN = 1;
          if (A1(CccolE, K) < 0.0) ISIGN = -1;
          int N = K + 1 - ISIGN;
          LIMITED_FLUX(A1(FBND, K), F0[K], F0[K + 1], F0[N], F0[N - 1], ISIGN, A1(CccolE, K) / A1(DE, K), BetaLim);
        }
        for (int K = 2; K <= NE; ++K) {
          F2_(S, I, J, K, L) = F2_(S, I, J, K, L) - A1(CccolE, K) / A1(WE, K) * A1(FBND, K) + A1(CccolE, K - 1) / A1(WE, K) * A1(FBND, K - 1);
          if (F2_(S, I, J, K, L) < 0.0) F2_(S, I, J, K, L) = 1E-15;
        }
      }
    }
}

// COULMU  src/ModRamCoul.f90:229-296  (T = TimeRamElapsed)
void coulmu(Orc* o, int S) {
  DIMS
  const double T = o->S("T");
  double* F2 = o->D("F2");
  const double *NECR = o->D("NECR"), *ATA = o->D("ATA"), *GTA = o->D("GTA"), *MU = o->D("MU"), *FNHS = o->D("FNHS"), *BOUNHS = o->D("BOUNHS"),
               *BOUNIS = o->D("BOUNIS");
  std::vector<double> RK(NPA, 0.0), RL(NPA, 0.0), BASCNE(NPA, 0.0);
  for (int J = 1; J <= NT; ++J)
    for (int I = 2; I <= NR; ++I) {
      double XNE = A2(NECR, NR, I, J);
      for (int K = 2; K <= NE; ++K) {
        A1(RK, 1) = 0.;
        A1(RL, 1) = -1.;
        A1(BASCNE, 1) = XNE * BOUNIS_(I, J, 1) / 2. / BOUNHS_(I, J, 1);
        for (int L = 2; L <= NPA - 1; ++L) {
          A1(BASCNE, L) = XNE * BOUNIS_(I, J, L) / 2. / BOUNHS_(I, J, L);
          double AN = C3(ATA, S, K, L) * A1(BASCNE, L) / FNHS_(I, J, L) * BOUNHS_(I, J, L);
          double GN = C3(GTA, S, K, L) * A1(BASCNE, L - 1) / FNHS_(I, J, L) * BOUNHS_(I, J, L - 1);
          double BN = AN + GN;
          double RP = F2_(S, I, J, K, L) / FNHS_(I, J, L) / A1(MU, L);
          double DENOM = BN + GN * A1(RL, L - 1) + 1;
          A1(RK, L) = (RP + GN * A1(RK, L - 1)) / DENOM;
          A1(RL, L) = -AN / DENOM;
        }
        F2_(S, I, J, K, NPA - 1) = A1(RK, NPA - 1) / (1 + A1(RL, NPA - 1));
        for (int L = NPA - 2; L >= 1; --L) F2_(S, I, J, K, L) = A1(RK, L) - A1(RL, L) * F2_(S, I, J, K, L + 1);
        F2_(S, I, J, K, NPA) = F2_(S, I, J, K, NPA - 1);
        for (int L = 1; L <= NPA; ++L) {
          F2_(S, I, J, K, L) = F2_(S, I, J, K, L) * FNHS_(I, J, L) * A1(MU, L);
          if ((T > 0.0) && (F2_(S, I, J, K, L) < 0.0)) F2_(S, I, J, K, L) = 1E-15;
        }
      }
    }
}

// -----------------------------------------------------------------------------
// SUMRC  src/ModRamRun.f90:231-259
void sumrc(Orc* o, int S) {
  DIMS
  const double *F2 = o->D("F2"), *EKEV = o->D("EKEV"), *WE = o->D("WE"), *WMU = o->D("WMU");
  double *SETRC = o->D("SETRC"), *ELORC = o->D("ELORC");
  A1(ELORC, S) = 0.;
  double ENOLD = A1(SETRC, S);
  A1(SETRC, S) = 0.;
  for (int I = 2; I <= NR; ++I)
    for (int K = 2; K <= NE; ++K)
      for (int L = 2; L <= NPA; ++L)
        for (int J = 1; J <= NT - 1; ++J) {
          double WEIGHT = F2_(S, I, J, K, L) * A1(WE, K) * A1(WMU, L);
          A1(SETRC, S) = A1(SETRC, S) + A1(EKEV, K) * WEIGHT;
        }
  A1(ELORC, S) = ENOLD - A1(SETRC, S);
}

// ANISCH moments  src/ModRamRun.f90:343-415 (pressure part; incl. the side
// effect F2(S,I,J,K,1)=F2(S,I,J,K,2) at :366).  khi(5) band edges are inputs.
void anisch(Orc* o, int S) {
  DIMS
  const double CS = 2.998E8, PI = 3.1415926535897932384626433832795;
  double* F2 = o->D("F2");
  const double *UPA = o->D("UPA"), *WMU = o->D("WMU"), *FFACTOR = o->D("FFACTOR"), *MU = o->D("MU"), *EKEV = o->D("EKEV"),
               *EPP = o->D("EPP"), *ERNH = o->D("ERNH"), *FNHS = o->D("FNHS");
  double *PPERT = o->D("PPERT"), *PPART = o->D("PPART");
  const int* khi = o->I("khi");
  const double cv = CS * 100;
  const double RFAC = 4 * PI / cv;
  for (int I = 2; I <= NR; ++I)
    for (int J = 1; J <= NT; ++J) {
      int klo = 2;
      A3(PPERT, nS, NR, S, I, J) = 0.;
      A3(PPART, nS, NR, S, I, J) = 0.;
      for (int iwa = 1; iwa <= 5; ++iwa) {
        double PPER = 0., PPAR = 0., RNHT = 0., EDEN = 0.;
        for (int K = klo; K <= A1(khi, iwa); ++K) {
          F2_(S, I, J, K, 1) = F2_(S, I, J, K, 2);
          double SUME = 0., SUMA = 0., SUMN = 0.;
          int u = (int)(A1(UPA, I) - 1);
          for (int L = 1; L <= u; ++L) {
            double ERNM = A1(WMU, L) / A4(FFACTOR, nS, NR, NE, S, I, K, L) / FNHS_(I, J, L);
            double EPMA = ERNM * A1(MU, L) * A1(MU, L);
            double EPME = ERNM - EPMA;
            SUME = SUME + F2_(S, I, J, K, L) * EPME;
            SUMA = SUMA + F2_(S, I, J, K, L) * EPMA;
            SUMN = SUMN + F2_(S, I, J, K, L) * ERNM;
          }
          PPER = PPER + A2(EPP, nS, S, K) * SUME;
          PPAR = PPAR + A2(EPP, nS, S, K) * SUMA;
          RNHT = RNHT + A2(ERNH, nS, S, K) * SUMN;
          EDEN = EDEN + A2(ERNH, nS, S, K) * A1(EKEV, K) * SUMN;
        }
        PPAR = 2 * RFAC * PPAR;
        PPER = RFAC * PPER;
        klo = A1(khi, iwa) + 1;
        A3(PPERT, nS, NR, S, I, J) = A3(PPERT, nS, NR, S, I, J) + PPER;
        A3(PPART, nS, NR, S, I, J) = A3(PPART, nS, NR, S, I, J) + PPAR;
        (void)RNHT; (void)EDEN;
      }
    }
}

// ---- GSL pieces of the diffusion-coefficient rebuild (GNU GSL 2.5 / 2.6, un-vendored: restated from the published
// algorithms, like the Steffen spline of oracle/scb_oracle.cpp) ----------------------------------------------------------
// GSL_Interpolation_1D -> Interpolation_1D_array (src/ModRamGSL.f90:240-311) + interpolation_1d_c (src/RamGSL.c:111-174):
// the wrapper drops abscissae that do not increase, the C driver extrapolates linearly outside [xa_0, xa_n-1] (end points
// included) and evaluates the Steffen cubic (gsl interpolation/steffen.c) of the bisection interval inside.
inline double copysign1(double y) { return (y < 0.0) ? -1.0 : 1.0; }   // steffen_copysign(1.0, y) of gsl steffen.c
int interp1d_steffen(int n0, const double* x1, const double* f1, int n2, const double* x2, double* f2) {
  std::vector<double> xa(n0), fa(n0);
  int n1 = 1;
  xa[0] = x1[0];
  fa[0] = f1[0];
  for (int i = 1; i < n0; ++i)
    if (x1[i] > xa[n1 - 1]) { xa[n1] = x1[i]; fa[n1] = f1[i]; ++n1; }
  if (n1 < 3) return 1;
  std::vector<double> yp(n1), a(n1 - 1), b(n1 - 1);
  yp[0] = (fa[1] - fa[0]) / (xa[1] - xa[0]);
  for (int i = 1; i < n1 - 1; ++i) {
    const double hi = xa[i + 1] - xa[i], him1 = xa[i] - xa[i - 1];
    const double si = (fa[i + 1] - fa[i]) / hi, sim1 = (fa[i] - fa[i - 1]) / him1;
    const double pi = (sim1 * hi + si * him1) / (him1 + hi);
    const double m1 = std::fabs(si) < 0.5 * std::fabs(pi) ? std::fabs(si) : 0.5 * std::fabs(pi);
    const double m2 = std::fabs(sim1) < m1 ? std::fabs(sim1) : m1;
    yp[i] = (copysign1(sim1) + copysign1(si)) * m2;
  }
  yp[n1 - 1] = (fa[n1 - 1] - fa[n1 - 2]) / (xa[n1 - 1] - xa[n1 - 2]);
  for (int i = 0; i < n1 - 1; ++i) {
    const double hi = xa[i + 1] - xa[i];
    const double si = (fa[i + 1] - fa[i]) / hi;
    a[i] = (yp[i] + yp[i + 1] - 2 * si) / hi / hi;
    b[i] = (3 * si - 2 * yp[i] - yp[i + 1]) / hi;
  }
  for (int q = 0; q < n2; ++q) {
    const double xb = x2[q];
    if (xb <= xa[0]) f2[q] = fa[0] + (xb - xa[0]) / (xa[1] - xa[0]) * (fa[1] - fa[0]);
    else if (xb >= xa[n1 - 1]) f2[q] = fa[n1 - 1] + (xb - xa[n1 - 1]) / (xa[n1 - 2] - xa[n1 - 1]) * (fa[n1 - 2] - fa[n1 - 1]);
    else if (xb == xb) {
      int ilo = 0, ihi = n1 - 1;
      while (ihi > ilo + 1) {
        const int i = (ihi + ilo) / 2;
        if (xa[i] > xb) ihi = i; else ilo = i;
      }
      const double delx = xb - xa[ilo];
      f2[q] = fa[ilo] + delx * (yp[ilo] + delx * (b[ilo] + delx * a[ilo]));
    } else return 1;
  }
  return 0;
}
// GSL_Interpolation_2D -> Interpolation_2D_point (src/ModRamGSL.f90:489-533) + interpolation_2d_c (src/RamGSL.c:178-214):
// gsl_interp2d_bilinear evaluated with gsl_interp2d_eval_extrap (no domain check: the edge cell is used outside the
// table).  Index search = gsl_interp_bsearch(xa, x, 0, n-1): xa[i] <= x < xa[i+1], clipped to [0, n-2].
inline int gsl_bsearch(const double* xa, double x, int n) {
  int ilo = 0, ihi = n - 1;
  while (ihi > ilo + 1) {
    const int i = (ihi + ilo) / 2;
    if (xa[i] > x) ihi = i; else ilo = i;
  }
  return ilo;
}
inline double interp2d_bilinear(int n1, int m1, const double* xa, const double* ya, const double* za, double x, double y) {
  const int xi = gsl_bsearch(xa, x, n1), yi = gsl_bsearch(ya, y, m1);
  const double xmin = xa[xi], xmax = xa[xi + 1], ymin = ya[yi], ymax = ya[yi + 1];
  const double zminmin = za[(size_t)yi * n1 + xi], zminmax = za[(size_t)(yi + 1) * n1 + xi];
  const double zmaxmin = za[(size_t)yi * n1 + xi + 1], zmaxmax = za[(size_t)(yi + 1) * n1 + xi + 1];
  const double dx = xmax - xmin, dy = ymax - ymin;
  const double t = (x - xmin) / dx, u = (y - ymin) / dy;
  return (1. - t) * (1. - u) * zminmin + t * (1. - u) * zmaxmin + (1. - t) * u * zminmax + t * u * zmaxmax;
}

// ANISCH, second half: the pitch-angle diffusion coefficients of WPADIF, rebuilt every Dt_bc
// (src/ModRamRun.f90:422-605).  Electrons with DoUseWPI: chorus outside the plasmapause (XNE <= 50; Steffen interpolation
// of log10 <Daa> from the table's pitch angles PA onto PAbn -> ATAC) and hiss inside (bilinear interpolation of the
// normalised table in (log10 E, fpe/fce) -> ATAW).  Species with EMIC (H+) and DoUseEMIC: bilinear interpolation of the
// H-band / He-band tables, scaled by the wave intensity of I_emic (src/ModRamWPI.f90:720-750) -> ATAW_emic_h / _he.
// The caller keeps the MOD(INT(T), INT(Dt_bc)) == 0 gate.  Returns the number of failed 1-D interpolations.
int anisch_diffcoef(Orc* o, int S, int flags) {
  DIMS
  const double CS = 2.998E8, PI = 3.1415926535897932384626433832795, Q = 1.602E-19;
  const int kind = A1(o->I("kind"), S);
  const bool DoUseWPI = flags & 1, DoUseEMIC = flags & 4;
  const double *MU = o->D("MU"), *WMU = o->D("WMU"), *EKEV = o->D("EKEV"), *RMAS = o->D("RMAS"), *BNES = o->D("BNES"),
               *BOUNHS = o->D("BOUNHS"), *GREL = o->D("GREL"), *XNE = o->D("XNE"), *PAbn = o->D("PAbn");
  const double cv = CS * 100, esu = Q * 3E9, gausgam = 1.E-5;
  int nerr = 0;
  if (DoUseWPI && kind == 3) {
    double *ATAW = o->D("ATAW"), *ATAC = o->D("ATAC");
    const double *CDAAR = o->D("CDAAR"), *BDAAR = o->D("BDAAR"), *NDAAJ = o->D("NDAAJ"), *ENOR = o->D("ENOR"), *fpofc = o->D("fpofc");
    const int ENG = (int)o->S("ENG"), NCF = (int)o->S("NCF");
    const bool DoUseBASdiff = o->S("DoUseBASdiff") != 0.0;
    for (size_t q = 0; q < (size_t)NR * NT * NE * NPA; ++q) { ATAW[q] = 0.0; ATAC[q] = 0.0; }
    std::vector<double> PA(NPA), DAMR1(NPA), Y(NPA);
    for (int L = 1; L <= NPA; ++L) A1(PA, L) = 180.0 / PI * std::acos(A1(MU, NPA - L + 1));    // ACOSD, src/ModRamFunctions.f90:172
    for (int I = 2; I <= NR; ++I)
      for (int J = 1; J <= NT; ++J)
        if (A2(XNE, NR, I, J) <= 50.) {                                                         // outside the plasmapause
          const double fnorm = 1;
          for (int K = 2; K <= NE; ++K) {
            for (int L = 1; L <= NPA; ++L)
              A1(DAMR1, L) = DoUseBASdiff ? std::log10(CD4(CDAAR, I, J, K, NPA - L + 1)) : std::log10(CD4(BDAAR, I, J, K, L));
            // the reference interpolates one target per call; the spline of a line is the same for all of them
            if (interp1d_steffen(NPA, PA.data(), DAMR1.data(), NPA, PAbn, Y.data())) ++nerr;
            for (int L = 1; L <= NPA; ++L) {
              const double MUBOUN = A1(MU, L) + 0.5 * A1(WMU, L);
              double taudaa = std::pow(10., A1(Y, L)) * fnorm;
              if (taudaa > 1e0) taudaa = 1e-1;
              if (taudaa < 1e-30) taudaa = 1e-30;
              CD4(ATAC, I, J, K, L) = taudaa * (1. - MUBOUN * MUBOUN) * MUBOUN * BOUNHS_(I, J, L);
            }
          }
        }
    double Bw = 30.;
    if (o->S("Kp") >= 4.0) Bw = 100.;
    std::vector<double> ALENOR(ENG), DUMP((size_t)ENG * NCF);
    for (int KN = 1; KN <= ENG; ++KN) A1(ALENOR, KN) = std::log10(A1(ENOR, KN));
    for (int I = 2; I <= NR; ++I)
      for (int J = 1; J <= NT; ++J)
        if (A2(XNE, NR, I, J) > 50.) {                                                          // inside the plasmapause
          const double omega = esu * 10 * BNES_(I, J) / (A1(RMAS, S) * cv);
          double xfrl = CS * std::sqrt(A2(XNE, NR, I, J) * A1(RMAS, S) * 40 * PI) / 10. / BNES_(I, J);
          if (xfrl > 18) xfrl = 18.;
          if (xfrl < 2) xfrl = 2.;
          const double fnorm = omega * ((Bw * 1e-3) * (Bw * 1e-3)) * (gausgam * gausgam) / 1e8 / BNES_(I, J) / BNES_(I, J);
          for (int L = 1; L <= NPA; ++L) {
            const double MUBOUN = A1(MU, L) + 0.5 * A1(WMU, L);
            for (int IZ = 1; IZ <= NCF; ++IZ)
              for (int KN = 1; KN <= ENG; ++KN) A2(DUMP, ENG, KN, IZ) = std::log10(A4(NDAAJ, NR, ENG, NPA, I, KN, L, IZ));
            for (int K = 2; K <= NE; ++K) {
              const double ER1 = std::log10(A1(EKEV, K));
              const double Yv = interp2d_bilinear(ENG, NCF, ALENOR.data(), fpofc, DUMP.data(), ER1, xfrl);
              CD4(ATAW, I, J, K, L) = std::pow(10., Yv) * fnorm / (GREL_(S, K) * GREL_(S, K)) * (1. - MUBOUN * MUBOUN) / MUBOUN;
            }
          }
        }
  }
  if (DoUseEMIC && kind == 0) {
    double *AH = o->D("ATAW_emic_h"), *AHE = o->D("ATAW_emic_he");
    const double *DH = o->D("Daa_emic_h"), *DHE = o->D("Daa_emic_he"), *EKEV_emic = o->D("EKEV_emic"), *fp2c = o->D("fp2c_emic"),
                 *Ihs = o->D("Ihs_emic"), *Ihes = o->D("Ihes_emic");
    const int ENGe = (int)o->S("ENG_emic"), NCFe = (int)o->S("NCF_emic");
    const int AE = (int)o->S("AE");
    const int eleS = (int)o->S("electron_species");        // RMAS(4) in the reference: the electron's index
    for (size_t q = 0; q < (size_t)NR * NT * NE * NPA; ++q) { AH[q] = 0.0; AHE[q] = 0.0; }
    std::vector<double> logE(ENGe), D1((size_t)ENGe * NCFe), D2((size_t)ENGe * NCFe);
    for (int KN = 1; KN <= ENGe; ++KN) A1(logE, KN) = std::log10(A1(EKEV_emic, KN));
    int cls = 0;                                            // I_emic: AE class 1..4 (0: none, intensities stay 0)
    if (AE >= 0 && AE < 100) cls = 1; else if (AE >= 100 && AE < 300) cls = 2; else if (AE >= 300 && AE < 400) cls = 3; else if (AE >= 400) cls = 4;
    for (int I = 2; I <= NR; ++I)
      for (int J = 1; J <= NT; ++J) {
        double xfrl = CS * std::sqrt(A2(XNE, NR, I, J) * A1(RMAS, eleS) * 40 * PI) / 10. / BNES_(I, J);
        if (xfrl > 20) xfrl = 20.;
        if (xfrl < 2) xfrl = 2.;
        const double fnorm_h = cls ? A3(Ihs, 4, NR, cls, I, J) : 0.0, fnorm_he = cls ? A3(Ihes, 4, NR, cls, I, J) : 0.0;
        for (int L = 1; L <= NPA; ++L) {
          const double MUBOUN = A1(MU, L) + 0.5 * A1(WMU, L);
          for (int IZ = 1; IZ <= NCFe; ++IZ)
            for (int KN = 1; KN <= ENGe; ++KN) {
              A2(D1, ENGe, KN, IZ) = std::log10(A4(DH, NR, ENGe, NPA, I, KN, L, IZ));
              A2(D2, ENGe, KN, IZ) = std::log10(A4(DHE, NR, ENGe, NPA, I, KN, L, IZ));
            }
          for (int K = 2; K <= NE; ++K) {
            const double ER1 = std::log10(A1(EKEV, K));
            double Yv = interp2d_bilinear(ENGe, NCFe, logE.data(), fp2c, D1.data(), ER1, xfrl);
            double vh = std::pow(10., Yv) * fnorm_h * (1. - MUBOUN * MUBOUN) * MUBOUN * BOUNHS_(I, J, L);
            Yv = interp2d_bilinear(ENGe, NCFe, logE.data(), fp2c, D2.data(), ER1, xfrl);
            double vhe = std::pow(10., Yv) * fnorm_he * (1. - MUBOUN * MUBOUN) * MUBOUN * BOUNHS_(I, J, L);
            if (vh <= 1.0e-20) vh = 1.0e-31;
            if (vhe <= 1.0e-20) vhe = 1.0e-31;
            CD4(AH, I, J, K, L) = vh;
            CD4(AHE, I, J, K, L) = vhe;
          }
        }
      }
  }
  return nerr;
}

// GEOSB (src/ModRamBoundary.f90:241-319), boundary == 'LANL': the outer-boundary distribution of species S from the
// interpolated geosynchronous flux FluxLanl(NT,NE) [1/cm2/s/sr/keV, isotropic] (get_geomlt_flux: file I/O, host), the
// composition factor species%s_comp and FFACTOR at the outermost radius, for pitch angles 2 .. UPA(NR)-1.
void geosb(Orc* o, int S) {
  DIMS
  double *FGEOS = o->D("FGEOS"), *Flux = o->D("FluxLanl");
  const double *FFACTOR = o->D("FFACTOR"), *UPA = o->D("UPA");
  const double comp = o->S("s_comp");
  for (int L = 1; L <= NPA; ++L)
    for (int K = 1; K <= NE; ++K)
      for (int J = 1; J <= NT; ++J) A4(FGEOS, nS, NT, NE, S, J, K, L) = 0.;
  for (int K = 1; K <= NE; ++K) A2(Flux, NT, 1, K) = A2(Flux, NT, NT, K);             // FluxLanl(1,:) = FluxLanl(nT,:)
  for (int K = 1; K <= NE; ++K)
    for (int J = 1; J <= NT; ++J) A2(Flux, NT, J, K) = A2(Flux, NT, J, K) * comp;     // FluxLanl = FluxLanl*s_comp
  const int u = (int)(A1(UPA, NR) - 1);
  for (int K = 1; K <= NE; ++K)
    for (int J = 1; J <= NT; ++J)
      for (int L = 2; L <= u; ++L) A4(FGEOS, nS, NT, NE, S, J, K, L) = A2(Flux, NT, J, K) * A4(FFACTOR, nS, NR, NE, S, NR, K, L);
}

// get_electric_field (src/ModRamEField.f90:14-63): VT(NR+1,NT) either interpolated in time between two potential maps
// (electric /= 'VOLS': VT = VTOL + (VTN - VTOL)*(TimeRamElapsed - TOLV)/DtEfi) or the Volland-Stern potential.
void get_electric_field(Orc* o, int vols) {
  DIMS
  double* VT = o->D("VT");
  const double RE = 6.371E6;
  if (!vols) {
    const double *VTOL = o->D("VTOL"), *VTN = o->D("VTN");
    const double t = o->S("TimeRamElapsed"), TOLV = o->S("TOLV"), DtEfi = o->S("DtEfi");
    for (int I = 1; I <= NR + 1; ++I)
      for (int J = 1; J <= NT; ++J) VT_(I, J) = A2(VTOL, NR + 1, I, J) + (A2(VTN, NR + 1, I, J) - A2(VTOL, NR + 1, I, J)) * (t - TOLV) / DtEfi;
  } else {
    const double KP = o->S("Kp"), PHIOFS = o->S("PHIOFS");
    const double *LZ = o->D("LZ"), *PHI = o->D("PHI");
    for (int I = 1; I <= NR + 1; ++I)
      for (int J = 1; J <= NT; ++J) {
        const double AVS = 7.05E-6 / ((1. - 0.159 * KP + 0.0093 * (KP * KP)) * (1. - 0.159 * KP + 0.0093 * (KP * KP)) * (1. - 0.159 * KP + 0.0093 * (KP * KP))) / RE;
        VT_(I, J) = AVS * ((A1(LZ, I) * RE) * (A1(LZ, I) * RE)) * std::sin(A1(PHI, J) - PHIOFS);
      }
  }
}

// flags for ram_run
enum { F_WPI = 1, F_COULOMB = 2, F_EMIC = 4 };

// the species-loop body of ram_run, src/ModRamRun.f90:64-185
void ram_species(Orc* o, int iS, int flags) {
  double *ELORC = o->D("ELORC"), *LSDR = o->D("LSDR"), *LSCHA = o->D("LSCHA"), *LSATM = o->D("LSATM"), *LSWAE = o->D("LSWAE"),
         *LSCOE = o->D("LSCOE"), *LSCSC = o->D("LSCSC");
  const int kind = A1(o->I("kind"), iS);
  const bool sWPI = (kind == 3), sCEX = (kind != 3), sEMIC = (kind == 0);
  const bool DoUseWPI = flags & F_WPI, DoUseCoulomb = flags & F_COULOMB, DoUseEMIC = flags & F_EMIC;
  cepara(o, iS);
  driftpara(o, iS);
  if (DoUseCoulomb) coulpara(o, iS);
  driftr(o, iS); driftp(o, iS); drifte(o, iS); driftmu(o, iS);
  sumrc(o, iS); A1(LSDR, iS) += A1(ELORC, iS);
  if (DoUseCoulomb) {
    coulen(o, iS); sumrc(o, iS); A1(LSCOE, iS) += A1(ELORC, iS);
    coulmu(o, iS); sumrc(o, iS); A1(LSCSC, iS) += A1(ELORC, iS);
  }
  if (sWPI) {
    if (DoUseWPI) wpadif(o, iS); else wavelo(o, iS);
    sumrc(o, iS); A1(LSWAE, iS) += A1(ELORC, iS);
  }
  if (sEMIC && DoUseEMIC) { wpadif(o, iS); sumrc(o, iS); A1(LSWAE, iS) += A1(ELORC, iS); }
  if (sCEX) { charexchange(o, iS); sumrc(o, iS); A1(LSCHA, iS) += A1(ELORC, iS); }
  atmol(o, iS); sumrc(o, iS); A1(LSATM, iS) += A1(ELORC, iS);
  // time splitting, reverse order
  atmol(o, iS); sumrc(o, iS); A1(LSATM, iS) += A1(ELORC, iS);
  if (sCEX) { charexchange(o, iS); sumrc(o, iS); A1(LSCHA, iS) += A1(ELORC, iS); }
  if (sEMIC && DoUseEMIC) { wpadif(o, iS); sumrc(o, iS); A1(LSWAE, iS) += A1(ELORC, iS); }
  if (sWPI) {
    if (DoUseWPI) wpadif(o, iS); else wavelo(o, iS);
    sumrc(o, iS); A1(LSWAE, iS) += A1(ELORC, iS);
  }
  if (DoUseCoulomb) {
    coulmu(o, iS); sumrc(o, iS); A1(LSCSC, iS) += A1(ELORC, iS);
    coulen(o, iS); sumrc(o, iS); A1(LSCOE, iS) += A1(ELORC, iS);
  }
  driftmu(o, iS); drifte(o, iS); driftp(o, iS); driftr(o, iS);
  sumrc(o, iS); A1(LSDR, iS) += A1(ELORC, iS);
}

// ram_run after the Volland-Stern block, src/ModRamRun.f90:64-222; returns DtsNext
double ram_run(Orc* o, int flags, int nthreads) {
  DIMS
  if (nthreads < 1) nthreads = 1;
  // make sure per-species work arrays exist before entering the parallel region
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
  for (int iS = 1; iS <= nS; ++iS) ram_species(o, iS, flags);
  double* F2 = o->D("F2");
  const int* outsideMGNP = o->I("outsideMGNP");
  for (int L = 1; L <= NPA; ++L)
    for (int K = 1; K <= NE; ++K)
      for (int I = 1; I <= NR; ++I)
        for (int S = 1; S <= nS; ++S) F2_(S, I, NT, K, L) = F2_(S, I, 1, K, L);
  for (int I = 1; I <= NR; ++I)
    for (int J = 1; J <= NT; ++J)
      if (OUT_(I, J) == 1)
        for (int L = 1; L <= NPA; ++L)
          for (int K = 1; K <= NE; ++K)
            for (int S = 1; S <= nS; ++S) F2_(S, I, J, K, L) = 1.e-31;
  const double *DtR = o->D("DtDriftR"), *DtP = o->D("DtDriftP"), *DtE = o->D("DtDriftE"), *DtM = o->D("DtDriftMu");
  double DtsNext = 1e300;
  for (int S = 0; S < nS; ++S) DtsNext = std::min(DtsNext, std::min(std::min(DtR[S], DtP[S]), std::min(DtE[S], DtM[S])));
  DtsNext = std::max(DtsNext, o->S("DtsMin"));
  // pressure totals + FLUX  (:208-222)
  double* FLUX = o->D("FLUX");
  const double *FFACTOR = o->D("FFACTOR"), *FNHS = o->D("FNHS");
  for (int iS = 1; iS <= nS; ++iS) {
    anisch(o, iS);
    for (int I = 2; I <= NR; ++I)
      for (int K = 2; K <= NE; ++K)
        for (int L = 2; L <= NPA; ++L)
          for (int J = 1; J <= NT - 1; ++J)
            A5(FLUX, nS, NR, NT, NE, iS, I, J, K, L) = F2_(iS, I, J, K, L) / A4(FFACTOR, nS, NR, NE, iS, I, K, L) / FNHS_(I, J, L);
  }
  return DtsNext;
}

}  // namespace

// =============================================================================
// C interface (ctypes)
// =============================================================================
extern "C" {

void* orc_create(int nS, int NR, int NT, int NE, int NPA) {
  Orc* o = new Orc();
  o->nS = nS; o->NR = NR; o->NT = NT; o->NE = NE; o->NPA = NPA;
  o->w.resize(nS);
  o->s["BetaLim"] = 1.5;   // src/ModRamParams.f90:86
  o->s["FracCFL"] = 0.8;   // src/ModRamVariables.f90:100
  o->s["DtsMin"] = 1.0;    // src/ModRamTiming.f90
  o->s["T"] = 0.0;
  return o;
}
void orc_destroy(void* h) { delete (Orc*)h; }
void orc_set_array(void* h, const char* name, double* p) { ((Orc*)h)->d[name] = p; }
void orc_set_iarray(void* h, const char* name, int* p) { ((Orc*)h)->i[name] = p; }
void orc_set_scalar(void* h, const char* name, double v) { ((Orc*)h)->s[name] = v; }
double orc_get_scalar(void* h, const char* name) { return ((Orc*)h)->S(name); }

void orc_driftpara(void* h, int S) { driftpara((Orc*)h, S); }
void orc_driftr(void* h, int S) { driftr((Orc*)h, S); }
void orc_driftp(void* h, int S) { driftp((Orc*)h, S); }
void orc_drifte(void* h, int S) { drifte((Orc*)h, S); }
void orc_driftmu(void* h, int S) { driftmu((Orc*)h, S); }
void orc_cepara(void* h, int S) { cepara((Orc*)h, S); }
void orc_charexchange(void* h, int S) { charexchange((Orc*)h, S); }
void orc_atmol(void* h, int S) { atmol((Orc*)h, S); }
void orc_wavelo(void* h, int S) { wavelo((Orc*)h, S); }
long orc_wpadif(void* h, int S) { return wpadif((Orc*)h, S); }
long orc_flcscatter(void* h, int S) { return flcscatter((Orc*)h, S); }
void orc_para_flc(void* h, int S) { para_flc((Orc*)h, S); }
void orc_coulpara(void* h, int S) { coulpara((Orc*)h, S); }
void orc_coulen(void* h, int S) { coulen((Orc*)h, S); }
void orc_coulmu(void* h, int S) { coulmu((Orc*)h, S); }
void orc_sumrc(void* h, int S) { sumrc((Orc*)h, S); }
void orc_anisch(void* h, int S) { anisch((Orc*)h, S); }
double orc_ram_run(void* h, int flags, int nthreads) { return ram_run((Orc*)h, flags, nthreads); }
int orc_anisch_diffcoef(void* h, int S, int flags) { return anisch_diffcoef((Orc*)h, S, flags); }
void orc_geosb(void* h, int S) { geosb((Orc*)h, S); }
void orc_get_electric_field(void* h, int vols) { get_electric_field((Orc*)h, vols); }
// copy of a species' drift coefficient array (NR,NT,NE,NPA), which: 0=R 1=P 2=E 3=Mu
void orc_get_cdrift(void* h, int S, int which, double* out) {
  Orc* o = (Orc*)h;
  SpeciesWork& w = o->w[S - 1];
  const std::vector<double>& v = which == 0 ? w.CDriftR : which == 1 ? w.CDriftP : which == 2 ? w.CDriftE : w.CDriftMu;
  std::memcpy(out, v.data(), v.size() * sizeof(double));
}
double orc_gcoul(double x) { return Gcoul(x); }
double orc_funt(double x) { return FUNT(x); }
double orc_funi(double x) { return FUNI(x); }
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}
