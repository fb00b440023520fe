// Host side of the SCB C ABI (include/ramscb_gpu.h, rsg_scb_*): device mirrors of
// the ModScbVariables arrays and kernel launches.  No CPU compute path: without
// a CUDA device every entry point fails with RSG_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ramscb_gpu.h"
#include "scb_kernels.cuh"
#include "scb_press_front.cuh"
#include "hi_kernels.cuh"

namespace {

thread_local std::string g_serr;
int sfail(int code, const std::string& msg) {
  g_serr = msg;
  return code;
}
#define SCK(call)                                                                                         \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return sfail(RSG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                     std::to_string(__LINE__) + ")");                                     \
  } while (0)
#define SCKL()                                                                                            \
  do {                                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                                  \
    if (e_ != cudaSuccess)                                                                                \
      return sfail(RSG_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                     std::to_string(__LINE__) + ")");                                     \
  } while (0)
#define SRET(x)                  \
  do {                           \
    int r_ = (x);                \
    if (r_ != RSG_OK) return r_; \
  } while (0)

const double PI_D = 3.141592653589793238462643383279502884197;
inline int nblk(long long n, int b) { return (int)((n + b - 1) / b); }

}  // namespace

extern "C" const char* rsg_scb_last_error(void) { return g_serr.c_str(); }

struct rsg_scb {
  int nthe, npsi, nzeta, device = 0;
  int isotropy = 0;
  ScbDev dev{};
  std::map<std::string, std::pair<double*, size_t>> arr;   // name -> (device ptr, elements)
  std::vector<void*> allocs;
  cudaStream_t st = nullptr, own_st = nullptr;   // st: where the work goes (own_st unless rsg_scb_set_stream)
  double *d_prev = nullptr, *d_part = nullptr, *d_resmax = nullptr;
  int *d_ni = nullptr, *d_fail = nullptr;
  size_t npart = 0;
  long long launches = 0;
  double bnormal, pnormal, pjconst;
  bool grid_set = false, geom_set = false, press_set = false, band_done = false;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double last_ms = 0.0;
  double *d_alphaVal = nullptr, *d_psiVal = nullptr, *d_chiVal = nullptr, *d_mapw = nullptr;   // map*: targets, workspace
  bool map_set = false;
  double *d_peq = nullptr, *d_tau = nullptr;   // pressure_aniso: equatorial inputs (2 x npsi x (nzeta+1)), tau
  // `pressure` front end on the device (rsg_scb_set_ram_pressure): extended + smoothed RAM pressures and their axes
  double *d_pf = nullptr, *d_rad2 = nullptr, *d_azim = nullptr, *d_rper = nullptr, *d_rpar = nullptr;
  int pf_nX = 0, pf_nAz = 0;
  bool pf_set = false;
  std::map<std::string, double*> snaps;   // "name#slot" -> device copy (rsg_scb_snapshot)
  // iterateAlpha sharded along zeta (rsg_scb_zsolve_*): per-surface state of the open solve
  double* d_zstate = nullptr;
  int *d_zdone = nullptr, *d_zpend = nullptr;
  ZArgs z{};
  int z_sweep = 0;
  bool z_open = false;
  bool use_cluster = true;   // 4-colour SOR on thread-block clusters with the problem resident on chip
  int last_cluster = 0;      // cluster size of the last SOR launch (0: one CTA per sub-problem)

  int dalloc(double** p, size_t n, const char* name) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(double));
    if (e != cudaSuccess) return sfail(RSG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cudaMemset(q, 0, n * sizeof(double));
    allocs.push_back(q);
    *p = (double*)q;
    if (name) arr[name] = {(double*)q, n};
    return RSG_OK;
  }
};

extern "C" {

int rsg_scb_create(rsg_scb** out, int nthe, int npsi, int nzeta, int device) {
  if (!out) return sfail(RSG_ERR_ARG, "null out");
  if (nthe < 12 || npsi < 6 || nzeta < 6) return sfail(RSG_ERR_ARG, "bad dimensions");
  int ndev = 0;
  SCK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return sfail(RSG_ERR_CUDA, "no CUDA device");
  if (device >= 0) SCK(cudaSetDevice(device));
  rsg_scb* h = new rsg_scb();
  SCK(cudaGetDevice(&h->device));
  h->nthe = nthe; h->npsi = npsi; h->nzeta = nzeta;
  ScbDev& d = h->dev;
  d.nthe = nthe; d.npsi = npsi; d.nzeta = nzeta;
  // src/ModScbInit.f90:131-148
  d.dr = 1.0 / (double)(npsi - 1);
  d.dt = PI_D / (double)(nthe - 1);
  d.dpPrime = 2 * PI_D / (double)(nzeta - 1);
  d.rdr = 1.0 / d.dr; d.rdt = 1.0 / d.dt; d.rdp = 1.0 / d.dpPrime;
  d.rdrsq = d.rdr * d.rdr; d.rdtsq = d.rdt * d.rdt; d.rdpsq = d.rdp * d.rdp;
  d.rdr2 = 0.5 * d.rdr; d.rdt2 = 0.5 * d.rdt; d.rdp2 = 0.5 * d.rdp;
  d.rdr4 = 0.25 * d.rdr; d.rdt4 = 0.25 * d.rdt; d.rdp4 = 0.25 * d.rdp;
  d.rdpdt4 = 0.25 * d.rdp * d.rdt;
  d.rdtdr4 = 0.25 * d.rdt * d.rdr;
  // src/ModScbInit.f90:246-273
  const double xzero3 = 6.6 * 6.6 * 6.6;
  h->bnormal = 0.31 / xzero3 * 1.E5;
  h->pnormal = h->bnormal * h->bnormal / (4. * PI_D * 1.E-7) * 1.E-9;
  h->pjconst = 1.e6 * 0.31E-4 / (xzero3 * 4. * PI_D * 1.E-7 * 6.4E6);
  const size_t n3 = (size_t)nthe * npsi * nzeta, n3p = (size_t)nthe * npsi * (nzeta + 1);
  SRET(h->dalloc((double**)&d.thetaVal, nthe, "thetaVal"));
  SRET(h->dalloc((double**)&d.rhoVal, npsi, "rhoVal"));
  SRET(h->dalloc((double**)&d.zetaVal, nzeta, "zetaVal"));
  SRET(h->dalloc((double**)&d.f, npsi, "f"));
  SRET(h->dalloc((double**)&d.fzet, nzeta + 1, "fzet"));
  struct { double** p; const char* n; } p1[] = {{&d.x, "x"}, {&d.y, "y"}, {&d.z, "z"}, {&d.alfa, "alfa"}, {&d.psi, "psi"},
      {&d.pper, "pper"}, {&d.ppar, "ppar"}, {&d.sigma, "sigma"}, {&d.bsq, "bsq"}, {&d.bf, "bf"}};
  for (auto& e : p1) SRET(h->dalloc(e.p, n3p, e.n));
  struct { double** p; const char* n; } p0[] = {
      {&d.dXT, "derivXTheta"}, {&d.dXR, "derivXRho"}, {&d.dXZ, "derivXZeta"}, {&d.dYT, "derivYTheta"}, {&d.dYR, "derivYRho"},
      {&d.dYZ, "derivYZeta"}, {&d.dZT, "derivZTheta"}, {&d.dZR, "derivZRho"}, {&d.dZZ, "derivZZeta"}, {&d.jac, "jacobian"},
      {&d.gRX, "gradRhoX"}, {&d.gRY, "gradRhoY"}, {&d.gRZ, "gradRhoZ"}, {&d.gZX, "gradZetaX"}, {&d.gZY, "gradZetaY"},
      {&d.gZZ, "gradZetaZ"}, {&d.gTX, "gradThetaX"}, {&d.gTY, "gradThetaY"}, {&d.gTZ, "gradThetaZ"}, {&d.GRS, "GradRhoSq"},
      {&d.GTS, "GradThetaSq"}, {&d.GZS, "GradZetaSq"}, {&d.GRGT, "GradRhoGradTheta"}, {&d.GRGZ, "GradRhoGradZeta"},
      {&d.GTGZ, "GradThetaGradZeta"}, {&d.Bx, "Bx"}, {&d.By, "By"}, {&d.Bz, "Bz"}, {&d.vecd, "vecd"}, {&d.vec1, "vec1"},
      {&d.vec2, "vec2"}, {&d.vec3, "vec3"}, {&d.vec4, "vec4"}, {&d.vec6, "vec6"}, {&d.vec7, "vec7"}, {&d.vec8, "vec8"},
      {&d.vec9, "vec9"}, {&d.vecx, "vecx"}, {&d.vecr, "vecr"}, {&d.dPT, "dPPerdTheta"}, {&d.dPR, "dPPerdRho"},
      {&d.dPZ, "dPPerdZeta"}, {&d.dBT, "dBsqdTheta"}, {&d.dBR, "dBsqdRho"}, {&d.dBZ, "dBsqdZeta"}, {&d.dPP, "dPPerdPsi"},
      {&d.dPA, "dPPerdAlpha"}, {&d.dBP, "dBsqdPsi"}, {&d.dBA, "dBsqdAlpha"}, {&d.dPdAlpha, "dPdAlpha"}, {&d.dPdPsi, "dPdPsi"},
      {&d.jGR, "jGradRho"}, {&d.jGZ, "jGradZeta"}, {&d.jGT, "jGradTheta"}, {&d.Jx, "Jx"}, {&d.Jy, "Jy"}, {&d.Jz, "Jz"},
      {&d.GPx, "GradPx"}, {&d.GPy, "GradPy"}, {&d.GPz, "GradPz"}, {&d.jCrossB, "jCrossB"}, {&d.GradP, "GradP"},
      {&d.w1, nullptr}, {&d.w2, nullptr}, {&d.w3, nullptr}, {&d.w4, nullptr}, {&d.w5, nullptr}};
  for (auto& e : p0) SRET(h->dalloc(e.p, n3, e.n));
  SRET(h->dalloc(&h->d_prev, n3p, nullptr));
  h->npart = (size_t)4 * nblk(nthe, 128) * npsi * nzeta + 2 * (size_t)nzeta;
  SRET(h->dalloc(&h->d_part, h->npart, nullptr));
  const int nsub = std::max(npsi, nzeta) + 1;
  SRET(h->dalloc(&h->d_resmax, nsub, nullptr));
  void* q = nullptr;
  SCK(cudaMalloc(&q, sizeof(int) * (nsub + 1)));
  h->allocs.push_back(q);
  h->d_ni = (int*)q;
  h->d_fail = h->d_ni + nsub;
  SCK(cudaMemset(q, 0, sizeof(int) * (nsub + 1)));
  SCK(cudaStreamCreateWithFlags(&h->own_st, cudaStreamNonBlocking));
  h->st = h->own_st;
  SCK(cudaEventCreate(&h->e0));
  SCK(cudaEventCreate(&h->e1));
  *out = h;
  return RSG_OK;
}

int rsg_scb_destroy(rsg_scb* h) {
  if (!h) return RSG_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  if (h->own_st) cudaStreamDestroy(h->own_st);
  if (h->e0) cudaEventDestroy(h->e0);
  if (h->e1) cudaEventDestroy(h->e1);
  delete h;
  return RSG_OK;
}

static int scb_up(rsg_scb* h, const char* name, const double* src) {
  auto it = h->arr.find(name);
  if (it == h->arr.end()) return sfail(RSG_ERR_ARG, std::string("unknown array ") + name);
  if (!src) return sfail(RSG_ERR_ARG, std::string("null pointer for ") + name);
  SCK(cudaMemcpyAsync(it->second.first, src, it->second.second * sizeof(double), cudaMemcpyHostToDevice, h->st));
  return RSG_OK;
}

int rsg_scb_set_grid(rsg_scb* h, const double* thetaVal, const double* rhoVal, const double* zetaVal, const double* f,
                     const double* fzet) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  SCK(cudaSetDevice(h->device));
  SRET(scb_up(h, "thetaVal", thetaVal)); SRET(scb_up(h, "rhoVal", rhoVal)); SRET(scb_up(h, "zetaVal", zetaVal));
  SRET(scb_up(h, "f", f)); SRET(scb_up(h, "fzet", fzet));
  SCK(cudaStreamSynchronize(h->st));
  h->grid_set = true;
  return RSG_OK;
}

int rsg_scb_set_geometry(rsg_scb* h, const double* x, const double* y, const double* z) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  SCK(cudaSetDevice(h->device));
  SRET(scb_up(h, "x", x)); SRET(scb_up(h, "y", y)); SRET(scb_up(h, "z", z));
  SCK(cudaStreamSynchronize(h->st));
  h->geom_set = true;
  h->band_done = false;
  return RSG_OK;
}

int rsg_scb_set_pressure(rsg_scb* h, int isotropy, const double* pper, const double* ppar, const double* sigma,
                         const double* dPPerdTheta, const double* dPPerdRho, const double* dPPerdZeta, const double* dBsqdTheta,
                         const double* dBsqdRho, const double* dBsqdZeta, const double* dPPerdPsi, const double* dPPerdAlpha,
                         const double* dBsqdPsi, const double* dBsqdAlpha, const double* dPdAlpha, const double* dPdPsi) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  SCK(cudaSetDevice(h->device));
  h->isotropy = isotropy;
  if (isotropy == 1) {
    SRET(scb_up(h, "dPdAlpha", dPdAlpha)); SRET(scb_up(h, "dPdPsi", dPdPsi));
    // Compute_convergence reads pper, ppar and dPPerd{Rho,Zeta,Theta} in both branches (:620-631):
    // whatever the host holds in them is mirrored when given
    if (pper) SRET(scb_up(h, "pper", pper));
    if (ppar) SRET(scb_up(h, "ppar", ppar));
    if (sigma) SRET(scb_up(h, "sigma", sigma));
    if (dPPerdTheta) SRET(scb_up(h, "dPPerdTheta", dPPerdTheta));
    if (dPPerdRho) SRET(scb_up(h, "dPPerdRho", dPPerdRho));
    if (dPPerdZeta) SRET(scb_up(h, "dPPerdZeta", dPPerdZeta));
  } else {
    SRET(scb_up(h, "pper", pper)); SRET(scb_up(h, "ppar", ppar)); SRET(scb_up(h, "sigma", sigma));
    SRET(scb_up(h, "dPPerdTheta", dPPerdTheta)); SRET(scb_up(h, "dPPerdRho", dPPerdRho)); SRET(scb_up(h, "dPPerdZeta", dPPerdZeta));
    SRET(scb_up(h, "dBsqdTheta", dBsqdTheta)); SRET(scb_up(h, "dBsqdRho", dBsqdRho)); SRET(scb_up(h, "dBsqdZeta", dBsqdZeta));
    SRET(scb_up(h, "dPPerdPsi", dPPerdPsi)); SRET(scb_up(h, "dPPerdAlpha", dPPerdAlpha));
    SRET(scb_up(h, "dBsqdPsi", dBsqdPsi)); SRET(scb_up(h, "dBsqdAlpha", dBsqdAlpha));
  }
  SCK(cudaStreamSynchronize(h->st));
  h->press_set = true;
  return RSG_OK;
}

int rsg_scb_set_field(rsg_scb* h, const char* name, const double* src) {
  if (!h || !name) return sfail(RSG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(h->device));
  SRET(scb_up(h, name, src));
  SCK(cudaStreamSynchronize(h->st));
  return RSG_OK;
}

int rsg_scb_get_field(rsg_scb* h, const char* name, double* dst) {
  if (!h || !name || !dst) return sfail(RSG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(h->device));
  auto it = h->arr.find(name);
  if (it == h->arr.end()) return sfail(RSG_ERR_ARG, std::string("unknown array ") + name);
  SCK(cudaMemcpyAsync(dst, it->second.first, it->second.second * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  return RSG_OK;
}

int rsg_scb_field_size(rsg_scb* h, const char* name, long long* n) {
  if (!h || !name || !n) return sfail(RSG_ERR_ARG, "null argument");
  auto it = h->arr.find(name);
  if (it == h->arr.end()) return sfail(RSG_ERR_ARG, std::string("unknown array ") + name);
  *n = (long long)it->second.second;
  return RSG_OK;
}

// computeBandJacob, src/ModScbCompute.f90:412-496
int rsg_scb_bandjacob(rsg_scb* h, int* sorfail) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  if (!h->grid_set || !h->geom_set) return sfail(RSG_ERR_STATE, "computeBandJacob before set_grid/set_geometry");
  SCK(cudaSetDevice(h->device));
  SCK(cudaMemsetAsync(h->d_fail, 0, sizeof(int), h->st));
  SCK(cudaEventRecord(h->e0, h->st));
  dim3 g(nblk(h->nthe, 128), h->npsi, h->nzeta);
  k_scb_bandjacob<<<g, 128, 0, h->st>>>(h->dev, h->d_fail);
  SCKL();
  k_scb_bwrap<<<dim3(nblk(h->nthe, 128), h->npsi), 128, 0, h->st>>>(h->dev);
  SCKL();
  SCK(cudaEventRecord(h->e1, h->st));
  h->launches += 2;
  int f = 0;
  SCK(cudaMemcpyAsync(&f, h->d_fail, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->last_ms = ms;
  if (sorfail) *sorfail = f;
  h->band_done = true;
  return RSG_OK;
}

static int scb_metric(rsg_scb* h, bool alpha) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  if (!h->geom_set) return sfail(RSG_ERR_STATE, "metric before set_geometry");
  SCK(cudaSetDevice(h->device));
  const size_t n3 = (size_t)h->nthe * h->npsi * h->nzeta;
  SCK(cudaEventRecord(h->e0, h->st));
  double* vs[] = {h->dev.vecd, h->dev.vec1, h->dev.vec2, h->dev.vec3, h->dev.vec4, h->dev.vec6, h->dev.vec7, h->dev.vec8, h->dev.vec9};
  for (double* v : vs) SCK(cudaMemsetAsync(v, 0, n3 * sizeof(double), h->st));
  dim3 g(nblk(h->nthe - 2, 128), h->npsi - 2, h->nzeta - 1);
  if (alpha) k_scb_metric<true><<<g, 128, 0, h->st>>>(h->dev);
  else k_scb_metric<false><<<g, 128, 0, h->st>>>(h->dev);
  SCKL();
  SCK(cudaEventRecord(h->e1, h->st));
  h->launches++;
  SCK(cudaStreamSynchronize(h->st));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->last_ms = ms;
  return RSG_OK;
}
int rsg_scb_metrica(rsg_scb* h) { return scb_metric(h, true); }   // src/ModScbEquation.f90:18-280
int rsg_scb_metric(rsg_scb* h) { return scb_metric(h, false); }   // :283-540

static int scb_rhs(rsg_scb* h, bool alpha) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  if (!h->band_done || !h->press_set) return sfail(RSG_ERR_STATE, "newk/newj before computeBandJacob/set_pressure");
  SCK(cudaSetDevice(h->device));
  dim3 g(nblk(h->nthe, 128), h->npsi, h->nzeta);
  if (alpha) k_scb_rhs<true><<<g, 128, 0, h->st>>>(h->dev, h->isotropy);
  else k_scb_rhs<false><<<g, 128, 0, h->st>>>(h->dev, h->isotropy);
  SCKL();
  h->launches++;
  SCK(cudaStreamSynchronize(h->st));
  return RSG_OK;
}
int rsg_scb_newk(rsg_scb* h) { return scb_rhs(h, true); }    // :546-604
int rsg_scb_newj(rsg_scb* h) { return scb_rhs(h, false); }   // :607-665

// iterateAlpha / iteratePsi, src/ModScbEuler.f90:160-299 / :469-612
// part 1: the SOR solves of sub-problems [sub0, sub0+nsub_l) (nsub_l < 0: all); part 2
// (scb_iterate_finish): sums, extrapolation / theta fill / periodic wrap, results.  Ranks that
// shard the independent sub-problems all-gather the solved planes between the two.
static int scb_iterate_finish(rsg_scb* h, bool alpha, int theChange, int psiChange, int* nisave, double* sumb, double* sumdb,
                              double* diffmx, int* sorfail, int* ni_out);
static int scb_iterate_part(rsg_scb* h, bool alpha, double tol, int nimax, int theChange, int psiChange, int ordering, int sub0,
                            int nsub_l) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  if (ordering != RSG_SOR_LEX && ordering != RSG_SOR_COLOR4) return sfail(RSG_ERR_ARG, "unknown SOR ordering");
  SCK(cudaSetDevice(h->device));
  const int nthe = h->nthe, npsi = h->npsi, nzeta = h->nzeta;
  const int nT = std::max(theChange, 1), nP = std::max(psiChange, 1);
  if (nT == 1) return sfail(RSG_ERR_UNSUPPORTED, "theChange <= 1 is not supported");
  if (nthe - 2 * nT < 3) return sfail(RSG_ERR_ARG, "theChange too large for nthe");
  double* u = alpha ? h->dev.alfa : h->dev.psi;
  const size_t n3p = (size_t)nthe * npsi * (nzeta + 1);
  SCK(cudaMemcpyAsync(h->d_prev, u, n3p * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  const int nsub_all = alpha ? (npsi - nP - 1) : (nzeta - 1);
  if (nsub_l < 0) { sub0 = 0; nsub_l = nsub_all; }
  if (sub0 < 0 || nsub_l < 0 || sub0 + nsub_l > nsub_all) return sfail(RSG_ERR_ARG, "sub-problem range out of bounds");
  const int nsub = nsub_l;
  SCK(cudaMemsetAsync(h->d_ni, 0, sizeof(int) * (std::max(npsi, nzeta) + 2), h->st));
  SCK(cudaMemsetAsync(h->d_resmax, 0, sizeof(double) * (std::max(npsi, nzeta) + 1), h->st));
  SorArgs a;
  a.sub0 = sub0;
  a.tol = tol;
  a.nimax = nimax;
  a.nT = nT;
  a.nP = nP;
  a.ni = h->d_ni;
  a.resmax = h->d_resmax;
  a.fail = h->d_fail;
  const double rjac = alpha ? 1.0 - 2.0 * PI_D * PI_D / ((double)nzeta * (double)nzeta + (double)nthe * (double)nthe)
                            : 1.0 - 2.0 * PI_D * PI_D / ((double)nthe * (double)nthe + (double)npsi * (double)npsi);
  a.omegaOpt = 2.0 / (1.0 + std::sqrt(1.0 - rjac * rjac));
  const int nrows = alpha ? nzeta + 1 : npsi;
  const size_t smem = sizeof(double) * (size_t)nrows * nthe;
  if (smem > 220 * 1024) return sfail(RSG_ERR_UNSUPPORTED, "SOR plane does not fit in shared memory");
  const int nr = alpha ? (nzeta - 1) : (npsi - nP - 1);
  const int threads = ordering == RSG_SOR_LEX ? std::min(1024, (nr + 31) / 32 * 32) : 1024;
  if (ordering == RSG_SOR_LEX && nr > 1024) return sfail(RSG_ERR_UNSUPPORTED, "too many rows for the lexicographic SOR kernel");
  // 4-colour ordering: a cluster of CL CTAs per sub-problem keeps the unknown AND the ten
  // coefficient arrays in (distributed) shared memory for the whole solve
  int CL = 0, nloc_max = 0, npc_max = 0;
  size_t csmem = 0;
  bool in_regs = false;     // coefficients in registers: one point per colour per thread (<= 576)
  if (ordering != RSG_SOR_LEX && h->use_cluster && !getenv("RSG_SCB_NO_CLUSTER")) {
    const int nc = nthe - 2 * nT;
    if (!getenv("RSG_SCB_NO_REGS")) {
      // smallest cluster whose CTAs hold at most one point per colour per thread; 768 threads (85 registers) are allowed
      // when that lets ALL sub-problems run in one wave of the device's SMs (alpha on the default grid: 43 x 3 CTAs
      // instead of 43 x 4 = 172 on 148 SMs, i.e. two waves)
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
      const int wide = getenv("RSG_SCB_NO_WIDE") ? 576 : 768;
      int best = 0;
      for (int c = 1; c <= 8; ++c) {
        const int nl = (nr + c - 1) / c;
        const int npc = ((nl + 1) / 2) * ((nc + 1) / 2);
        if (nl < 2 || npc > wide) continue;
        const bool one_wave = nsub * c <= sms;
        if (npc <= 576 && (c & (c - 1)) == 0 && !best) best = c;          // round 1's choice: power of two, 576 threads
        if (one_wave && (npc <= 576 || wide > 576)) { best = c; break; }   // first (smallest) cluster that fits one wave
      }
      if (best) {
        const int nl = (nr + best - 1) / best;
        CL = best; nloc_max = nl; npc_max = ((nl + 1) / 2) * ((nc + 1) / 2); in_regs = true;
        csmem = sizeof(double) * (size_t)(nl + 2) * nthe;
      }
    }
    for (int c = 1; c <= 8 && !in_regs; c *= 2) {
      const int nl = (nr + c - 1) / c;
      const int npc = ((nl + 1) / 2) * ((nc + 1) / 2);
      const size_t b = sizeof(double) * ((size_t)(nl + 2) * nthe + (size_t)40 * npc);
      if (b <= 220 * 1024 && nl >= 2) { CL = c; nloc_max = nl; npc_max = npc; csmem = b; break; }
    }
  }
  h->last_cluster = CL;
  SCK(cudaEventRecord(h->e0, h->st));
  if (nsub == 0) {
    // nothing to solve on this rank
  } else if (CL > 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nsub * CL);
    cfg.blockDim = dim3(std::min(1024, (npc_max + 31) / 32 * 32));
    cfg.dynamicSmemBytes = csmem;
    cfg.stream = h->st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const bool wideT = npc_max > 576;
    if (in_regs && alpha && wideT) {
      SCK(cudaFuncSetAttribute(k_scb_sor_cluster_reg<true, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      SCK(cudaLaunchKernelEx(&cfg, k_scb_sor_cluster_reg<true, 768>, h->dev, a, nloc_max));
    } else if (in_regs && wideT) {
      SCK(cudaFuncSetAttribute(k_scb_sor_cluster_reg<false, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      SCK(cudaLaunchKernelEx(&cfg, k_scb_sor_cluster_reg<false, 768>, h->dev, a, nloc_max));
    } else if (in_regs && alpha) {
      SCK(cudaFuncSetAttribute(k_scb_sor_cluster_reg<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      SCK(cudaLaunchKernelEx(&cfg, k_scb_sor_cluster_reg<true>, h->dev, a, nloc_max));
    } else if (in_regs) {
      SCK(cudaFuncSetAttribute(k_scb_sor_cluster_reg<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      SCK(cudaLaunchKernelEx(&cfg, k_scb_sor_cluster_reg<false>, h->dev, a, nloc_max));
    } else if (alpha) {
      SCK(cudaFuncSetAttribute(k_scb_sor_cluster<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      SCK(cudaLaunchKernelEx(&cfg, k_scb_sor_cluster<true>, h->dev, a, nloc_max, npc_max));
    } else {
      SCK(cudaFuncSetAttribute(k_scb_sor_cluster<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      SCK(cudaLaunchKernelEx(&cfg, k_scb_sor_cluster<false>, h->dev, a, nloc_max, npc_max));
    }
  } else {
#define LAUNCH_SOR(A, O)                                                                                   \
  do {                                                                                                     \
    SCK(cudaFuncSetAttribute(k_scb_sor<A, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    k_scb_sor<A, O><<<nsub, threads, smem, h->st>>>(h->dev, a);                                            \
  } while (0)
  if (alpha && ordering == RSG_SOR_LEX) LAUNCH_SOR(true, 0);
  else if (alpha) LAUNCH_SOR(true, 1);
  else if (ordering == RSG_SOR_LEX) LAUNCH_SOR(false, 0);
  else LAUNCH_SOR(false, 1);
#undef LAUNCH_SOR
  }
  SCKL();
  SCK(cudaEventRecord(h->e1, h->st));
  if (nsub > 0) h->launches++;
  return RSG_OK;
}
static int scb_iterate_finish(rsg_scb* h, bool alpha, int theChange, int psiChange, int* nisave, double* sumb, double* sumdb,
                              double* diffmx, int* sorfail, int* ni_out) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  SCK(cudaSetDevice(h->device));
  const int nthe = h->nthe, npsi = h->npsi, nzeta = h->nzeta;
  const int nT = std::max(theChange, 1), nP = std::max(psiChange, 1);
  double* u = alpha ? h->dev.alfa : h->dev.psi;
  const int nsub = alpha ? (npsi - nP - 1) : (nzeta - 1);
  k_scb_sums<<<nzeta - 1, 256, 0, h->st>>>(h->dev, u, h->d_prev, h->d_part);
  SCKL();
  k_scb_post_extap<<<dim3(nblk(nthe, 128), nzeta - 1), 128, 0, h->st>>>(h->dev, u, nT, nP);
  SCKL();
  k_scb_post_theta<<<dim3(nblk(npsi, 64), nzeta), 64, 0, h->st>>>(h->dev, u, nT);
  SCKL();
  k_scb_post_wrap<<<dim3(nblk(nthe, 128), npsi), 128, 0, h->st>>>(h->dev, u, alpha ? 2.0 * PI_D : 0.0);
  SCKL();
  h->launches += 4;
  std::vector<int> ni(nsub + 1);
  std::vector<double> rm(nsub), part(2 * (size_t)(nzeta - 1));
  int f = 0;
  SCK(cudaMemcpyAsync(ni.data(), h->d_ni, sizeof(int) * nsub, cudaMemcpyDeviceToHost, h->st));
  SCK(cudaMemcpyAsync(rm.data(), h->d_resmax, sizeof(double) * nsub, cudaMemcpyDeviceToHost, h->st));
  SCK(cudaMemcpyAsync(part.data(), h->d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaMemcpyAsync(&f, h->d_fail, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->last_ms = ms;
  int nmax = 0;
  double dmx = 0.0, sb = 0.0, sdb = 0.0;
  for (int q = 0; q < nsub; ++q) {
    nmax = std::max(nmax, ni[q]);
    dmx = std::max(dmx, rm[q]);
  }
  for (int k = 0; k < nzeta - 1; ++k) {
    sb += part[2 * k];
    sdb += part[2 * k + 1];
  }
  if (nisave) *nisave = nmax;
  if (sumb) *sumb = sb;
  if (sumdb) *sumdb = sdb;
  if (diffmx) *diffmx = dmx;
  if (sorfail) *sorfail = f;
  if (ni_out) {
    // Fortran ni(1:npsi) (alpha, entries jz=2..npsi-nP) / ni(1:nzeta) (psi, entries k=2..nzeta)
    const int n = alpha ? npsi : nzeta;
    for (int q = 0; q < n; ++q) ni_out[q] = 0;
    for (int q = 0; q < nsub; ++q) ni_out[q + 1] = ni[q];
  }
  if (f) SCK(cudaMemsetAsync(h->d_fail, 0, sizeof(int), h->st));
  return RSG_OK;
}
static int scb_iterate(rsg_scb* h, bool alpha, double tol, int nimax, int theChange, int psiChange, int ordering, int* nisave,
                       double* sumb, double* sumdb, double* diffmx, int* sorfail, int* ni_out) {
  const int rc = scb_iterate_part(h, alpha, tol, nimax, theChange, psiChange, ordering, 0, -1);
  if (rc != RSG_OK) return rc;
  return scb_iterate_finish(h, alpha, theChange, psiChange, nisave, sumb, sumdb, diffmx, sorfail, ni_out);
}
// The independent sub-problems (psi surfaces jz for alpha, zeta planes k for psi) of one solve split
// among ranks: part solves [sub0, sub0+nsub) (0-based; sub-problem q is jz = q+2 / k = q+2); the
// caller all-gathers the solved planes of the field (rsg_scb_field_device) and calls finish,
// whose nisave / diffmx / ni cover this rank's sub-problems only (reduce with max over ranks).
int rsg_scb_iterate_part(rsg_scb* h, int alpha, double tol, int nimax, int theChange, int psiChange, int ordering, int sub0,
                         int nsub) {
  if (nsub < 0) return sfail(RSG_ERR_ARG, "negative sub-problem count");
  return scb_iterate_part(h, alpha != 0, tol, nimax, theChange, psiChange, ordering, sub0, nsub);
}
int rsg_scb_iterate_finish(rsg_scb* h, int alpha, int theChange, int psiChange, int* nisave, double* sumb, double* sumdb,
                           double* diffmx, int* sorfail, int* ni) {
  return scb_iterate_finish(h, alpha != 0, theChange, psiChange, nisave, sumb, sumdb, diffmx, sorfail, ni);
}
int rsg_scb_field_device(rsg_scb* h, const char* name, void** ptr, long long* n) {
  if (!h || !name || !ptr || !n) return sfail(RSG_ERR_ARG, "null argument");
  auto it = h->arr.find(name);
  if (it == h->arr.end()) return sfail(RSG_ERR_ARG, std::string("unknown field ") + name);
  *ptr = it->second.first;
  *n = (long long)it->second.second;
  return RSG_OK;
}
// iterateAlpha sharded along ZETA: this rank relaxes the zeta planes (0-based rows)
// [k0, k0+nk) of every psi surface; the caller exchanges the edge planes after every half-sweep and
// all-reduces (MAX) the state vector after every sweep.  Bit-identical to RSG_SOR_COLOR4 on one GPU.
int rsg_scb_zsolve_begin(rsg_scb* h, double tol, int nimax, int theChange, int psiChange, int k0, int nk) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  SCK(cudaSetDevice(h->device));
  const int nthe = h->nthe, npsi = h->npsi, nzeta = h->nzeta;
  const int nT = std::max(theChange, 1), nP = std::max(psiChange, 1);
  if (nT == 1) return sfail(RSG_ERR_UNSUPPORTED, "theChange <= 1 is not supported");
  if (nthe - 2 * nT < 3) return sfail(RSG_ERR_ARG, "theChange too large for nthe");
  if (nk < 0 || (nk > 0 && (k0 < 1 || k0 + nk > nzeta))) return sfail(RSG_ERR_ARG, "zeta range outside the updated planes 1..nzeta-1");
  const int nsub = npsi - nP - 1;
  if (!h->d_zstate) {
    const int cap = npsi + 1;
    SRET(h->dalloc(&h->d_zstate, 2 * (size_t)cap, nullptr));
    void* q = nullptr;
    SCK(cudaMalloc(&q, sizeof(int) * (cap + 1)));
    h->allocs.push_back(q);
    h->d_zdone = (int*)q;
    h->d_zpend = h->d_zdone + cap;
  }
  const size_t n3p = (size_t)nthe * npsi * (nzeta + 1);
  SCK(cudaMemcpyAsync(h->d_prev, h->dev.alfa, n3p * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  std::vector<int> one(std::max(npsi, nzeta) + 2, 0);
  for (int q = 0; q < nsub; ++q) one[q] = 1;                       // ni = 1 before the first sweep (:205)
  SCK(cudaMemcpyAsync(h->d_ni, one.data(), sizeof(int) * one.size(), cudaMemcpyHostToDevice, h->st));
  SCK(cudaMemsetAsync(h->d_resmax, 0, sizeof(double) * (std::max(npsi, nzeta) + 1), h->st));
  SCK(cudaMemsetAsync(h->d_zstate, 0, sizeof(double) * 2 * (size_t)(npsi + 1), h->st));
  std::vector<int> done(npsi + 2, nimax < 1 ? 1 : 0);
  done[npsi + 1] = nimax < 1 ? 0 : nsub;                            // pending
  SCK(cudaMemcpyAsync(h->d_zdone, done.data(), sizeof(int) * done.size(), cudaMemcpyHostToDevice, h->st));
  SCK(cudaStreamSynchronize(h->st));                                // `one` / `done` are stack-owned
  ZArgs& a = h->z;
  a.tol = tol;
  a.nT = nT; a.k0 = k0; a.nk = nk; a.nsub = nsub; a.nimax = nimax;
  a.state = h->d_zstate;
  a.done = h->d_zdone;
  a.u0 = h->d_prev;
  const double rjac = 1.0 - 2.0 * PI_D * PI_D / ((double)nzeta * (double)nzeta + (double)nthe * (double)nthe);
  a.om = 2.0 / (1.0 + std::sqrt(1.0 - rjac * rjac));                // omegaOpt; the first sweep runs with 1 (:207-214)
  h->z_sweep = 0;
  h->z_open = true;
  SCK(cudaEventRecord(h->e0, h->st));
  SCK(cudaEventRecord(h->e1, h->st));
  return RSG_OK;
}
int rsg_scb_zsolve_half(rsg_scb* h, int parity) {
  if (!h || !h->z_open) return sfail(RSG_ERR_ARG, "no open zeta-sharded solve");
  if (parity != 0 && parity != 1) return sfail(RSG_ERR_ARG, "parity is 0 or 1");
  SCK(cudaSetDevice(h->device));
  ZArgs a = h->z;
  if (h->z_sweep == 0) a.om = 1.0;
  const int rs = a.k0 + (((a.k0 & 1) == parity) ? 0 : 1);
  const int nrows = rs < a.k0 + a.nk ? (a.k0 + a.nk - 1 - rs) / 2 + 1 : 0;
  if (nrows > 0 && a.nsub > 0) {
    k_scb_zhalf<<<dim3(nrows, a.nsub), 64, 0, h->st>>>(h->dev, a, parity);
    SCKL();
    h->launches++;
  }
  return RSG_OK;
}
int rsg_scb_zsolve_state_device(rsg_scb* h, void** ptr, long long* n) {
  if (!h || !h->z_open || !ptr || !n) return sfail(RSG_ERR_ARG, "no open zeta-sharded solve");
  *ptr = h->d_zstate;
  *n = 2 * (long long)h->z.nsub;
  return RSG_OK;
}
int rsg_scb_zsolve_commit(rsg_scb* h) {
  if (!h || !h->z_open) return sfail(RSG_ERR_ARG, "no open zeta-sharded solve");
  SCK(cudaSetDevice(h->device));
  k_scb_zcommit<<<1, 64, 0, h->st>>>(h->z, h->d_ni, h->d_resmax, h->d_fail, h->d_zpend);
  SCKL();
  h->launches++;
  h->z_sweep++;
  SCK(cudaEventRecord(h->e1, h->st));
  return RSG_OK;
}
int rsg_scb_zsolve_pending(rsg_scb* h, int* pending) {
  if (!h || !h->z_open || !pending) return sfail(RSG_ERR_ARG, "no open zeta-sharded solve");
  SCK(cudaSetDevice(h->device));
  SCK(cudaMemcpyAsync(pending, h->d_zpend, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  return RSG_OK;
}
int rsg_scb_set_stream(rsg_scb* h, void* stream) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  SCK(cudaSetDevice(h->device));
  SCK(cudaStreamSynchronize(h->st));
  h->st = stream ? (cudaStream_t)stream : h->own_st;
  return RSG_OK;
}
int rsg_scb_iterate_alpha(rsg_scb* h, double InConAlpha, int nimax, int theChange, int psiChange, int ordering, int* nisave,
                          double* sumb, double* sumdb, double* diffmx, int* sorfail, int* ni) {
  return scb_iterate(h, true, InConAlpha, nimax, theChange, psiChange, ordering, nisave, sumb, sumdb, diffmx, sorfail, ni);
}
int rsg_scb_iterate_psi(rsg_scb* h, double InConPsi, int nimax, int theChange, int psiChange, int ordering, int* nisave,
                        double* sumb, double* sumdb, double* diffmx, int* sorfail, int* ni) {
  return scb_iterate(h, false, InConPsi, nimax, theChange, psiChange, ordering, nisave, sumb, sumdb, diffmx, sorfail, ni);
}

// Compute_convergence, src/ModScbCompute.f90:499-754 (anisotropic branch)
int rsg_scb_convergence(rsg_scb* h, double* normDiff, double* normJxB, double* normGradP, int* sorfail) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  if (!h->band_done || !h->press_set) return sfail(RSG_ERR_STATE, "Compute_convergence before computeBandJacob/set_pressure");
  SCK(cudaSetDevice(h->device));
  dim3 g(nblk(h->nthe, 128), h->npsi, h->nzeta);
  SCK(cudaEventRecord(h->e0, h->st));
  k_scb_conv1<<<g, 128, 0, h->st>>>(h->dev, h->isotropy);
  SCKL();
  k_scb_derivs<<<g, 128, 0, h->st>>>(h->dev, h->dev.w1, nullptr, h->dev.w4, nullptr);   // d/drho of jGradThetaPartialRho
  SCKL();
  k_scb_derivs<<<g, 128, 0, h->st>>>(h->dev, h->dev.w2, nullptr, nullptr, h->dev.w5);   // d/dzeta of jGradThetaPartialZeta
  SCKL();
  k_scb_derivs<<<g, 128, 0, h->st>>>(h->dev, h->dev.w3, h->dev.w1, nullptr, nullptr);   // d/dtheta of J(pper-ppar) -> w1
  SCKL();
  k_scb_conv2<<<g, 128, 0, h->st>>>(h->dev, h->bnormal, h->pnormal, h->pjconst, h->d_part, h->isotropy);
  SCKL();
  SCK(cudaEventRecord(h->e1, h->st));
  h->launches += 5;
  const size_t ncta = (size_t)g.x * g.y * g.z;
  std::vector<double> part(4 * ncta);
  SCK(cudaMemcpyAsync(part.data(), h->d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->e0, h->e1);
  h->last_ms = ms;
  double s[4] = {0, 0, 0, 0};
  for (size_t c = 0; c < ncta; ++c)
    for (int m = 0; m < 4; ++m) s[m] += part[4 * c + m];
  const double nd = s[0] / s[3], nj = s[1] / s[3], ng = s[2] / s[3];
  if (normDiff) *normDiff = nd;
  if (normJxB) *normJxB = nj;
  if (normGradP) *normGradP = ng;
  if (sorfail) *sorfail = (std::isnan(nd) || std::isnan(nj) || std::isnan(ng)) ? 1 : 0;
  return RSG_OK;
}

// GSL_Derivs (Steffen) of an arbitrary (nthe,npsi,nzeta) host field -- exposed for tests and
// for host code that still needs the derivative of its own arrays (e.g. `pressure`)
int rsg_scb_derivs(rsg_scb* h, const double* f, double* dT, double* dR, double* dZ) {
  if (!h || !f) return sfail(RSG_ERR_ARG, "null argument");
  if (!h->grid_set) return sfail(RSG_ERR_STATE, "derivs before set_grid");
  SCK(cudaSetDevice(h->device));
  const size_t n3 = (size_t)h->nthe * h->npsi * h->nzeta;
  SCK(cudaMemcpyAsync(h->dev.w1, f, n3 * sizeof(double), cudaMemcpyHostToDevice, h->st));
  dim3 g(nblk(h->nthe, 128), h->npsi, h->nzeta);
  k_scb_derivs<<<g, 128, 0, h->st>>>(h->dev, h->dev.w1, dT ? h->dev.w2 : nullptr, dR ? h->dev.w3 : nullptr, dZ ? h->dev.w4 : nullptr);
  SCKL();
  h->launches++;
  if (dT) SCK(cudaMemcpyAsync(dT, h->dev.w2, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (dR) SCK(cudaMemcpyAsync(dR, h->dev.w3, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (dZ) SCK(cudaMemcpyAsync(dZ, h->dev.w4, n3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  return RSG_OK;
}

// prescribed node values of the Euler potentials and of the field-line coordinate
// (alphaVal(nzeta+1), psiVal(npsi): src/ModScbIO.f90:179-203; chiVal(nthe): :167)
int rsg_scb_set_map_targets(rsg_scb* h, const double* alphaVal, const double* psiVal, const double* chiVal) {
  if (!h || !alphaVal || !psiVal || !chiVal) return sfail(RSG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(h->device));
  if (!h->d_alphaVal) {
    SRET(h->dalloc(&h->d_alphaVal, h->nzeta + 1, "alphaVal"));
    SRET(h->dalloc(&h->d_psiVal, h->npsi, "psiVal"));
    SRET(h->dalloc(&h->d_chiVal, h->nthe, "chiVal"));
    const size_t nmax = std::max({(size_t)(h->nzeta + 1) * h->nthe * h->npsi, (size_t)h->npsi * h->nthe * (h->nzeta - 1),
                                  (size_t)h->nthe * h->npsi * (h->nzeta - 1)});
    SRET(h->dalloc(&h->d_mapw, 7 * nmax, nullptr));
  }
  SRET(scb_up(h, "alphaVal", alphaVal)); SRET(scb_up(h, "psiVal", psiVal)); SRET(scb_up(h, "chiVal", chiVal));
  SCK(cudaStreamSynchronize(h->st));
  h->map_set = true;
  return RSG_OK;
}

// mapAlpha / mapPsi / mapTheta (src/ModScbEuler.f90:97-147, :403-457, :15-75): x, y, z (and the
// reset alfa / psi) stay on the device for the next computeBandJacob; *sorfail != 0 when a line
// could not be interpolated (the reference sets SORFail and the caller rolls back)
static int scb_map(rsg_scb* h, int mode, int* sorfail) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  if (!h->geom_set || !h->map_set) return sfail(RSG_ERR_STATE, "map* before set_geometry / set_map_targets");
  SCK(cudaSetDevice(h->device));
  const int nthe = h->nthe, npsi = h->npsi, nzeta = h->nzeta;
  SCK(cudaMemsetAsync(h->d_fail, 0, sizeof(int), h->st));
  SCK(cudaEventRecord(h->e0, h->st));
  constexpr int LPB = 8;                                    // lines per CTA (8 warps): 64-byte runs of the tile loads
  const bool serial = getenv("RSG_SCB_MAP_SERIAL") != nullptr;   // round 1's thread-per-line kernel (A/B)
  if (mode == 0) {
    const int nl = nthe * npsi;
    const size_t sm = sizeof(double) * LPB * 7 * (size_t)(nzeta + 1);
    if (serial) k_scb_map<0><<<nblk(nl, 64), 64, 0, h->st>>>(h->dev, h->d_alphaVal, h->d_mapw, nl, h->d_fail);
    else {
      if (sm > 48 * 1024) SCK(cudaFuncSetAttribute(k_scb_map_w<0, LPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_scb_map_w<0, LPB><<<nblk(nl, LPB), 32 * LPB, sm, h->st>>>(h->dev, h->d_alphaVal, nl, h->d_fail);
    }
  } else if (mode == 1) {
    const int nl = nthe * (nzeta - 1);
    const size_t sm = sizeof(double) * LPB * 7 * (size_t)npsi;
    if (serial) k_scb_map<1><<<nblk(nl, 64), 64, 0, h->st>>>(h->dev, h->d_psiVal, h->d_mapw, nl, h->d_fail);
    else {
      if (sm > 48 * 1024) SCK(cudaFuncSetAttribute(k_scb_map_w<1, LPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_scb_map_w<1, LPB><<<nblk(nl, LPB), 32 * LPB, sm, h->st>>>(h->dev, h->d_psiVal, nl, h->d_fail);
    }
  } else {
    const int nl = npsi * (nzeta - 1);
    const size_t sm = sizeof(double) * LPB * 7 * (size_t)nthe;
    if (serial) k_scb_map<2><<<nblk(nl, 64), 64, 0, h->st>>>(h->dev, h->d_chiVal, h->d_mapw, nl, h->d_fail);
    else {
      if (sm > 48 * 1024) SCK(cudaFuncSetAttribute(k_scb_map_w<2, LPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      k_scb_map_w<2, LPB><<<nblk(nl, LPB), 32 * LPB, sm, h->st>>>(h->dev, h->d_chiVal, nl, h->d_fail);
    }
  }
  SCKL();
  h->launches++;
  SCK(cudaEventRecord(h->e1, h->st));
  int f = 0;
  SCK(cudaMemcpyAsync(&f, h->d_fail, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  float ms = 0.f;
  SCK(cudaEventElapsedTime(&ms, h->e0, h->e1));
  h->last_ms = ms;
  h->band_done = false;
  if (sorfail) *sorfail = f != 0;
  return RSG_OK;
}
// Anisotropic branch of `pressure` from the normalised equatorial pressures pperEq, pparEq
// (npsi, nzeta+1) on (src/ModScbRun.f90:1087-1175): everything that is 3-D is produced on the device
// from bf / bsq of the last computeBandJacob -- pper, ppar, sigma, tau, the Steffen derivatives of
// pper and bsq and their Euler-potential forms.  The host keeps the part before it (RAM pressures
// interpolated to the SCB equatorial points, smoothing) and uploads two 2-D arrays instead of fifteen
// 3-D ones.
static int scb_pressure_aniso_core(rsg_scb* h, const double* d_pe, const double* d_pa, int iLossCone, int iReduceAnisotropy) {
  const int nthe = h->nthe, npsi = h->npsi, nzeta = h->nzeta;
  SCK(cudaEventRecord(h->e0, h->st));
  const dim3 g(nblk(nthe, 128), npsi, nzeta);
  k_scb_press_aniso<<<g, 128, 0, h->st>>>(h->dev, d_pe, d_pa, h->d_tau, iLossCone, iReduceAnisotropy, (nthe + 1) / 2 - 1);
  SCKL();
  k_scb_derivs<<<g, 128, 0, h->st>>>(h->dev, h->dev.pper, h->dev.dPT, h->dev.dPR, h->dev.dPZ);
  SCKL();
  k_scb_derivs<<<g, 128, 0, h->st>>>(h->dev, h->dev.bsq, h->dev.dBT, h->dev.dBR, h->dev.dBZ);
  SCKL();
  k_scb_press_scale<<<g, 128, 0, h->st>>>(h->dev);
  SCKL();
  h->launches += 4;
  SCK(cudaEventRecord(h->e1, h->st));
  SCK(cudaStreamSynchronize(h->st));
  float ms = 0.f;
  SCK(cudaEventElapsedTime(&ms, h->e0, h->e1));
  h->last_ms = ms;
  h->isotropy = 0;
  h->press_set = true;
  return RSG_OK;
}
static int scb_press_buffers(rsg_scb* h) {
  const size_t n2 = (size_t)h->npsi * (h->nzeta + 1);
  if (!h->d_peq) {
    SRET(h->dalloc(&h->d_peq, 2 * n2, nullptr));
    SRET(h->dalloc(&h->d_tau, (size_t)h->nthe * h->npsi * (h->nzeta + 1), "tau"));
  }
  return RSG_OK;
}
int rsg_scb_pressure_aniso(rsg_scb* h, const double* pperEq, const double* pparEq, int iLossCone, int iReduceAnisotropy) {
  if (!h || !pperEq || !pparEq) return sfail(RSG_ERR_ARG, "null argument");
  if (iLossCone != 1 && iLossCone != 2) return sfail(RSG_ERR_ARG, "iLossCone must be 1 or 2");
  if (!h->grid_set || !h->geom_set) return sfail(RSG_ERR_STATE, "pressure before set_grid / set_geometry");
  SCK(cudaSetDevice(h->device));
  const size_t n2 = (size_t)h->npsi * (h->nzeta + 1);
  SRET(scb_press_buffers(h));
  SCK(cudaMemcpyAsync(h->d_peq, pperEq, n2 * sizeof(double), cudaMemcpyHostToDevice, h->st));
  SCK(cudaMemcpyAsync(h->d_peq + n2, pparEq, n2 * sizeof(double), cudaMemcpyHostToDevice, h->st));
  return scb_pressure_aniso_core(h, h->d_peq, h->d_peq + n2, iLossCone, iReduceAnisotropy);
}

// ---- `pressure`, 2-D front end on the device (src/ModScbRun.f90:838-1086; scb_press_front.cuh) ------------------
// set_ram_pressure, once per scb_run (the RAM pressures do not change while SCB iterates): PPerT, PParT (nS,NR,NT) of
// ram_run / ANISCH, scb[nS] = species%SCB, LZ(NR+1), PHI(NT); PressMode 0 SKD | 1 ROE | 2 EXT | 3 FLT (ModScbParams
// default SKD), iSm2 0 none | 1 SavGol7 | 3 Gaussian | 4 both (default 4), SavGolIters (default 11).
int rsg_scb_set_ram_pressure(rsg_scb* h, int nS, int NR, int NT, const double* PPerT, const double* PParT, const int* scb,
                             const double* LZ, const double* PHI, int PressMode, int iSm2, int SavGolIters) {
  if (!h || !PPerT || !PParT || !scb || !LZ || !PHI) return sfail(RSG_ERR_ARG, "null argument");
  if (nS < 1 || NR < 10 || NT < 8) return sfail(RSG_ERR_ARG, "bad RAM dimensions");
  if (PressMode < 0 || PressMode > 3) return sfail(RSG_ERR_UNSUPPORTED, "PressMode must be SKD (0), ROE (1), EXT (2) or FLT (3); BAT needs the SWMF coupler");
  if (iSm2 != 0 && iSm2 != 1 && iSm2 != 3 && iSm2 != 4) return sfail(RSG_ERR_UNSUPPORTED, "iSm2 = 2 (GSL B-spline fit) is not available on the device");
  SCK(cudaSetDevice(h->device));
  SCK(cudaStreamSynchronize(h->st));
  const int nXRaw = NR - 1, nAz = NT, nX = NR + 2 * (int)std::floor(1.5 / (5. / NR));
  if (nXRaw < 7 || nX < 9 || nAz < 9) return sfail(RSG_ERR_ARG, "RAM grid too small for the 7-point filter");
  const size_t n = (size_t)nX * nAz;
  std::vector<double> radExt(nX), rad2(nX), azim(nAz), fac(nX, 0.0), w(81);
  for (int j1 = 1; j1 <= nXRaw; ++j1) radExt[j1 - 1] = LZ[j1];
  for (int j1 = nXRaw + 1; j1 <= nX; ++j1)
    radExt[j1 - 1] = radExt[nXRaw - 1] + (double)(j1 - nXRaw) * (radExt[nXRaw - 1] - radExt[0]) / ((double)(nXRaw - 1));
  for (int k1 = 0; k1 < nAz; ++k1) azim[k1] = ((PHI[k1] * 12 / PI_D) * 360. / 24) * PI_D / 180.;
  for (int j1 = 0; j1 < nX; ++j1) {
    const double r = radExt[j1];
    rad2[j1] = r * r;
    fac[j1] = (PressMode == 0) ? 89. * std::exp(-0.59 * r) + 8.9 * std::pow(r, -1.53)
                               : 8.4027 * std::exp(-1.7845 * r) + 90.3150 * std::exp(-0.7659 * r);
  }
  {
    double sum = 0.0;
    for (int j = -4; j <= 4; ++j)
      for (int i = -4; i <= 4; ++i) {
        const double x = i, y = j;
        const double v = 2.0 * std::exp(-0.5 * (x * x + y * y) / 1.0);
        w[(i + 4) + 9 * (j + 4)] = v;
        sum += v;
      }
    for (double& v : w) v = v / sum;
  }
  // (re)allocate: the RAM grid may differ between calls only if the handle is rebuilt; sizes are fixed afterwards
  if (!h->d_rad2 || h->pf_nX != nX || h->pf_nAz != nAz) {
    SRET(h->dalloc(&h->d_rad2, nX, nullptr));
    SRET(h->dalloc(&h->d_azim, nAz, nullptr));
    SRET(h->dalloc(&h->d_rper, n, nullptr));
    SRET(h->dalloc(&h->d_rpar, n, nullptr));
    h->pf_nX = nX; h->pf_nAz = nAz;
  }
  double *d_in = nullptr, *d_tab = nullptr, *d_tmp = nullptr;
  int* d_scb = nullptr;
  const size_t nin = (size_t)nS * NR * NT;
  SCK(cudaMalloc((void**)&d_in, 2 * nin * sizeof(double)));
  SCK(cudaMalloc((void**)&d_tab, (2 * (size_t)nX + 81) * sizeof(double)));
  SCK(cudaMalloc((void**)&d_tmp, 6 * n * sizeof(double)));
  SCK(cudaMalloc((void**)&d_scb, nS * sizeof(int)));
  SCK(cudaMemcpy(d_in, PPerT, nin * sizeof(double), cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(d_in + nin, PParT, nin * sizeof(double), cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(d_tab, radExt.data(), nX * sizeof(double), cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(d_tab + nX, fac.data(), nX * sizeof(double), cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(d_tab + 2 * nX, w.data(), 81 * sizeof(double), cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(d_scb, scb, nS * sizeof(int), cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(h->d_rad2, rad2.data(), nX * sizeof(double), cudaMemcpyHostToDevice));
  SCK(cudaMemcpy(h->d_azim, azim.data(), nAz * sizeof(double), cudaMemcpyHostToDevice));
  PressRawArgs A{};
  A.nS = nS; A.NR = NR; A.NT = NT; A.nX = nX; A.nXRaw = nXRaw; A.nAz = nAz; A.mode = PressMode; A.iSm2 = iSm2; A.iters = SavGolIters;
  A.PPerT = d_in; A.PParT = d_in + nin; A.scb = d_scb; A.radExt = d_tab; A.fac = d_tab + nX; A.w = d_tab + 2 * nX;
  A.per = h->d_rper; A.par = h->d_rpar; A.t0 = d_tmp; A.t1 = d_tmp + 2 * n; A.t2 = d_tmp + 4 * n;
  k_scb_press_raw<<<1, 256, 0, h->st>>>(A);
  SCKL();
  h->launches++;
  SCK(cudaStreamSynchronize(h->st));
  cudaFree(d_in); cudaFree(d_tab); cudaFree(d_tmp); cudaFree(d_scb);
  h->pf_set = true;
  return RSG_OK;
}
// the extended, smoothed RAM pressures and their axes (diagnostics / tests): rad2(nX), azim(nAz), per, par (nX,nAz)
int rsg_scb_get_ram_pressure(rsg_scb* h, int* nX, int* nAz, double* rad2, double* azim, double* per, double* par) {
  if (!h || !nX || !nAz) return sfail(RSG_ERR_ARG, "null argument");
  if (!h->pf_set) return sfail(RSG_ERR_STATE, "no RAM pressures on the device");
  *nX = h->pf_nX; *nAz = h->pf_nAz;
  SCK(cudaSetDevice(h->device));
  const size_t n = (size_t)h->pf_nX * h->pf_nAz;
  if (rad2) SCK(cudaMemcpy(rad2, h->d_rad2, h->pf_nX * sizeof(double), cudaMemcpyDeviceToHost));
  if (azim) SCK(cudaMemcpy(azim, h->d_azim, h->pf_nAz * sizeof(double), cudaMemcpyDeviceToHost));
  if (per) SCK(cudaMemcpy(per, h->d_rper, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (par) SCK(cudaMemcpy(par, h->d_rpar, n * sizeof(double), cudaMemcpyDeviceToHost));
  return RSG_OK;
}
// `pressure` (anisotropic, RAM-coupled) entirely on the device: front end + rsg_scb_pressure_aniso's 3-D tail.
// pperEq / pparEq (npsi, nzeta+1; may be NULL) return the normalised equatorial pressures.
static int scb_pressure_front(rsg_scb* h, int iLossCone, int iReduceAnisotropy, double* pperEq, double* pparEq) {
  if (!h->pf_set) return sfail(RSG_ERR_STATE, "pressure front end before rsg_scb_set_ram_pressure");
  if (iLossCone != 1 && iLossCone != 2) return sfail(RSG_ERR_ARG, "iLossCone must be 1 or 2");
  const int npsi = h->npsi, nzeta = h->nzeta;
  const size_t n2 = (size_t)npsi * (nzeta + 1);
  SRET(scb_press_buffers(h));
  if (!h->d_pf) SRET(h->dalloc(&h->d_pf, 4 * n2, nullptr));
  k_scb_gather_eq<<<nblk((long long)n2, 128), 128, 0, h->st>>>(h->dev, (h->nthe + 1) / 2 - 1, h->d_peq);
  SCKL();
  k_scb_press_eq<<<nblk(nzeta + 1, 64), 64, 0, h->st>>>(npsi, nzeta, h->d_peq, h->pf_nX, h->pf_nAz, h->d_rad2, h->d_azim, h->d_rper,
                                                        h->d_rpar, h->pnormal, h->d_pf);
  SCKL();
  k_scb_press_wrap<<<nblk((long long)(2 * n2), 128), 128, 0, h->st>>>(npsi, nzeta, h->pnormal, h->d_pf);
  SCKL();
  h->launches += 3;
  if (pperEq) SCK(cudaMemcpyAsync(pperEq, h->d_pf + 2 * n2, n2 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  if (pparEq) SCK(cudaMemcpyAsync(pparEq, h->d_pf + 3 * n2, n2 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  return scb_pressure_aniso_core(h, h->d_pf + 2 * n2, h->d_pf + 3 * n2, iLossCone, iReduceAnisotropy);
}
int rsg_scb_pressure_front(rsg_scb* h, int iLossCone, int iReduceAnisotropy, double* pperEq, double* pparEq) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  if (!h->grid_set || !h->geom_set) return sfail(RSG_ERR_STATE, "pressure before set_grid / set_geometry");
  SCK(cudaSetDevice(h->device));
  return scb_pressure_front(h, iLossCone, iReduceAnisotropy, pperEq, pparEq);
}

// FLC_Radius (src/ModRamLoss.f90:176-336) from the geometry and field of the last computeBandJacob, both resident:
// r_curvEq, zeta1Eq, zeta2Eq (nR,nT) for PARA_FLC (rsg_para_flc).  radRaw(1:nR), azimRaw(1:nT) as the reference holds them
// (src/ModRamEField.f90:100-107); REarth = 6.4e6 (src/ModScbMain.f90:15).  The caller keeps the "every Dt_bc" gate (:207).
int rsg_scb_flc_radius(rsg_scb* h, int nR, int nT, const double* radRaw, const double* azimRaw, double REarth, double* r_curvEq,
                       double* zeta1Eq, double* zeta2Eq) {
  if (!h || !radRaw || !azimRaw || !r_curvEq || !zeta1Eq || !zeta2Eq) return sfail(RSG_ERR_ARG, "null argument");
  if (nR < 2 || nT < 2) return sfail(RSG_ERR_ARG, "bad RAM dimensions");
  if (!h->geom_set || !h->band_done) return sfail(RSG_ERR_STATE, "FLC_Radius needs computeBandJacob on the current geometry");
  const int nthe = h->nthe, npsi = h->npsi, nzeta = h->nzeta, m1 = nzeta - 1;
  const int ie = (nthe + 1) / 2 - 1;
  if (ie + 1 < 2 || ie + 1 > nthe - 3 || (long long)npsi * m1 < 9) return sfail(RSG_ERR_ARG, "grid too small for FLC_Radius");
  SCK(cudaSetDevice(h->device));
  const size_t ne = (size_t)npsi * m1, nq = (size_t)nR * nT;
  const size_t smem = 2 * ne * sizeof(double);
  if (smem > 200 * 1024) return sfail(RSG_ERR_UNSUPPORTED, "equatorial plane too large for the shared-memory candidate set");
  double* buf = nullptr;
  SCK(cudaMalloc((void**)&buf, (5 * ne + 5 * nq) * sizeof(double)));
  std::vector<double> q(2 * nq, 0.0);
  const double PI = 3.1415926535897932384626433832795;
  for (int i = 2; i <= nR; ++i)
    for (int j = 1; j <= nT - 1; ++j) {
      q[(size_t)(i - 1) + (size_t)nR * (j - 1)] = radRaw[i - 1] * std::cos(azimRaw[j - 1] * 2 * PI / 24. - PI);
      q[nq + (size_t)(i - 1) + (size_t)nR * (j - 1)] = radRaw[i - 1] * std::sin(azimRaw[j - 1] * 2 * PI / 24. - PI);
    }
  FlcArgs A{};
  A.nthe = nthe; A.npsi = npsi; A.nzeta = nzeta; A.nR = nR; A.nT = nT; A.ie = ie; A.bnormal = h->bnormal; A.REarth = REarth;
  A.x = h->dev.x; A.y = h->dev.y; A.z = h->dev.z; A.bx = h->dev.Bx; A.by = h->dev.By; A.bz = h->dev.Bz;
  A.xe = buf; A.ye = buf + ne; A.rc = buf + 2 * ne; A.z1 = buf + 3 * ne; A.z2 = buf + 4 * ne;
  double* dq = buf + 5 * ne;
  A.qx = dq; A.qy = dq + nq; A.rcEq = dq + 2 * nq; A.z1Eq = dq + 3 * nq; A.z2Eq = dq + 4 * nq;
  SCK(cudaMemcpyAsync(dq, q.data(), 2 * nq * sizeof(double), cudaMemcpyHostToDevice, h->st));
  SCK(cudaMemsetAsync(dq + 2 * nq, 0, 3 * nq * sizeof(double), h->st));
  SCK(cudaEventRecord(h->e0, h->st));
  k_flc_curv<<<nblk((long long)ne, 128), 128, 0, h->st>>>(A);
  SCKL();
  if (smem > 48 * 1024) SCK(cudaFuncSetAttribute(k_flc_nn9, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_flc_nn9<<<std::max(1, std::min(148, (int)(((size_t)(nR - 1) * (nT - 1) + 7) / 8))), 256, smem, h->st>>>(A);
  SCKL();
  k_flc_edges<<<nblk((long long)nq, 128), 128, 0, h->st>>>(A);
  SCKL();
  h->launches += 3;
  SCK(cudaEventRecord(h->e1, h->st));
  SCK(cudaMemcpyAsync(r_curvEq, A.rcEq, nq * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaMemcpyAsync(zeta1Eq, A.z1Eq, nq * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaMemcpyAsync(zeta2Eq, A.z2Eq, nq * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  float ms = 0.f;
  SCK(cudaEventElapsedTime(&ms, h->e0, h->e1));
  h->last_ms = ms;
  cudaFree(buf);
  return RSG_OK;
}

// ---- glue of the outer iteration (src/ModScbRun.f90:232-262, 418-440) ------------------------
// snapshot slots hold device copies of a named (nthe,npsi,nzeta+1) field: alfaSav1 / alphaPrev /
// xPrev ... of the reference.  slot 0..3.
static int scb_slot(rsg_scb* h, const char* name, int slot, double** field, double** snap, size_t* n) {
  if (!h || !name) return sfail(RSG_ERR_ARG, "null argument");
  if (slot < 0 || slot > 3) return sfail(RSG_ERR_ARG, "snapshot slot must be 0..3");
  auto it = h->arr.find(name);
  if (it == h->arr.end()) return sfail(RSG_ERR_ARG, std::string("unknown array ") + name);
  const std::string key = std::string(name) + "#" + std::to_string(slot);
  auto sn = h->snaps.find(key);
  if (sn == h->snaps.end()) {
    double* p = nullptr;
    SRET(h->dalloc(&p, it->second.second, nullptr));
    sn = h->snaps.emplace(key, p).first;
  }
  *field = it->second.first; *snap = sn->second; *n = it->second.second;
  return RSG_OK;
}
int rsg_scb_snapshot(rsg_scb* h, const char* name, int slot) {
  double *f, *sn; size_t n;
  SRET(scb_slot(h, name, slot, &f, &sn, &n));
  SCK(cudaSetDevice(h->device));
  SCK(cudaMemcpyAsync(sn, f, n * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  return RSG_OK;
}
int rsg_scb_restore(rsg_scb* h, const char* name, int slot) {
  double *f, *sn; size_t n;
  SRET(scb_slot(h, name, slot, &f, &sn, &n));
  SCK(cudaSetDevice(h->device));
  SCK(cudaMemcpyAsync(f, sn, n * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
  if (!std::strcmp(name, "x") || !std::strcmp(name, "y") || !std::strcmp(name, "z")) h->band_done = false;
  return RSG_OK;
}
// field = snap(slot_new)*blend + snap(slot_sav)*(1 - blend)       (src/ModScbRun.f90:236, :422)
int rsg_scb_blend(rsg_scb* h, const char* name, int slot_new, int slot_sav, double blend) {
  double *f, *a, *b; size_t n;
  SRET(scb_slot(h, name, slot_new, &f, &a, &n));
  SRET(scb_slot(h, name, slot_sav, &f, &b, &n));
  SCK(cudaSetDevice(h->device));
  k_scb_blend<<<(unsigned)((n + 255) / 256), 256, 0, h->st>>>(f, a, b, blend, n);
  SCKL();
  h->launches++;
  return RSG_OK;
}
// MINVAL(jacobian(2:nthe-1,2:npsi-1,2:nzeta)) of the last computeBandJacob (:248); -1e300 if a NaN is present
int rsg_scb_min_jacobian(rsg_scb* h, double* minjac) {
  if (!h || !minjac) return sfail(RSG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(h->device));
  const int nb = h->nzeta - 1;
  k_scb_minjac<<<nb, 256, 0, h->st>>>(h->dev, h->d_part);
  SCKL();
  h->launches++;
  std::vector<double> part(nb);
  SCK(cudaMemcpyAsync(part.data(), h->d_part, nb * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  double m = part[0];
  for (int q = 1; q < nb; ++q) m = part[q] < m ? part[q] : m;
  *minjac = m;
  return RSG_OK;
}

int rsg_scb_map_alpha(rsg_scb* h, int* sorfail) { return scb_map(h, 0, sorfail); }
int rsg_scb_map_psi(rsg_scb* h, int* sorfail) { return scb_map(h, 1, sorfail); }
int rsg_scb_map_theta(rsg_scb* h, int* sorfail) { return scb_map(h, 2, sorfail); }

// ---------------------------------------------------------------------------------------------
// scb_run, src/ModScbRun.f90:149-440 (method = 2, iAMR = 0): the whole outer iteration of the
// Euler-potential solve with every 3-D array resident.  Only `pressure`'s 2-D front end is the
// caller's: the callback receives the equatorial foot points of the field lines and returns the
// normalised equatorial pressures (npsi, nzeta+1); everything downstream (field-line mapping,
// derivatives, coefficients, SOR, re-gridding, blending, Jacobian test, convergence norms) is here.
// ---------------------------------------------------------------------------------------------
static int scb_run_pressure(rsg_scb* h, const rsg_scb_run_params* p, rsg_scb_pressure_fn fn, void* user, std::vector<double>& eq,
                            std::vector<double>& pe, std::vector<double>& pa) {
  const size_t n2 = (size_t)h->npsi * (h->nzeta + 1);
  if (!fn) return scb_pressure_front(h, p->iLossCone, p->iReduceAnisotropy, nullptr, nullptr);   // no host hop at all
  SRET(scb_press_buffers(h));
  k_scb_gather_eq<<<nblk((long long)n2, 128), 128, 0, h->st>>>(h->dev, (h->nthe + 1) / 2 - 1, h->d_peq);
  SCKL();
  h->launches++;
  SCK(cudaMemcpyAsync(eq.data(), h->d_peq, 2 * n2 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  if (fn(user, h->npsi, h->nzeta + 1, eq.data(), eq.data() + n2, pe.data(), pa.data()) != 0)
    return sfail(RSG_ERR_ARG, "the pressure callback reported a failure");
  return rsg_scb_pressure_aniso(h, pe.data(), pa.data(), p->iLossCone, p->iReduceAnisotropy);
}

int rsg_scb_run(rsg_scb* h, const rsg_scb_run_params* p, rsg_scb_pressure_fn pressure, void* user, rsg_scb_run_result* out) {
  if (!h || !p || !out) return sfail(RSG_ERR_ARG, "null argument");
  if (!h->grid_set || !h->geom_set || !h->map_set) return sfail(RSG_ERR_STATE, "scb_run before set_grid / set_geometry / set_map_targets");
  if (!pressure && !h->pf_set) return sfail(RSG_ERR_STATE, "scb_run without a pressure callback needs rsg_scb_set_ram_pressure");
  SCK(cudaSetDevice(h->device));
  const size_t n2 = (size_t)h->npsi * (h->nzeta + 1);
  std::vector<double> eq(2 * n2), pe(n2), pa(n2);
  *out = rsg_scb_run_result{};
  int fail = 0, f = 0;
  const char* xyz[3] = {"x", "y", "z"};
  // xStart ... alphaStart (:134-143): slot 2
  for (const char* n : xyz) SRET(rsg_scb_snapshot(h, n, 2));
  SRET(rsg_scb_snapshot(h, "alfa", 2));
  SRET(rsg_scb_snapshot(h, "psi", 2));
  SRET(rsg_scb_bandjacob(h, &f)); fail |= f;                                                     // :148-150
  SRET(scb_run_pressure(h, p, pressure, user, eq, pe, pa));
  SRET(rsg_scb_convergence(h, &out->normDiffStart, &out->normJxBStart, &out->normGradPStart, &f)); fail |= f;
  double blendAlpha = p->blendInitial, blendPsi = p->blendInitial;                               // :178-180
  double errorAlpha = 0.0, errorPsi = 0.0;
  int iteration = 1, iConvGlobal = 0;
  SRET(rsg_scb_snapshot(h, "alfa", 1));                                                          // alfaSav1, psiSav1 (:206-207)
  SRET(rsg_scb_snapshot(h, "psi", 1));
  while (!fail) {                                                                                // Outeriters (:209)
    // ---- equation 1 (:213-262)
    SRET(rsg_scb_bandjacob(h, &f)); if (f) { fail = 1; break; }
    SRET(scb_run_pressure(h, p, pressure, user, eq, pe, pa));
    SRET(rsg_scb_metrica(h));
    SRET(rsg_scb_newk(h));
    blendAlpha = std::min(std::max(blendAlpha, p->blendMin), p->blendMax);
    double sb, sdb, dmx;
    int nis;
    SRET(rsg_scb_iterate_alpha(h, p->InConAlpha, p->nimax, p->theChange, p->psiChange, p->ordering, &nis, &sb, &sdb, &dmx, &f, nullptr));
    if (f) { fail = 1; break; }
    out->nisaveAlpha = nis; out->sumbAlpha = sb; out->sumdbAlpha = sdb;
    errorAlpha = dmx;
    for (const char* n : xyz) SRET(rsg_scb_snapshot(h, n, 0));                                   // xPrev, alphaPrev
    SRET(rsg_scb_snapshot(h, "alfa", 0));
    for (;;) {                                                                                   // Move_points_in_alpha_theta
      SRET(rsg_scb_blend(h, "alfa", 0, 1, blendAlpha));
      SRET(rsg_scb_map_alpha(h, &f)); if (f) { fail = 1; break; }
      SRET(rsg_scb_map_theta(h, &f)); if (f) { fail = 1; break; }
      SRET(rsg_scb_bandjacob(h, &f)); fail |= f;
      double mj;
      SRET(rsg_scb_min_jacobian(h, &mj));
      if (mj < 0.0) {
        for (const char* n : xyz) SRET(rsg_scb_restore(h, n, 0));
        blendAlpha = p->damp * blendAlpha;
        out->blendRetries++;
        if (blendAlpha < p->blendMin) { fail = 1; break; }
        continue;
      }
      break;
    }
    if (fail) break;
    // ---- equation 2 (:285-349)
    SRET(rsg_scb_bandjacob(h, &f)); if (f) { fail = 1; break; }
    SRET(scb_run_pressure(h, p, pressure, user, eq, pe, pa));
    { double a, b, c; SRET(rsg_scb_convergence(h, &a, &b, &c, &f)); fail |= f; }
    SRET(rsg_scb_metric(h));
    SRET(rsg_scb_newj(h));
    blendPsi = std::min(std::max(blendPsi, p->blendMin), p->blendMax);
    SRET(rsg_scb_iterate_psi(h, p->InConPsi, p->nimax, p->theChange, p->psiChange, p->ordering, &nis, &sb, &sdb, &dmx, &f, nullptr));
    if (f) { fail = 1; break; }
    out->nisavePsi = nis; out->sumbPsi = sb; out->sumdbPsi = sdb;
    errorPsi = dmx;
    for (const char* n : xyz) SRET(rsg_scb_snapshot(h, n, 0));
    SRET(rsg_scb_snapshot(h, "psi", 0));
    for (;;) {                                                                                   // Move_points_in_psi_theta
      SRET(rsg_scb_blend(h, "psi", 0, 1, blendPsi));
      SRET(rsg_scb_map_psi(h, &f)); if (f) { fail = 1; break; }
      SRET(rsg_scb_map_theta(h, &f)); if (f) { fail = 1; break; }
      SRET(rsg_scb_bandjacob(h, &f)); fail |= f;
      double mj;
      SRET(rsg_scb_min_jacobian(h, &mj));
      if (mj < 0.0) {
        for (const char* n : xyz) SRET(rsg_scb_restore(h, n, 0));
        blendPsi = p->damp * blendPsi;
        out->blendRetries++;
        if (blendPsi < p->blendMin) { fail = 1; break; }
        continue;
      }
      break;
    }
    if (fail) break;
    // ---- convergence control (:363-378); SCBIterNeeded = 1 and outDistance > convDistance are constants there
    if (errorAlpha < p->decreaseConvAlpha && errorPsi < p->decreaseConvPsi) iConvGlobal = 1;
    if ((iteration < p->numit && iConvGlobal == 0) || iteration < p->MinSCBIterations) { iteration++; continue; }
    break;
  }
  out->iterations = iteration;
  out->iConvGlobal = iConvGlobal;
  out->blendAlpha = blendAlpha; out->blendPsi = blendPsi;
  out->errorAlpha = errorAlpha; out->errorPsi = errorPsi;
  out->SORFail = fail ? 1 : 0;
  if (fail) {                                                                                    // :397-413: back to the start
    for (const char* n : xyz) SRET(rsg_scb_restore(h, n, 2));
    SRET(rsg_scb_restore(h, "alfa", 2));
    SRET(rsg_scb_restore(h, "psi", 2));
    return RSG_OK;
  }
  SRET(rsg_scb_bandjacob(h, &f));                                                                // :427-429
  SRET(scb_run_pressure(h, p, pressure, user, eq, pe, pa));
  SRET(rsg_scb_convergence(h, &out->normDiff, &out->normJxB, &out->normGradP, &f));
  return RSG_OK;
}

// ---- launch sequences of the three computehI blocks, shared by the stateless calls and the resident handle (rsg_hi) ----
static int hi_launch_lines(const HiArgs& A, cudaStream_t st) {
  const int threads = std::min(256, ((A.nPa + 31) / 32) * 32);
  const size_t smem = (size_t)(4 * A.nthe + 4 * A.nPa) * sizeof(double) + 2 * (size_t)A.nPa * sizeof(int);
  if (smem > 200 * 1024) return sfail(RSG_ERR_UNSUPPORTED, "field line does not fit shared memory");
  if (smem > 48 * 1024) SCK(cudaFuncSetAttribute(k_hi_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_hi_lines<<<(unsigned)(A.nR * A.nT), threads, smem, st>>>(A);
  SCK(cudaGetLastError());
  return RSG_OK;
}
static void hi_gauss_weights(double* w) {                             // gaussian_kernel(1.0), srcExternal/gaussian_filter.f90:19-56
  double sum = 0.0;
  for (int j = -4; j <= 4; j++)
    for (int i = -4; i <= 4; i++) {
      const double x = i, y = j;
      const double v = 2.0 * std::exp(-0.5 * (x * x + y * y) / 1.0);
      w[(i + 4) + 9 * (j + 4)] = v;
      sum += v;
    }
  for (int q = 0; q < 81; q++) w[q] = w[q] / sum;
}
static int hi_launch_tail(const HiTailArgs& A, cudaStream_t st) {
  const int ncol = A.nT * A.nPa;
  const size_t nl = (size_t)A.nR * A.nT, n3 = nl * A.nPa;
  const int lthreads = std::min(256, ((A.nPa + 31) / 32) * 32);
  const size_t lsmem = (size_t)8 * A.nPa * sizeof(double);
  if (lsmem > 48 * 1024) return sfail(RSG_ERR_UNSUPPORTED, "pitch-angle line does not fit shared memory");
  k_hi_tail_cols<<<nblk(ncol, 128), 128, 0, st>>>(A);
  SCK(cudaGetLastError());
  k_hi_tail_lines<<<(unsigned)nl, lthreads, lsmem, st>>>(A);
  SCK(cudaGetLastError());
  if (A.smooth) {
    k_hi_smooth<<<nblk((long long)n3, 128), 128, 0, st>>>(A);
    SCK(cudaGetLastError());
  }
  k_hi_fill<<<nblk(ncol, 128), 128, 0, st>>>(A);
  SCK(cudaGetLastError());
  return RSG_OK;
}
static int hi_launch_convert(const HiConvArgs& A, cudaStream_t st) {
  const size_t nl = (size_t)A.nR * A.nT;
  const size_t smem = 2 * (size_t)A.npsi * (A.nzeta - 1) * sizeof(double);
  if (smem > 200 * 1024) return sfail(RSG_ERR_UNSUPPORTED, "SCB surface does not fit shared memory");
  SCK(cudaFuncSetAttribute(k_hi_nn9<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SCK(cudaFuncSetAttribute(k_hi_nn9<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int q0 = std::max(1, std::min(148, (int)((nl + 7) / 8)));   // one point set, queries spread over the SMs
  k_hi_nn9<0><<<dim3(1, q0), 256, smem, st>>>(A);
  SCK(cudaGetLastError());
  const int q1 = std::max(1, std::min((int)((nl + 7) / 8), (3 * 148) / A.nthe));   // one wave: 3 resident CTAs (69 KB each) per SM
  k_hi_nn9<1><<<dim3(A.nthe, q1), 256, smem, st>>>(A);
  SCK(cudaGetLastError());
  return RSG_OK;
}
// xo, yo of the RAM equatorial points and alphaRAM (src/ModRamScb.f90:258-259, :283-284; cos / sin of the host libm)
static void hi_query_tables(int nR, int nT, const double* Lz, const double* MLT, std::vector<double>& qx, std::vector<double>& qy,
                            std::vector<double>& al) {
  qx.resize((size_t)nR * nT); qy.resize((size_t)nR * nT); al.resize(nT);
  const double twopi = 2.0 * PI_D;
  for (int j = 0; j < nT; j++) {
    for (int i = 0; i < nR; i++) {
      qx[i + (size_t)nR * j] = Lz[i + 1] * std::cos(MLT[j] * 2.0 * PI_D / 24.0 - PI_D);
      qy[i + (size_t)nR * j] = Lz[i + 1] * std::sin(MLT[j] * 2.0 * PI_D / 24.0 - PI_D);
    }
    al[j] = MLT[j] * PI_D / 12.0 + PI_D;
    if (al[j] > twopi) al[j] = al[j] - twopi;
  }
}

// computehI's integral block (src/ModRamScb.f90:372-410): I_cart, H_cart, HDens_cart, bZEq_Cart of every RAM
// field line from the traced lines (xRAM, yRAM, zRAM, bRAM, density: (nthe,nR,nT)); HDens_cart is in/out
// (lines with outsideMGNP != 0 keep their value, I_cart / H_cart / bZEq_Cart are zero there).
// Stateless: buffers live for the call (the arrays are a few MB and the call runs once per SCB update).
int rsg_hI_integrals(int device, int nthe, int nR, int nT, int nPa, int nThetaEquator, double bnormal, const double* chiVal,
                     const double* mu, const double* xRAM, const double* yRAM, const double* zRAM, const double* bRAM,
                     const double* density, const int* outsideMGNP, double* I_cart, double* H_cart, double* HDens_cart,
                     double* bZEq_cart, double* ms) {
  if (!chiVal || !mu || !xRAM || !yRAM || !zRAM || !bRAM || !density || !outsideMGNP || !I_cart || !H_cart || !HDens_cart || !bZEq_cart)
    return sfail(RSG_ERR_ARG, "null argument");
  if (nthe < 3 || nR < 1 || nT < 1 || nPa < 3 || nThetaEquator < 1 || nThetaEquator > nthe)
    return sfail(RSG_ERR_ARG, "bad dimensions");
  SCK(cudaSetDevice(device));
  const size_t n3 = (size_t)nthe * nR * nT, nl = (size_t)nR * nT, no = nl * nPa;
  const size_t nd = 5 * n3 + nthe + nPa + 3 * no + nl;
  double* d = nullptr;
  int* dout = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = RSG_OK;
  auto done = [&](int code, const std::string& msg) {
    if (d) cudaFree(d);
    if (dout) cudaFree(dout);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    return code == RSG_OK ? RSG_OK : sfail(code, msg);
  };
#define HCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return done(RSG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
  HCK(cudaStreamCreate(&st));
  HCK(cudaEventCreate(&e0));
  HCK(cudaEventCreate(&e1));
  HCK(cudaMalloc(&d, nd * sizeof(double)));
  HCK(cudaMalloc(&dout, nl * sizeof(int)));
  HiArgs A;
  A.nthe = nthe; A.nR = nR; A.nT = nT; A.nPa = nPa; A.nThetaEquator = nThetaEquator; A.bnormal = bnormal;
  double* p = d;
  cudaError_t uerr = cudaSuccess;      // first failed upload (checked after the batch)
  auto up = [&](const double* src, size_t n) {
    double* q = p; p += n;
    const cudaError_t e_ = cudaMemcpyAsync(q, src, n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e_ != cudaSuccess && uerr == cudaSuccess) uerr = e_;
    return q;
  };
  A.x = up(xRAM, n3); A.y = up(yRAM, n3); A.z = up(zRAM, n3); A.b = up(bRAM, n3); A.dens = up(density, n3);
  A.chi = up(chiVal, nthe); A.mu = up(mu, nPa);
  A.Icart = p; p += no;
  A.Hcart = p; p += no;
  A.Dcart = up(HDens_cart, no);
  A.bzeq = p;
  HCK(uerr);
  HCK(cudaMemcpyAsync(dout, outsideMGNP, nl * sizeof(int), cudaMemcpyHostToDevice, st));
  A.outside = dout;
  HCK(cudaEventRecord(e0, st));
  if ((rc = hi_launch_lines(A, st)) != RSG_OK) return done(rc, g_serr);
  HCK(cudaEventRecord(e1, st));
  HCK(cudaMemcpyAsync(I_cart, A.Icart, no * sizeof(double), cudaMemcpyDeviceToHost, st));
  HCK(cudaMemcpyAsync(H_cart, A.Hcart, no * sizeof(double), cudaMemcpyDeviceToHost, st));
  HCK(cudaMemcpyAsync(HDens_cart, A.Dcart, no * sizeof(double), cudaMemcpyDeviceToHost, st));
  HCK(cudaMemcpyAsync(bZEq_cart, A.bzeq, nl * sizeof(double), cudaMemcpyDeviceToHost, st));
  HCK(cudaStreamSynchronize(st));
  if (ms) {
    float t = 0.f;
    HCK(cudaEventElapsedTime(&t, e0, e1));
    *ms = t;
  }
#undef HCK
  return done(rc, "");
}

// computehI after the integral block (src/ModRamScb.f90:413-637): outer-boundary scaling, MLT continuity, near-90-degree
// corrections, repairs, Steffen interpolation of h / I onto PAbn, Gaussian smoothing, the RAM variables (FNHS, FNIS,
// BOUNHS, BOUNIS, HDNS, BNES) with their time derivatives, the I = 1 row and the NaN repair.  I_cart, H_cart,
// HDens_cart, bZEq_cart come back as the reference leaves them.  *gslerr != 0: a line could not be interpolated.
int rsg_hI_tail(int device, int nR, int nT, int nPa, double* I_cart, double* H_cart, double* HDens_cart, double* bZEq_cart,
                const int* ScaleAt, const int* outsideMGNP, const double* Lz, const double* PA, const double* PAbn,
                int integral_smooth, double DthI, double* FNHS, double* FNIS, double* BOUNHS, double* BOUNIS, double* HDNS,
                double* BNES, double* dIdt, double* dHdt, double* dIbndt, double* dBdt, int* gslerr, double* ms) {
  if (!I_cart || !H_cart || !HDens_cart || !bZEq_cart || !ScaleAt || !outsideMGNP || !Lz || !PA || !PAbn || !FNHS || !FNIS ||
      !BOUNHS || !BOUNIS || !HDNS || !BNES || !dIdt || !dHdt || !dIbndt || !dBdt)
    return sfail(RSG_ERR_ARG, "null argument");
  if (nR < 3 || nT < 2 || nPa < 5) return sfail(RSG_ERR_ARG, "bad dimensions");
  if (integral_smooth && (nR < 9 || nT < 9)) return sfail(RSG_ERR_ARG, "grid smaller than the 9 x 9 smoothing kernel");
  for (int j = 0; j < nT; j++)
    if (ScaleAt[j] != 0 && (ScaleAt[j] < 3 || ScaleAt[j] > nR)) return sfail(RSG_ERR_ARG, "ScaleAt out of range");
  SCK(cudaSetDevice(device));
  const size_t nl = (size_t)nR * nT, n3 = nl * nPa, nr2 = (size_t)(nR + 1) * nT, nr3 = nr2 * nPa;
  const size_t nd = 13 * n3 + 2 * nl + 8 * nr3 + 2 * nr2 + (nR + 1) + 2 * (size_t)nPa;
  double* d = nullptr;
  int* di = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto done = [&](int code, const std::string& msg) {
    if (d) cudaFree(d);
    if (di) cudaFree(di);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    return code == RSG_OK ? RSG_OK : sfail(code, msg);
  };
#define HCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return done(RSG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
  HCK(cudaStreamCreate(&st));
  HCK(cudaEventCreate(&e0));
  HCK(cudaEventCreate(&e1));
  HCK(cudaMalloc(&d, nd * sizeof(double)));
  HCK(cudaMalloc(&di, (nl + nT + 1) * sizeof(int)));
  HiTailArgs A;
  A.nR = nR; A.nT = nT; A.nPa = nPa; A.smooth = integral_smooth ? 1 : 0; A.DthI = DthI;
  A.bnes1 = 0.32 / (Lz[0] * Lz[0] * Lz[0]) / 1.e4;                                  // :609
  hi_gauss_weights(A.w);
  double* p = d;
  auto take = [&](size_t n) { double* q = p; p += n; return q; };
  cudaError_t uerr = cudaSuccess;      // first failed upload (checked after the batch)
  auto up = [&](const double* src, size_t n) {
    double* q = take(n);
    const cudaError_t e_ = cudaMemcpyAsync(q, src, n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e_ != cudaSuccess && uerr == cudaSuccess) uerr = e_;
    return q;
  };
  A.I0 = up(I_cart, n3); A.H0 = up(H_cart, n3); A.D0 = up(HDens_cart, n3); A.bz0 = up(bZEq_cart, nl);
  A.I1 = take(n3); A.H1 = take(n3); A.D1 = take(n3); A.bz1 = take(nl); A.hI = take(n3); A.iI = take(n3);
  A.I2 = take(n3); A.H2 = take(n3); A.D2 = take(n3); A.hI2 = take(n3); A.iI2 = take(n3);
  A.FNHS = up(FNHS, nr3); A.FNIS = up(FNIS, nr3); A.BOUNHS = up(BOUNHS, nr3); A.BOUNIS = up(BOUNIS, nr3); A.HDNS = up(HDNS, nr3);
  A.dIdt = take(nr3); A.dHdt = take(nr3); A.dIbndt = take(nr3);
  A.BNES = up(BNES, nr2); A.dBdt = take(nr2);
  A.Lz = up(Lz, nR + 1); A.PA = up(PA, nPa); A.PAbn = up(PAbn, nPa);
  HCK(uerr);
  HCK(cudaMemcpyAsync(di, outsideMGNP, nl * sizeof(int), cudaMemcpyHostToDevice, st));
  HCK(cudaMemcpyAsync(di + nl, ScaleAt, nT * sizeof(int), cudaMemcpyHostToDevice, st));
  HCK(cudaMemsetAsync(di + nl + nT, 0, sizeof(int), st));
  A.outside = di; A.ScaleAt = di + nl; A.fail = di + nl + nT;
  HCK(cudaEventRecord(e0, st));
  { const int rc = hi_launch_tail(A, st); if (rc != RSG_OK) return done(rc, g_serr); }
  HCK(cudaEventRecord(e1, st));
  auto down = [&](double* dst, const double* src, size_t n) { return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, st); };
  HCK(down(I_cart, A.smooth ? A.I2 : A.I1, n3));
  HCK(down(H_cart, A.smooth ? A.H2 : A.H1, n3));
  HCK(down(HDens_cart, A.smooth ? A.D2 : A.D1, n3));
  HCK(down(bZEq_cart, A.bz1, nl));
  HCK(down(FNHS, A.FNHS, nr3)); HCK(down(FNIS, A.FNIS, nr3)); HCK(down(BOUNHS, A.BOUNHS, nr3)); HCK(down(BOUNIS, A.BOUNIS, nr3));
  HCK(down(HDNS, A.HDNS, nr3)); HCK(down(dIdt, A.dIdt, nr3)); HCK(down(dHdt, A.dHdt, nr3)); HCK(down(dIbndt, A.dIbndt, nr3));
  HCK(down(BNES, A.BNES, nr2)); HCK(down(dBdt, A.dBdt, nr2));
  int nfail = 0;
  HCK(cudaMemcpyAsync(&nfail, A.fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  HCK(cudaStreamSynchronize(st));
  if (gslerr) *gslerr = nfail;
  if (ms) {
    float t = 0.f;
    HCK(cudaEventElapsedTime(&t, e0, e1));
    *ms = t;
  }
#undef HCK
  return done(RSG_OK, "");
}

// computehI, "Convert SCB field lines to RAM field lines" (src/ModRamScb.f90:252-300): xRAM, yRAM, zRAM, bRAM
// (nthe,nR,nT) and outsideSCB(nR,nT) from the SCB arrays x, y, z, bf, psi, alfa (nthe,npsi,nzeta+1), Lz(nR+1), MLT(nT).
// Lines outside the SCB domain come back zero with outsideSCB = 1 (the reference traces them with Geopack afterwards).
int rsg_hI_convert_lines(int device, int nthe, int npsi, int nzeta, int nR, int nT, int nThetaEquator, const double* x,
                         const double* y, const double* z, const double* bf, const double* psi, const double* alfa, const double* Lz,
                         const double* MLT, double* xRAM, double* yRAM, double* zRAM, double* bRAM, int* outsideSCB, double* ms) {
  if (!x || !y || !z || !bf || !psi || !alfa || !Lz || !MLT || !xRAM || !yRAM || !zRAM || !bRAM || !outsideSCB)
    return sfail(RSG_ERR_ARG, "null argument");
  if (nthe < 1 || npsi < 3 || nzeta < 3 || nR < 1 || nT < 1 || nThetaEquator < 1 || nThetaEquator > nthe || (long long)npsi * (nzeta - 1) < 9)
    return sfail(RSG_ERR_ARG, "bad dimensions");
  SCK(cudaSetDevice(device));
  const size_t n3 = (size_t)nthe * npsi * (nzeta + 1), nl = (size_t)nR * nT, no = (size_t)nthe * nl;
  const size_t nd = 6 * n3 + 3 * nl + nT + 4 * no;
  double* d = nullptr;
  int* dout = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto done = [&](int code, const std::string& msg) {
    if (d) cudaFree(d);
    if (dout) cudaFree(dout);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    return code == RSG_OK ? RSG_OK : sfail(code, msg);
  };
#define HCK(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return done(RSG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
  HCK(cudaStreamCreate(&st));
  HCK(cudaEventCreate(&e0));
  HCK(cudaEventCreate(&e1));
  HCK(cudaMalloc(&d, nd * sizeof(double)));
  HCK(cudaMalloc(&dout, nl * sizeof(int)));
  std::vector<double> qx, qy, al;
  hi_query_tables(nR, nT, Lz, MLT, qx, qy, al);
  HiConvArgs A;
  A.nthe = nthe; A.npsi = npsi; A.nzeta = nzeta; A.nR = nR; A.nT = nT; A.nThetaEquator = nThetaEquator;
  double* p = d;
  auto take = [&](size_t n) { double* q = p; p += n; return q; };
  cudaError_t uerr = cudaSuccess;      // first failed upload (checked after the batch)
  auto up = [&](const double* src, size_t n) {
    double* q = take(n);
    const cudaError_t e_ = cudaMemcpyAsync(q, src, n * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e_ != cudaSuccess && uerr == cudaSuccess) uerr = e_;
    return q;
  };
  A.x = up(x, n3); A.y = up(y, n3); A.z = up(z, n3); A.bf = up(bf, n3); A.psi = up(psi, n3); A.alfa = up(alfa, n3);
  A.qx = up(qx.data(), nl); A.qy = up(qy.data(), nl); A.alphaRAM = up(al.data(), nT);
  HCK(uerr);
  A.psiRAM = take(nl);
  A.xRAM = take(no); A.yRAM = take(no); A.zRAM = take(no); A.bRAM = take(no);
  A.outside = dout;
  HCK(cudaStreamSynchronize(st));                                   // the host tables go out of scope with the call only, but keep it simple
  HCK(cudaEventRecord(e0, st));
  { const int rc = hi_launch_convert(A, st); if (rc != RSG_OK) return done(rc, g_serr); }
  HCK(cudaEventRecord(e1, st));
  HCK(cudaMemcpyAsync(xRAM, A.xRAM, no * sizeof(double), cudaMemcpyDeviceToHost, st));
  HCK(cudaMemcpyAsync(yRAM, A.yRAM, no * sizeof(double), cudaMemcpyDeviceToHost, st));
  HCK(cudaMemcpyAsync(zRAM, A.zRAM, no * sizeof(double), cudaMemcpyDeviceToHost, st));
  HCK(cudaMemcpyAsync(bRAM, A.bRAM, no * sizeof(double), cudaMemcpyDeviceToHost, st));
  HCK(cudaMemcpyAsync(outsideSCB, dout, nl * sizeof(int), cudaMemcpyDeviceToHost, st));
  HCK(cudaStreamSynchronize(st));
  if (ms) {
    float t = 0.f;
    HCK(cudaEventElapsedTime(&t, e0, e1));
    *ms = t;
  }
#undef HCK
  return done(RSG_OK, "");
}

// =============================================================================
// rsg_hi: computehI (src/ModRamScb.f90:249-637) with everything resident -- the SCB arrays come from the host once
// per call or straight from an rsg_scb handle's device arrays, xRAM .. bRAM, I_cart .. bZEq_Cart and the RAM variables
// never leave the device between the three blocks, HDens_cart and the previous FNHS .. BNES persist between calls like
// the reference's module variables, and the new field arrays can go device-to-device into an rsg_ram handle.
// =============================================================================
struct rsg_hi {
  int device = 0, nthe = 0, npsi = 0, nzeta = 0, nR = 0, nT = 0, nPa = 0, nThetaEquator = 0;
  double bnormal = 1.0, Lz0 = 1.0;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  std::vector<void*> allocs;
  std::map<std::string, std::pair<double*, size_t>> arr;
  double *d_scb[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // x, y, z, bf, psi, alfa (own copies when the host supplies them)
  HiConvArgs C{};
  HiArgs I{};
  HiTailArgs T{};
  double* d_dens = nullptr;
  int *d_outSCB = nullptr, *d_outMGNP = nullptr, *d_scale = nullptr, *d_fail = nullptr;
  bool converted = false, fields_set = false, finished = false;
  double last_ms = 0.0;
  long long launches = 0;
  int dalloc(double** p, size_t n, const char* name) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(double));
    if (e != cudaSuccess) return sfail(RSG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cudaMemset(q, 0, n * sizeof(double));
    allocs.push_back(q);
    *p = (double*)q;
    if (name) arr[name] = {(double*)q, n};
    return RSG_OK;
  }
};

void rsg_hi_destroy(rsg_hi* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->st) { cudaStreamSynchronize(h->st); cudaStreamDestroy(h->st); }
  if (h->e0) cudaEventDestroy(h->e0);
  if (h->e1) cudaEventDestroy(h->e1);
  for (void* q : h->allocs) cudaFree(q);
  delete h;
}

int rsg_hi_create(rsg_hi** out, int device, int nthe, int npsi, int nzeta, int nR, int nT, int nPa, int nThetaEquator, double bnormal,
                  const double* chiVal, const double* mu, const double* Lz, const double* MLT, const double* PA, const double* PAbn) {
  if (!out || !chiVal || !mu || !Lz || !MLT || !PA || !PAbn) return sfail(RSG_ERR_ARG, "null argument");
  if (nthe < 3 || npsi < 3 || nzeta < 3 || nR < 3 || nT < 2 || nPa < 5 || nThetaEquator < 1 || nThetaEquator > nthe ||
      (long long)npsi * (nzeta - 1) < 9)
    return sfail(RSG_ERR_ARG, "bad dimensions");
  if (device >= 0) SCK(cudaSetDevice(device));
  rsg_hi* h = new rsg_hi();
  auto bail = [&](int rc) { rsg_hi_destroy(h); return rc; };
  if (cudaGetDevice(&h->device) != cudaSuccess) return bail(sfail(RSG_ERR_CUDA, "cudaGetDevice"));
  h->nthe = nthe; h->npsi = npsi; h->nzeta = nzeta; h->nR = nR; h->nT = nT; h->nPa = nPa; h->nThetaEquator = nThetaEquator;
  h->bnormal = bnormal; h->Lz0 = Lz[0];
  if (cudaStreamCreate(&h->st) != cudaSuccess || cudaEventCreate(&h->e0) != cudaSuccess || cudaEventCreate(&h->e1) != cudaSuccess)
    return bail(sfail(RSG_ERR_CUDA, "stream / event creation failed"));
  const size_t n3s = (size_t)nthe * npsi * (nzeta + 1), nl = (size_t)nR * nT, no = (size_t)nthe * nl, n3 = nl * nPa;
  const size_t nr2 = (size_t)(nR + 1) * nT, nr3 = nr2 * nPa;
#define HA(ptr, n, name) do { int rc_ = h->dalloc((double**)&(ptr), (n), (name)); if (rc_ != RSG_OK) return bail(rc_); } while (0)
  static const char* scbn[6] = {"x", "y", "z", "bf", "psi", "alfa"};
  for (int q = 0; q < 6; q++) HA(h->d_scb[q], n3s, scbn[q]);
  HiConvArgs& C = h->C;
  C.nthe = nthe; C.npsi = npsi; C.nzeta = nzeta; C.nR = nR; C.nT = nT; C.nThetaEquator = nThetaEquator;
  HA(C.qx, nl, nullptr); HA(C.qy, nl, nullptr); HA(C.alphaRAM, nT, nullptr);
  HA(C.psiRAM, nl, "psiRAM");
  HA(C.xRAM, no, "xRAM"); HA(C.yRAM, no, "yRAM"); HA(C.zRAM, no, "zRAM"); HA(C.bRAM, no, "bRAM");
  HiArgs& I = h->I;
  I.nthe = nthe; I.nR = nR; I.nT = nT; I.nPa = nPa; I.nThetaEquator = nThetaEquator; I.bnormal = bnormal;
  I.x = C.xRAM; I.y = C.yRAM; I.z = C.zRAM; I.b = C.bRAM;
  HA(h->d_dens, no, "density");
  I.dens = h->d_dens;
  HA(I.chi, nthe, nullptr); HA(I.mu, nPa, nullptr);
  HA(I.Icart, n3, nullptr); HA(I.Hcart, n3, nullptr); HA(I.Dcart, n3, nullptr); HA(I.bzeq, nl, nullptr);
  HiTailArgs& T = h->T;
  T.nR = nR; T.nT = nT; T.nPa = nPa; T.smooth = 0; T.DthI = 0.0;
  T.bnes1 = 0.32 / (Lz[0] * Lz[0] * Lz[0]) / 1.e4;                                  // :609
  hi_gauss_weights(T.w);
  T.I0 = I.Icart; T.H0 = I.Hcart; T.D0 = I.Dcart; T.bz0 = I.bzeq;
  HA(T.I1, n3, nullptr); HA(T.H1, n3, nullptr); HA(T.D1, n3, nullptr); HA(T.bz1, nl, nullptr); HA(T.hI, n3, nullptr); HA(T.iI, n3, nullptr);
  HA(T.I2, n3, nullptr); HA(T.H2, n3, nullptr); HA(T.D2, n3, nullptr); HA(T.hI2, n3, nullptr); HA(T.iI2, n3, nullptr);
  HA(T.FNHS, nr3, "FNHS"); HA(T.FNIS, nr3, "FNIS"); HA(T.BOUNHS, nr3, "BOUNHS"); HA(T.BOUNIS, nr3, "BOUNIS"); HA(T.HDNS, nr3, "HDNS");
  HA(T.dIdt, nr3, "dIdt"); HA(T.dHdt, nr3, "dHdt"); HA(T.dIbndt, nr3, "dIbndt");
  HA(T.BNES, nr2, "BNES"); HA(T.dBdt, nr2, "dBdt");
  HA(T.Lz, nR + 1, nullptr); HA(T.PA, nPa, nullptr); HA(T.PAbn, nPa, nullptr);
#undef HA
  {
    void* q = nullptr;
    if (cudaMalloc(&q, (2 * nl + nT + 1) * sizeof(int)) != cudaSuccess) return bail(sfail(RSG_ERR_CUDA, "cudaMalloc"));
    cudaMemset(q, 0, (2 * nl + nT + 1) * sizeof(int));
    h->allocs.push_back(q);
    h->d_outSCB = (int*)q; h->d_outMGNP = h->d_outSCB + nl; h->d_scale = h->d_outMGNP + nl; h->d_fail = h->d_scale + nT;
  }
  C.outside = h->d_outSCB;
  I.outside = h->d_outMGNP;
  T.outside = h->d_outMGNP; T.ScaleAt = h->d_scale; T.fail = h->d_fail;
  std::vector<double> qx, qy, al;
  hi_query_tables(nR, nT, Lz, MLT, qx, qy, al);
  auto upl = [&](const double* dst, const double* src, size_t n) {
    return cudaMemcpy((void*)dst, src, n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  if (!upl(C.qx, qx.data(), nl) || !upl(C.qy, qy.data(), nl) || !upl(C.alphaRAM, al.data(), nT) || !upl(I.chi, chiVal, nthe) ||
      !upl(I.mu, mu, nPa) || !upl(T.Lz, Lz, nR + 1) || !upl(T.PA, PA, nPa) || !upl(T.PAbn, PAbn, nPa))
    return bail(sfail(RSG_ERR_CUDA, "upload of the grid tables failed"));
  *out = h;
  return RSG_OK;
}

static int hi_named(rsg_hi* h, const char* name, double** p, size_t* n) {
  if (!h || !name) return sfail(RSG_ERR_ARG, "null argument");
  const size_t nl = (size_t)h->nR * h->nT, n3 = nl * h->nPa;
  const std::string k(name);
  // the *_cart arrays as the reference leaves them after the routine: smoothed when IntegralSmooth
  if (k == "I_cart") { *p = h->T.smooth ? h->T.I2 : h->T.I1; *n = n3; return RSG_OK; }
  if (k == "H_cart") { *p = h->T.smooth ? h->T.H2 : h->T.H1; *n = n3; return RSG_OK; }
  if (k == "HDens_cart") { *p = h->I.Dcart; *n = n3; return RSG_OK; }
  if (k == "bZEq_cart") { *p = h->T.bz1; *n = nl; return RSG_OK; }
  auto it = h->arr.find(k);
  if (it == h->arr.end()) return sfail(RSG_ERR_ARG, std::string("unknown array '") + name + "'");
  *p = it->second.first; *n = it->second.second;
  return RSG_OK;
}

/* previous values of the RAM variables (the reference's module arrays at entry): FNHS, FNIS, BOUNHS, BOUNIS, HDNS
 * (nR+1,nT,nPa), BNES (nR+1,nT); HDens_cart (nR,nT,nPa) may be NULL (zeros / what the last call left) */
int rsg_hi_set_ram_fields(rsg_hi* h, const double* FNHS, const double* FNIS, const double* BOUNHS, const double* BOUNIS, const double* HDNS,
                          const double* BNES, const double* HDens_cart) {
  if (!h || !FNHS || !FNIS || !BOUNHS || !BOUNIS || !HDNS || !BNES) return sfail(RSG_ERR_ARG, "null argument");
  SCK(cudaSetDevice(h->device));
  const size_t nr2 = (size_t)(h->nR + 1) * h->nT, nr3 = nr2 * h->nPa;
  auto up = [&](double* dst, const double* src, size_t n) { return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, h->st); };
  SCK(up(h->T.FNHS, FNHS, nr3)); SCK(up(h->T.FNIS, FNIS, nr3)); SCK(up(h->T.BOUNHS, BOUNHS, nr3)); SCK(up(h->T.BOUNIS, BOUNIS, nr3));
  SCK(up(h->T.HDNS, HDNS, nr3)); SCK(up(h->T.BNES, BNES, nr2));
  if (HDens_cart) SCK(up(h->I.Dcart, HDens_cart, (size_t)h->nR * h->nT * h->nPa));
  SCK(cudaStreamSynchronize(h->st));
  h->fields_set = true;
  return RSG_OK;
}

/* block 1 (src/ModRamScb.f90:252-300).  The SCB arrays x, y, z, bf, psi, alfa (nthe,npsi,nzeta+1) come from the host, or --
 * all six NULL and scb given -- from the device arrays of an rsg_scb handle on the same device (no copy at all).
 * outsideSCB (nR,nT; may be NULL) comes back; *nOutside (may be NULL) counts its ones. */
int rsg_hi_convert(rsg_hi* h, const double* x, const double* y, const double* z, const double* bf, const double* psi, const double* alfa,
                   rsg_scb* scb, int* outsideSCB, int* nOutside) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  SCK(cudaSetDevice(h->device));
  const size_t n3s = (size_t)h->nthe * h->npsi * (h->nzeta + 1), nl = (size_t)h->nR * h->nT;
  HiConvArgs& C = h->C;
  if (x || y || z || bf || psi || alfa) {
    if (!x || !y || !z || !bf || !psi || !alfa) return sfail(RSG_ERR_ARG, "give all six SCB arrays or none (with an rsg_scb handle)");
    const double* src[6] = {x, y, z, bf, psi, alfa};
    for (int q = 0; q < 6; q++) SCK(cudaMemcpyAsync(h->d_scb[q], src[q], n3s * sizeof(double), cudaMemcpyHostToDevice, h->st));
    C.x = h->d_scb[0]; C.y = h->d_scb[1]; C.z = h->d_scb[2]; C.bf = h->d_scb[3]; C.psi = h->d_scb[4]; C.alfa = h->d_scb[5];
  } else {
    if (!scb) return sfail(RSG_ERR_ARG, "no SCB arrays and no rsg_scb handle");
    if (scb->device != h->device || scb->nthe != h->nthe || scb->npsi != h->npsi || scb->nzeta != h->nzeta)
      return sfail(RSG_ERR_ARG, "rsg_scb handle on another device or with other dimensions");
    SCK(cudaStreamSynchronize(scb->st));                           // what the SCB solve left
    C.x = scb->dev.x; C.y = scb->dev.y; C.z = scb->dev.z; C.bf = scb->dev.bf; C.psi = scb->dev.psi; C.alfa = scb->dev.alfa;
  }
  SCK(cudaEventRecord(h->e0, h->st));
  SRET(hi_launch_convert(C, h->st));
  h->launches += 2;
  h->converted = true;
  h->finished = false;
  if (outsideSCB || nOutside) {
    std::vector<int> tmp;
    int* dst = outsideSCB;
    if (!dst) { tmp.resize(nl); dst = tmp.data(); }
    SCK(cudaMemcpyAsync(dst, h->d_outSCB, nl * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    SCK(cudaStreamSynchronize(h->st));
    if (nOutside) {
      int c = 0;
      for (size_t q = 0; q < nl; q++) c += dst[q] != 0;
      *nOutside = c;
    }
  }
  return RSG_OK;
}

/* one traced line (src/ModRamScb.f90:330-362, the host's Geopack tracer) for the RAM point (i, j), 1-based: nthe nodes each */
int rsg_hi_set_line(rsg_hi* h, int i, int j, const double* xl, const double* yl, const double* zl, const double* bl) {
  if (!h || !xl || !yl || !zl || !bl) return sfail(RSG_ERR_ARG, "null argument");
  if (i < 1 || i > h->nR || j < 1 || j > h->nT) return sfail(RSG_ERR_ARG, "point out of range");
  if (!h->converted) return sfail(RSG_ERR_STATE, "rsg_hi_set_line before rsg_hi_convert");
  SCK(cudaSetDevice(h->device));
  const size_t o = (size_t)h->nthe * ((i - 1) + (size_t)h->nR * (j - 1)), nb = (size_t)h->nthe * sizeof(double);
  SCK(cudaMemcpyAsync(h->C.xRAM + o, xl, nb, cudaMemcpyHostToDevice, h->st));
  SCK(cudaMemcpyAsync(h->C.yRAM + o, yl, nb, cudaMemcpyHostToDevice, h->st));
  SCK(cudaMemcpyAsync(h->C.zRAM + o, zl, nb, cudaMemcpyHostToDevice, h->st));
  SCK(cudaMemcpyAsync(h->C.bRAM + o, bl, nb, cudaMemcpyHostToDevice, h->st));
  SCK(cudaStreamSynchronize(h->st));
  return RSG_OK;
}

/* blocks 2 and 3 (:302-637 without the tracer): ScaleAt(nT) / outsideMGNP(nR,nT) from the host's magnetopause logic, or both
 * NULL: derived on the device as the reference's 'SWMF' branch does (:306-314: ScaleAt = first outside point of the MLT,
 * every outside line flagged).  density (nthe,nR,nT) from the host, or NULL: the RAIRDEN polynomial of the distance
 * (:365-371) on the device.  Then the integral block, the tail and the RAM variables, all on the handle's stream. */
int rsg_hi_finish(rsg_hi* h, const int* ScaleAt, const int* outsideMGNP, const double* density, int integral_smooth, double DthI,
                  int* gslerr) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  if (!h->converted) return sfail(RSG_ERR_STATE, "rsg_hi_finish before rsg_hi_convert");
  if (!h->fields_set) return sfail(RSG_ERR_STATE, "rsg_hi_finish before rsg_hi_set_ram_fields");
  if ((ScaleAt == nullptr) != (outsideMGNP == nullptr)) return sfail(RSG_ERR_ARG, "give ScaleAt and outsideMGNP together or neither");
  if (integral_smooth && (h->nR < 9 || h->nT < 9)) return sfail(RSG_ERR_ARG, "grid smaller than the 9 x 9 smoothing kernel");
  SCK(cudaSetDevice(h->device));
  const size_t nl = (size_t)h->nR * h->nT, no = (size_t)h->nthe * nl, n3 = nl * h->nPa;
  cudaStream_t st = h->st;
  if (ScaleAt) {
    for (int j = 0; j < h->nT; j++)
      if (ScaleAt[j] != 0 && (ScaleAt[j] < 3 || ScaleAt[j] > h->nR)) return sfail(RSG_ERR_ARG, "ScaleAt out of range");
    SCK(cudaMemcpyAsync(h->d_scale, ScaleAt, h->nT * sizeof(int), cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(h->d_outMGNP, outsideMGNP, nl * sizeof(int), cudaMemcpyHostToDevice, st));
  } else {
    k_hi_scaleat<<<nblk(h->nT, 64), 64, 0, st>>>(h->nR, h->nT, h->d_outSCB, h->d_outMGNP, h->d_scale, h->d_fail + 0);
    SCK(cudaGetLastError());
    h->launches++;
  }
  if (density) {
    SCK(cudaMemcpyAsync(h->d_dens, density, no * sizeof(double), cudaMemcpyHostToDevice, st));
  } else {
    k_hi_rairden<<<nblk((long long)no, 256), 256, 0, st>>>(no, h->C.xRAM, h->C.yRAM, h->C.zRAM, h->d_dens);
    SCK(cudaGetLastError());
    h->launches++;
  }
  SCK(cudaMemsetAsync(h->d_fail, 0, sizeof(int), st));
  SRET(hi_launch_lines(h->I, st));
  h->T.smooth = integral_smooth ? 1 : 0;
  h->T.DthI = DthI;
  SRET(hi_launch_tail(h->T, st));
  h->launches += 4 + (integral_smooth ? 1 : 0);
  // HDens_cart persists into the next call as the reference's module variable does (skipped lines keep their value, :399)
  SCK(cudaMemcpyAsync(h->I.Dcart, h->T.smooth ? h->T.D2 : h->T.D1, n3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
  SCK(cudaEventRecord(h->e1, st));
  int nfail = 0;
  SCK(cudaMemcpyAsync(&nfail, h->d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  SCK(cudaStreamSynchronize(st));
  if (!ScaleAt) {                                                  // the device-derived ScaleAt obeys the same range rule
    std::vector<int> sa(h->nT);
    SCK(cudaMemcpy(sa.data(), h->d_scale, h->nT * sizeof(int), cudaMemcpyDeviceToHost));
    for (int j = 0; j < h->nT; j++)
      if (sa[j] != 0 && sa[j] < 3) return sfail(RSG_ERR_ARG, "ScaleAt out of range (SCB domain ends inside the second RAM shell)");
  }
  if (gslerr) *gslerr = nfail;
  float t = 0.f;
  if (cudaEventElapsedTime(&t, h->e0, h->e1) == cudaSuccess) h->last_ms = t;
  h->finished = true;
  return RSG_OK;
}

/* the whole routine in one call: convert + finish with the device-side defaults ('SWMF' boundary branch, RAIRDEN density) */
int rsg_computehI(rsg_hi* h, const double* x, const double* y, const double* z, const double* bf, const double* psi, const double* alfa,
                  rsg_scb* scb, int integral_smooth, double DthI, int* gslerr) {
  SRET(rsg_hi_convert(h, x, y, z, bf, psi, alfa, scb, nullptr, nullptr));
  return rsg_hi_finish(h, nullptr, nullptr, nullptr, integral_smooth, DthI, gslerr);
}

/* any resident array to the host by name: xRAM yRAM zRAM bRAM density (nthe,nR,nT), psiRAM (nR,nT), I_cart H_cart HDens_cart
 * (nR,nT,nPa), bZEq_cart (nR,nT), FNHS FNIS BOUNHS BOUNIS HDNS dIdt dHdt dIbndt (nR+1,nT,nPa), BNES dBdt (nR+1,nT) */
int rsg_hi_get(rsg_hi* h, const char* name, double* host) {
  if (!host) return sfail(RSG_ERR_ARG, "null argument");
  double* p = nullptr;
  size_t n = 0;
  SRET(hi_named(h, name, &p, &n));
  SCK(cudaSetDevice(h->device));
  SCK(cudaMemcpyAsync(host, p, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  return RSG_OK;
}
/* integer results: which = 0 outsideSCB (nR,nT), 1 outsideMGNP (nR,nT), 2 ScaleAt (nT) */
int rsg_hi_get_int(rsg_hi* h, int which, int* host) {
  if (!h || !host) return sfail(RSG_ERR_ARG, "null argument");
  if (which < 0 || which > 2) return sfail(RSG_ERR_ARG, "which out of range");
  SCK(cudaSetDevice(h->device));
  const size_t nl = (size_t)h->nR * h->nT;
  const int* src = which == 0 ? h->d_outSCB : (which == 1 ? h->d_outMGNP : h->d_scale);
  SCK(cudaMemcpyAsync(host, src, (which == 2 ? (size_t)h->nT : nl) * sizeof(int), cudaMemcpyDeviceToHost, h->st));
  SCK(cudaStreamSynchronize(h->st));
  return RSG_OK;
}
double rsg_hi_last_ms(rsg_hi* h) { return h ? h->last_ms : 0.0; }
long long rsg_hi_launch_count(rsg_hi* h) { return h ? h->launches : 0; }

/* device pointers of the new field arrays for rsg_ram_set_fields_device (same process, same device):
 * order BNES, dBdt, FNHS, FNIS, BOUNHS, BOUNIS, HDNS, dIdt, dIbndt; *outsideMGNP the integer array (nR,nT) */
int rsg_hi_device_fields(rsg_hi* h, const double** ptrs9, const int** outsideMGNP) {
  if (!h || !ptrs9 || !outsideMGNP) return sfail(RSG_ERR_ARG, "null argument");
  if (!h->finished) return sfail(RSG_ERR_STATE, "rsg_hi_device_fields before rsg_hi_finish");
  const HiTailArgs& T = h->T;
  const double* p[9] = {T.BNES, T.dBdt, T.FNHS, T.FNIS, T.BOUNHS, T.BOUNIS, T.HDNS, T.dIdt, T.dIbndt};
  for (int q = 0; q < 9; q++) ptrs9[q] = p[q];
  *outsideMGNP = h->d_outMGNP;
  return RSG_OK;
}

double rsg_scb_last_ms(rsg_scb* h) { return h ? h->last_ms : 0.0; }
int rsg_scb_use_cluster(rsg_scb* h, int on) {
  if (!h) return sfail(RSG_ERR_ARG, "null handle");
  h->use_cluster = on != 0;
  return RSG_OK;
}
int rsg_scb_last_cluster(rsg_scb* h) { return h ? h->last_cluster : 0; }
long long rsg_scb_launch_count(rsg_scb* h) { return h ? h->launches : 0; }

}  // extern "C"
