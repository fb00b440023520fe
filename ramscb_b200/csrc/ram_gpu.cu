// Host side of the RAM C ABI (include/ramscb_gpu.h): device mirrors, per-species
// streams, small host tables (in the reference's operation order) and kernel
// launches.  No CPU compute path exists here: without a CUDA device every entry
// point fails with RSG_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unistd.h>
#include <string>
#include <vector>

#include "../../include/ramscb_gpu.h"
#include "ram_kernels.cuh"
#include "ram_fused.cuh"
#include "ram_coulomb.cuh"
#include "ram_diffcoef.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(RSG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                    std::to_string(__LINE__) + ")");                                  \
  } while (0)
#define CKL()                                                                                         \
  do {                                                                                                \
    cudaError_t e_ = cudaGetLastError();                                                              \
    if (e_ != cudaSuccess)                                                                            \
      return fail(RSG_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                    std::to_string(__LINE__) + ")");                                  \
  } while (0)
#define RET(x)                   \
  do {                           \
    int r_ = (x);                \
    if (r_ != RSG_OK) return r_; \
  } while (0)

constexpr double kQ = 1.602E-19, kCS = 2.998E8, kPI = 3.1415926535897932384626433832795;
constexpr int NSUM = 16;           // moment slots per species
constexpr int RES_N = 4 + NSUM + 1 + 4;  // [4] dt bits, [NSUM] sums, [1] WPADIF violations, [4] dt of the forward half step
constexpr int DTF_OFF = 4 + NSUM + 1;

struct Spec {
  cudaStream_t own = nullptr;
  cudaEvent_t ev = nullptr;
  double DTs = -1.0;        // DTs of the last DRIFTPARA
  double DTs_ce = -1.0;     // DTs of the last CEPARA
  double DTs_wl = -1.0;     // DTs of the last WAVELO table
  double setrc = 0.0;
  int cur = 0;              // which ping-pong buffer holds F2
  double* d_tab = nullptr;   // drift tables
  double* h_tab = nullptr;   // pinned staging for d_tab
  size_t n_tab = 0;
  double* d_ce = nullptr;    // sv[NE], ATLOS[NE][NR]
  double* h_ce = nullptr;
  double* d_wfac = nullptr;  // [NE][Pp]
  double* h_wfac = nullptr;
  double* d_FF = nullptr;    // FFACTOR [l][k][i]
  double* d_EPP = nullptr;   // [NE]
  double* d_FGEOS = nullptr; // [l][k][j]
  int* d_last = nullptr;     // DRIFTR inflow scan
  double* d_ghost = nullptr; // DRIFTR ghost cells per line
  double* d_part = nullptr;  // moment partials [nblk_sum][RSG_NMOM]
  double* d_tE = nullptr;    // ANISCH scratch [2][nch][NE][Pp]
  double* d_aE2 = nullptr;   // fused step: ANISCH rows of k_plane_rp<REV> [2][NPA * energy chunks][Pp] (allocated on first use)
  double* d_rFFA = nullptr;  // FAST ANISCH: 1/A(S,I,K) [k][i]
  double* d_flc = nullptr;   // FLC_coef of this species [l][k][Pp] (allocated by rsg_ram_set_flc_coef)
  double* d_wtab = nullptr;  // fused WPADIF: elimination factors (cA,cB) pairs [2*n] then RL [n], n = NPA*NE*Pp (k_wpadif_tables)
  double wtab_DTs = -1.0;    // DTs the factors were tabulated for (< 0: stale)
  double* d_coul = nullptr;  // COULPARA tables COULE, COULI, ATA, GTA, each [k][l], then cK = (COULE + COULI)(K) [NE] for the fused COULEN
  double* d_ctab = nullptr;  // fused COULMU: elimination factors (cA,cB) pairs [2*n] then RL [n] (k_coulmu_tables)
  double ctab_DTs = -1.0;    // DTs the factors were tabulated for (< 0: stale)
  double DTs_coul = -1.0;    // DTs of the last COULPARA
  unsigned long long* d_res = nullptr;  // slice of rsg_ram::d_res_all
  unsigned long long* h_res = nullptr;  // slice of rsg_ram::h_res_all (pinned)
  double* d_pp = nullptr;    // slice of d_pp_all: [2][Pp]
  SpecDev sd{};
};

}  // namespace

struct rsg_shard;   // multi-GPU state (ram_shard.inl)

struct rsg_ram {
  int nS, NR, NT, NE, NPA, NR1, P, Pp;
  rsg_shard* shard = nullptr;
  int device = 0, mode = RSG_MODE_EXACT;
  bool grids_set = false, fields_set = false, efield_set = false;
  double Kp = 0, Kpmax12 = 0;
  std::vector<double> RLZ, LZ, EKEV, WE, DE, EBND, MU, WMU, DMU, UPA, GREL, GRBND, V, VBND, EPP, ERNH, RMAS;
  std::vector<double> WALOS1, WALOS2, WALOS3;
  std::vector<int> QS, kind, khi;
  RamDev dev{};
  std::vector<void*> allocs;
  double* d_F2[2] = {nullptr, nullptr};   // ping-pong buffers [nS][NPA][NE][Pp]
  size_t specStride = 0;
  double* d_stage = nullptr;  // host-layout image of F2 (nS*P*NE*NPA)
  double* d_diff[4] = {nullptr, nullptr, nullptr, nullptr};
  double* d_zero4 = nullptr;  // all-zero diffusion coefficient
  double* d_flctab = nullptr; // PARA_FLC inputs: r_curvEq, zeta1Eq, zeta2Eq (P each), V(S,:), LZ
  double* d_NECR = nullptr;
  DiffTabs dtab{};             // wave tables of the ANISCH diffusion-coefficient rebuild (rsg_ram_set_wave_tables)
  bool dtab_set = false;
  double* d_XNE = nullptr;
  int* d_dcerr = nullptr;
  double* d_dcgrel = nullptr;
  double* d_dtinit = nullptr;
  int* d_outlist = nullptr;   // plane indices p of flagged (outsideMGNP) cells with 2 <= J <= NT-1
  int nout = 0;
  int* d_tilemax = nullptr;
  int ntiles = 0;
  unsigned long long* d_res_all = nullptr;
  unsigned long long* h_res_all = nullptr;
  double* d_pp_all = nullptr;
  double* h_pp_all = nullptr;
  cudaStream_t ext = nullptr;   // user stream (rsg_ram_set_stream)
  cudaStream_t prepst = nullptr;  // prep kernels and the batched rsg_ram_run
  cudaEvent_t prepev = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  double prep_DTs = -1.0;
  bool step_dirty = true;       // e-field or fields changed since last k_prep_step
  std::mutex mu;
  Spec sp[RSG_MAX_SPECIES];
  long long launches = 0;
  // optional per-stage device timing of rsg_ram_run (CUDA events on the run stream)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<std::string> prof_stage;   // stage that starts at event i
  size_t prof_n = 0;
  std::vector<std::pair<std::string, std::pair<double, long long>>> prof_acc;  // name -> (ms, count)
  int nblk_sum = 0, sum_threads = 256;
  int segE = 12, segMU = 12, segP = 12, kcR = 7;
  // DRIFTR inflow scan: valid while DTs, the E field, the fields and the mode are unchanged
  bool inflow_ok[RSG_MAX_SPECIES] = {false};
  bool cfl_ok[RSG_MAX_SPECIES] = {false};   // cached CFL limits of the fused path (same dependencies)
  unsigned long long* d_cfl_all = nullptr;  // [nS][4]
  unsigned long long* hd_res_all = nullptr; // device-side addresses of the pinned h_res_all / h_pp_all
  double* hd_pp_all = nullptr;
  // CUDA graph of the single-GPU step (valid for one (DTs, flags, mode); rebuilt when they change)
  cudaGraphExec_t gexec = nullptr;
  double g_DTs = -1.0;
  int g_flags = -1, g_mode = -1;
  long long g_launches = 0;
  int g_s0 = 0, g_ns = 0;
  bool g_tpos = false;   // COULMU's T > 0 switch is baked into the captured launch
  cudaStream_t g_stream = nullptr;
  double T_elapsed = 0.0;
  int fp_nb = 0;   // column blocks of the last rsg_ram_fpart_columns (k_finalize sums that many partials)
  bool use_graph = true;
  // fused FAST path (ram_fused.cuh): shared-memory plane / column kernels
  bool use_fused = true;
  bool use_fused_wpi = true;   // WPADIF inside the column kernel (RSG_NO_FUSE_WPI=1 / rsg_ram_use_fused(h, 1): one kernel per operator with WPI / EMIC)
  int kcPlane = 0, colT = 0, planeT = 0;
  bool planeOdd = false;
  bool planeTma = true;     // k_plane_rp stages plane rows with TMA bulk copies (RSG_PLANE_TMA=0: cp.async chunks)
  int anischLch = 12;   // pitch angles per thread in the ANISCH pitch-angle sums
  bool in_step = false, fwd_half = false;   // set by rsg_ram_part_*: CFL slots are reset once per step
  unsigned long long* d_res_init = nullptr;
  unsigned long long* d_wviol = nullptr;    // [nS] rows of the WPADIF matrices that are not diagonally dominant (per tabulation)

  // rsg_ram_run_host: copy stream and per-chunk events of the pipelined host <-> device step
  cudaStream_t copyst = nullptr;
  cudaEvent_t evh[2 * 16] = {nullptr};

  cudaStream_t st(int s) { return ext ? ext : sp[s].own; }
  cudaStream_t pst() { return ext ? ext : prepst; }
  template <class T>
  int dalloc(T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) return fail(RSG_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    e = cudaMemset(q, 0, n * sizeof(T));
    if (e != cudaSuccess) return fail(RSG_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
    allocs.push_back(q);
    *p = (T*)q;
    return RSG_OK;
  }
};

namespace {

void shard_release(rsg_ram* h);   // ram_shard.inl

inline int nblk(long long n, int b) { return (int)((n + b - 1) / b); }

int check_S(rsg_ram* h, int S) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (S < 1 || S > h->nS) return fail(RSG_ERR_ARG, "species index out of range");
  return RSG_OK;
}

template <class T>
int up(T* dst, const T* src, size_t n) {
  CK(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return RSG_OK;
}

// (NR,NT,NE,NPA) Fortran array -> device [l][k][Pp]
void to_planes4(const rsg_ram* h, const double* src, std::vector<double>& out) {
  out.assign((size_t)h->NPA * h->NE * h->Pp, 0.0);
  for (int l = 0; l < h->NPA; ++l)
    for (int k = 0; k < h->NE; ++k)
      std::memcpy(&out[((size_t)l * h->NE + k) * h->Pp], &src[((size_t)l * h->NE + k) * h->P], sizeof(double) * h->P);
}

// species pack with the current ping-pong assignment
void make_pack(rsg_ram* h, SpecPack& pk, int s0 = 0, int ns = -1) {
  if (ns < 0) ns = h->nS - s0;
  for (int s = s0; s < s0 + ns; ++s) {
    Spec& sp = h->sp[s];
    sp.sd.F = h->d_F2[sp.cur] + h->specStride * s;
    sp.sd.Fo = h->d_F2[sp.cur ^ 1] + h->specStride * s;
    // inside the fused step the forward sweeps park their CFL minima in a scratch slot: the
    // reference resets DtDrift* at every call, so only the reverse half step's values survive
    sp.sd.dtw = sp.d_res + ((h->in_step && h->fwd_half) ? DTF_OFF : 0);
    pk.s[s] = sp.sd;
  }
}
void flip(rsg_ram* h, int s0, int ns) {
  for (int s = s0; s < s0 + ns; ++s) h->sp[s].cur ^= 1;
}

// species-independent coefficient planes that depend on DTs / the E field
int ensure_step(rsg_ram* h, double DTs, cudaStream_t only = nullptr) {
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->step_dirty && h->prep_DTs == DTs) return RSG_OK;
  if (!h->fields_set || !h->efield_set) return fail(RSG_ERR_STATE, "DRIFTPARA before set_fields/set_efield");
  cudaStream_t ps = only ? only : h->pst();
  const bool join = !h->ext && !only;
  if (join)
    for (int s = 0; s < h->nS; ++s) {
      CK(cudaEventRecord(h->sp[s].ev, h->sp[s].own));
      CK(cudaStreamWaitEvent(ps, h->sp[s].ev, 0));
    }
  RamDev dv = h->dev;
  dv.DTs = DTs;
  k_prep_step<<<nblk((long long)h->NPA * h->Pp, 256), 256, 0, ps>>>(dv);
  CKL();
  k_transpose_rcoef<<<dim3(nblk(h->P, 256), h->NPA), 256, 0, ps>>>(dv);
  CKL();
  h->launches += 2;
  if (join) {
    CK(cudaEventRecord(h->prepev, ps));
    for (int s = 0; s < h->nS; ++s) CK(cudaStreamWaitEvent(h->sp[s].own, h->prepev, 0));
  }
  h->prep_DTs = DTs;
  h->step_dirty = false;
  for (int s = 0; s < h->nS; ++s) h->inflow_ok[s] = h->cfl_ok[s] = false;   // CR changed
  return RSG_OK;
}

// mark the start of a stage of rsg_ram_run on stream st
int prof_mark(rsg_ram* h, const char* stage, cudaStream_t st) {
  if (!h->prof_on) return RSG_OK;
  if (h->prof_n == h->prof_ev.size()) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    h->prof_ev.push_back(e);
    h->prof_stage.push_back("");
  }
  h->prof_stage[h->prof_n] = stage;
  CK(cudaEventRecord(h->prof_ev[h->prof_n], st));
  h->prof_n++;
  return RSG_OK;
}
// after the stream has been synchronised: fold the intervals into the accumulators
int prof_fold(rsg_ram* h) {
  if (!h->prof_on) return RSG_OK;
  for (size_t q = 0; q + 1 < h->prof_n; ++q) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->prof_ev[q], h->prof_ev[q + 1]));
    const std::string& nm = h->prof_stage[q];
    bool found = false;
    for (auto& a : h->prof_acc)
      if (a.first == nm) { a.second.first += ms; a.second.second++; found = true; break; }
    if (!found) h->prof_acc.push_back({nm, {(double)ms, 1}});
  }
  h->prof_n = 0;
  return RSG_OK;
}

RamDev devfor(rsg_ram* h, double DTs) {
  RamDev dv = h->dev;
  dv.DTs = DTs;
  return dv;
}

int seg_count(int ncell, int seg) { return (ncell + seg - 1) / seg; }

// ---- batched launches: species s0 .. s0+ns-1 on stream st ----------------------
int L_inflow(rsg_ram* h, int s0, int ns, cudaStream_t st) {
  bool all_ok = true;
  for (int s = s0; s < s0 + ns; ++s) all_ok = all_ok && h->inflow_ok[s];
  if (all_ok) return RSG_OK;   // coefficients unchanged since the last scan
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  RamDev dv = devfor(h, h->sp[s0].DTs);
  dim3 g(h->ntiles, ns);
  if (h->mode == RSG_MODE_FAST) k_driftr_inflow<true><<<g, SCAN_TILE, 0, st>>>(dv, pk, s0, h->d_tilemax, h->ntiles);
  else k_driftr_inflow<false><<<g, SCAN_TILE, 0, st>>>(dv, pk, s0, h->d_tilemax, h->ntiles);
  CKL();
  k_driftr_scan<<<g, SCAN_TILE, 0, st>>>(dv, pk, s0, h->d_tilemax, h->ntiles);
  CKL();
  h->launches += 2;
  for (int s = s0; s < s0 + ns; ++s) h->inflow_ok[s] = true;
  return RSG_OK;
}
int reset_dt(rsg_ram* h, int s0, int ns, int which, cudaStream_t st) {
  if (h->in_step) return RSG_OK;   // done once for the whole step in rsg_ram_part_fwd
  for (int s = s0; s < s0 + ns; ++s)
    CK(cudaMemcpyAsync(h->sp[s].d_res + which, h->d_dtinit + which, 8, cudaMemcpyDeviceToDevice, st));
  return RSG_OK;
}
// mom_slot >= 0: also produce the SUMRC moment of the updated F2 into that result slot
int L_driftr(rsg_ram* h, int s0, int ns, cudaStream_t st, int l0 = 0, int nl = -1, int mom_slot = -1) {
  if (nl < 0) nl = h->NPA - l0;
  RET(reset_dt(h, s0, ns, 0, st));
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  const int KC = h->kcR, KG = (h->NE + KC - 1) / KC;
  dim3 g(nblk(h->P, 248), nl * KG, ns);   // 8 warps x 31 cells per CTA, KC energies per thread
  const RamDev dv = devfor(h, h->sp[s0].DTs);
  const bool fast = h->mode == RSG_MODE_FAST;
  if (mom_slot >= 0) {
    if (fast) k_driftr<true, true><<<g, 256, 0, st>>>(dv, pk, s0, KC, KG, l0);
    else k_driftr<false, true><<<g, 256, 0, st>>>(dv, pk, s0, KC, KG, l0);
  } else {
    if (fast) k_driftr<true, false><<<g, 256, 0, st>>>(dv, pk, s0, KC, KG, l0);
    else k_driftr<false, false><<<g, 256, 0, st>>>(dv, pk, s0, KC, KG, l0);
  }
  CKL();
  h->launches++;
  flip(h, s0, ns);
  if (mom_slot >= 0) {
    k_sum_final<<<dim3(1, ns), 256, 0, st>>>(pk, s0, (int)(g.x * g.y * 8), 1, mom_slot);
    CKL();
    h->launches++;
  }
  return RSG_OK;
}
int L_driftp(rsg_ram* h, int s0, int ns, cudaStream_t st, int l0 = 0, int nl = -1) {
  if (nl < 0) nl = h->NPA - l0;
  RET(reset_dt(h, s0, ns, 1, st));
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  const int nseg = seg_count(h->NT - 1, h->segP);
  dim3 g(nblk((long long)h->NE * h->NR, 128), nl * nseg, ns);
  if (h->mode == RSG_MODE_FAST) k_driftp<true><<<g, 128, 0, st>>>(devfor(h, h->sp[s0].DTs), pk, s0, h->segP, nseg, l0);
  else k_driftp<false><<<g, 128, 0, st>>>(devfor(h, h->sp[s0].DTs), pk, s0, h->segP, nseg, l0);
  CKL();
  h->launches++;
  flip(h, s0, ns);
  return RSG_OK;
}
int L_drifte(rsg_ram* h, int s0, int ns, cudaStream_t st, int l0 = 0, int nl = -1) {
  if (nl < 0) nl = h->NPA - l0;
  RET(reset_dt(h, s0, ns, 2, st));
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  const int nseg = seg_count(h->NE, h->segE);
  dim3 g(nblk(h->P, 128), nl * nseg, ns);
  if (h->mode == RSG_MODE_FAST) k_drifte<true><<<g, 128, 0, st>>>(devfor(h, h->sp[s0].DTs), pk, s0, h->segE, nseg, l0);
  else k_drifte<false><<<g, 128, 0, st>>>(devfor(h, h->sp[s0].DTs), pk, s0, h->segE, nseg, l0);
  CKL();
  h->launches++;
  flip(h, s0, ns);
  return RSG_OK;
}
int L_driftmu(rsg_ram* h, int s0, int ns, cudaStream_t st, int k0 = 0, int nk = -1, int mom_slot = -1) {
  if (nk < 0) nk = h->NE - k0;
  RET(reset_dt(h, s0, ns, 3, st));
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  const int nseg = seg_count(h->NPA - 2, h->segMU);
  dim3 g(nblk(h->P, 128), nk * nseg, ns);
  const RamDev dv = devfor(h, h->sp[s0].DTs);
  const bool fast = h->mode == RSG_MODE_FAST;
  if (mom_slot >= 0) {
    if (fast) k_driftmu<true, true><<<g, 128, 0, st>>>(dv, pk, s0, h->segMU, nseg, k0);
    else k_driftmu<false, true><<<g, 128, 0, st>>>(dv, pk, s0, h->segMU, nseg, k0);
  } else {
    if (fast) k_driftmu<true, false><<<g, 128, 0, st>>>(dv, pk, s0, h->segMU, nseg, k0);
    else k_driftmu<false, false><<<g, 128, 0, st>>>(dv, pk, s0, h->segMU, nseg, k0);
  }
  CKL();
  h->launches++;
  flip(h, s0, ns);
  if (mom_slot >= 0) {
    k_sum_final<<<dim3(1, ns), 256, 0, st>>>(pk, s0, (int)(g.x * g.y * 4), 1, mom_slot);
    CKL();
    h->launches++;
  }
  return RSG_OK;
}
PlaneRange full_range(rsg_ram* h) { return PlaneRange{0, h->NPA, 0, h->NE}; }
int L_sumrc(rsg_ram* h, int s0, int ns, int slot, cudaStream_t st, PlaneRange pr = PlaneRange{0, -1, 0, -1}) {
  if (pr.nl < 0) pr = full_range(h);
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  const int nb = pr.nl * pr.nk;
  k_sumrc_partial<<<dim3(nb, ns), h->sum_threads, 0, st>>>(h->dev, pk, s0, pr);
  CKL();
  k_sum_final<<<dim3(1, ns), 256, 0, st>>>(pk, s0, nb, 1, slot);
  CKL();
  h->launches += 2;
  return RSG_OK;
}
int L_loss(rsg_ram* h, int s, int op, double DTs, cudaStream_t st) {
  SpecPack pk;
  make_pack(h, pk, s, 1);
  k_loss<<<dim3(nblk(h->P, 256), h->NPA * h->NE, 1), 256, 0, st>>>(devfor(h, DTs), pk, s, op);
  CKL();
  h->launches++;
  return RSG_OK;
}
int L_loss_mid(rsg_ram* h, int s0, int ns, int doA, double DTs, int slot, cudaStream_t st, PlaneRange pr = PlaneRange{0, -1, 0, -1}) {
  if (pr.nl < 0) pr = full_range(h);
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  const int nb = pr.nl * pr.nk;
  if (h->mode == RSG_MODE_FAST) k_loss_mid<true><<<dim3(nb, ns), h->sum_threads, 0, st>>>(devfor(h, DTs), pk, s0, doA, pr);
  else k_loss_mid<false><<<dim3(nb, ns), h->sum_threads, 0, st>>>(devfor(h, DTs), pk, s0, doA, pr);
  CKL();
  k_sum_final<<<dim3(4, ns), 256, 0, st>>>(pk, s0, nb, 4, slot);
  CKL();
  h->launches += 2;
  return RSG_OK;
}
// flc: FLCscatter (src/ModRamLoss.f90:513-575) = the same kernel with the species' FLC_coef as
// the only coefficient array
int L_wpadif(rsg_ram* h, int s, double DTs, cudaStream_t st, int k0 = 0, int nk = -1, bool flc = false) {
  if (nk < 0) nk = h->NE - k0;
  Spec& sp = h->sp[s];
  const double *DA, *DB;
  if (flc) { DA = sp.d_flc; DB = nullptr; }
  else if (h->kind[s] == RSG_KIND_E) { DA = h->d_diff[0]; DB = h->d_diff[1]; }
  else { DA = h->d_diff[2]; DB = h->d_diff[3]; }
  if (!DA && !DB) return fail(RSG_ERR_STATE, flc ? "FLCscatter before set_flc_coef" : "WPADIF before set_diffcoef");
  sp.sd.DA = DA ? DA : h->d_zero4;
  sp.sd.DB = DB ? DB : h->d_zero4;
  SpecPack pk;
  make_pack(h, pk, s, 1);
  const int T = 64;
  const size_t smem = sizeof(double) * 2 * h->NPA * T;
  static thread_local size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    CK(cudaFuncSetAttribute(k_wpadif, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  CK(cudaMemsetAsync(sp.d_res + 4 + NSUM, 0, sizeof(unsigned long long), st));
  k_wpadif<<<dim3(nblk((long long)nk * h->Pp, T), 1), T, smem, st>>>(devfor(h, DTs), pk, s, k0, nk);
  CKL();
  h->launches++;
  return RSG_OK;
}
int L_anisch(rsg_ram* h, int s0, int ns, cudaStream_t st, int l0 = 0, int nl = -1) {
  if (nl < 0) nl = h->NPA - l0;
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  if (h->mode == RSG_MODE_FAST) {
    const int LCH = 6;
    const int nch = std::max(1, std::min(16, (nl + LCH - 1) / LCH));
    const int lch = (nl + nch - 1) / nch;
    k_anisch_pa_fast<<<dim3(nblk(h->Pp, 32), h->NE, ns), dim3(32, nch), 0, st>>>(h->dev, pk, s0, l0, nl, lch);
  } else {
    k_anisch_pa<<<dim3(nblk(h->Pp, 128), h->NE, ns), 128, 0, st>>>(h->dev, pk, s0, l0, nl);
  }
  CKL();
  const double cv = kCS * 100;
  const double RFAC = 4 * kPI / cv;
  k_anisch_en<<<dim3(nblk(h->P, 128), ns), 128, 0, st>>>(h->dev, pk, s0, RFAC, h->khi[0], h->khi[1], h->khi[2], h->khi[3], h->khi[4]);
  CKL();
  h->launches += 2;
  return RSG_OK;
}

// ---- fused FAST path --------------------------------------------------------------
constexpr int COL_PG = 4;
struct ColPlan { ColCfg cfg; int T; size_t smem; };
ColPlan col_plan(const rsg_ram* h, bool wpi = false) {
  ColPlan c{};
  const int NE = h->NE, NPA = h->NPA;
  c.cfg.NEs = NE | 1;
  const int linesE = NPA * COL_PG, linesM = NE * COL_PG;
  int T = h->colT ? h->colT : ((linesE + 31) / 32 * 32) * (NE > 48 ? 2 : 1);
  c.T = std::min(1024, (T + 31) / 32 * 32);
  auto segs = [](int n, int want, int* nseg, int* seg) {
    want = std::max(1, want);
    *seg = (n + want - 1) / want;
    *nseg = (n + *seg - 1) / *seg;
  };
  segs(NE, c.T / linesE, &c.cfg.nsegE, &c.cfg.segE);
  segs(NPA - 2, c.T / linesM, &c.cfg.nsegM, &c.cfg.segM);
  segs(NPA, c.T / linesM, &c.cfg.nsegL, &c.cfg.segL);
  c.smem = sizeof(double) * ((size_t)NPA * c.cfg.NEs * COL_PG + 6 * (size_t)NPA * COL_PG + 8 * (size_t)NE + 2 * (size_t)NE * COL_PG +
                             4 * (size_t)NPA + 64 + 5 * 32 + (wpi ? 2 * (size_t)NPA * COL_PG + 4 * (size_t)NE : 0));
  return c;
}
int fused_part_off(const rsg_ram* h) { return (((h->P + COL_PG - 1) / COL_PG) * 5 + 15) & ~15; }
// WPADIF moments of the column kernel: behind the reverse plane kernel's partials (at most NE*NPA of them)
int fused_wpart_off(const rsg_ram* h) { return (fused_part_off(h) + h->NE * h->NPA + 15) & ~15; }
// Coulomb moments of the column kernel: behind the two WPADIF moments of every block
int fused_cpart_off(const rsg_ram* h) { return (fused_wpart_off(h) + ((h->P + COL_PG - 1) / COL_PG) * 2 + 15) & ~15; }
struct PlanePlan { PlaneCfg cfg; int T; size_t smem; };
PlanePlan plane_plan(const rsg_ram* h) {
  PlanePlan c{};
  const int NR = h->NR, NT = h->NT;
  // row stride of the shared copy: odd, or 2*odd when NR is even (16-byte rows; the radial walks of
  // 16 consecutive MLT lines then hit 8 different 8-byte banks instead of 1-4)
  c.cfg.NRp = (NR & 1) ? NR : (((NR / 2) & 1) ? NR : NR + 2);
  if (h->planeOdd) c.cfg.NRp = NR | 1;   // odd stride: conflict-free walks, 8-byte staging
  c.cfg.PS = (NT * c.cfg.NRp + 1) & ~1;
  const size_t plane_bytes = sizeof(double) * (size_t)c.cfg.PS;
  // planes per CTA: enough lines for a few warps, at most ~100 KB so two CTAs share an SM
  int KC = h->kcPlane > 0 ? h->kcPlane : std::max(1, std::min(12, (int)(100 * 1024 / plane_bytes)));
  KC = std::max(1, std::min({KC, h->NE, (int)(200 * 1024 / plane_bytes)}));
  c.cfg.KC = KC;
  const int linesR = NT * KC, linesP = NR * KC;
  // one thread per line; lines longer than ~32 cells are split so that a CTA has >= 8 warps
  int T = h->planeT ? h->planeT : std::max(linesR, linesP);
  if (!h->planeT) {
    const int longest = std::max(NR, NT);
    if (longest > 40) T = std::min(512, T * ((longest + 31) / 32));
  }
  c.T = std::max(64, std::min(512, (T + 31) / 32 * 32));
  auto segs = [](int n, int want, int* nseg, int* seg) {
    want = std::max(1, want);
    *seg = std::max(2, (n + want - 1) / want);
    *nseg = (n + *seg - 1) / *seg;
  };
  segs(NR - 1, c.T / linesR, &c.cfg.nsegR, &c.cfg.segR);
  segs(NT - 1, c.T / linesP, &c.cfg.nsegP, &c.cfg.segP);
  // DRIFTP: the cells J = NT-1 and J = NT must belong to ONE thread -- the step at interface NT-1 re-reads the stored F(1)
  // (colp[0]), which the thread that owns J = NT overwrites at its end (k_plane_rp, wrapfix).  A last segment of one cell
  // is avoided by one more cell per segment.
  while (c.cfg.nsegP > 1 && (NT - 1) - (c.cfg.nsegP - 1) * c.cfg.segP < 2) {
    c.cfg.segP++;
    c.cfg.nsegP = (NT - 1 + c.cfg.segP - 1) / c.cfg.segP;
  }
  c.smem = sizeof(double) * ((size_t)KC * c.cfg.PS + 2 * (size_t)KC * NT + 32) + sizeof(int) * (((size_t)KC * NT + 1) & ~(size_t)1);
  return c;
}
size_t plane_smem(const rsg_ram* h) { return plane_plan(h).smem; }
// species whose step contains WPADIF (src/ModRamRun.f90:91-104): electrons with #USEWPI, H+ with #USEEMIC
int wpadif_mask(const rsg_ram* h, int flags) {
  int m = 0;
  for (int s = 0; s < h->nS; ++s)
    if (((flags & RSG_F_WPI) && h->kind[s] == RSG_KIND_E) || ((flags & RSG_F_EMIC) && h->kind[s] == RSG_KIND_H)) m |= 1 << s;
  return m;
}
// the fused kernels cover the default operator set, with or without the WPI / EMIC pitch-angle
// diffusion, on a whole grid in FAST mode (the Coulomb operators run one kernel per operator)
bool fused_ok(const rsg_ram* h, int flags) {
  if (!h->use_fused || h->mode != RSG_MODE_FAST || (flags & ~(RSG_F_WPI | RSG_F_EMIC | RSG_F_COULOMB)) != 0) return false;
  if (h->NR < 4 || h->NT < 5 || h->NE < 3 || h->NPA < 13) return false;
  const bool wpi = wpadif_mask(h, flags) != 0 || (flags & RSG_F_COULOMB);
  if (wpi && !h->use_fused_wpi) return false;
  return col_plan(h, wpi).smem <= 220 * 1024 && plane_smem(h) <= 220 * 1024;
}
template <typename K>
int opt_in_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return RSG_OK;
}
void split_range(int n, int parts, int idx, int* start, int* count);
// host layout <-> device layout for ALL species of a range of (l, k) planes in one pass: a thread moves the nS contiguous
// doubles of one (plane, position) between the staging image and the species buffers (k_f2_from_host / to_host make one
// strided pass per species).  grid: x = tiles of p, y = planes of the range
struct SpecPtrs { double* F[RSG_MAX_SPECIES]; };
template <bool TO_HOST>
__global__ void __launch_bounds__(256) k_f2_host_all(RamDev d, double* __restrict__ stage, SpecPtrs sp, int plane0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const size_t plane = (size_t)plane0 + blockIdx.y;
  if (p >= d.Pp) return;
  if (p < d.P) {
    double* q = stage + (plane * d.P + p) * d.nS;
    for (int s = 0; s < d.nS; ++s) {
      if (TO_HOST) q[s] = sp.F[s][plane * d.Pp + p];
      else sp.F[s][plane * d.Pp + p] = q[s];
    }
  } else if (!TO_HOST) {
    for (int s = 0; s < d.nS; ++s) sp.F[s][plane * d.Pp + p] = 0.0;
  }
}

int L_plane_rp(rsg_ram* h, int s0, int ns, cudaStream_t st, bool rev, int l0 = 0, int nl = -1, const PeerView* peer = nullptr,
               int part_l0 = -1) {
  if (nl < 0) nl = h->NPA - l0;
  static const PeerView kNoPeer{};
  const PeerView& pv = peer ? *peer : kNoPeer;
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  const RamDev dv = devfor(h, h->sp[s0].DTs);
  PlanePlan c = plane_plan(h);
  const int KG = (h->NE + c.cfg.KC - 1) / c.cfg.KC;
  const dim3 g(KG, nl, ns);
  c.cfg.part_off = fused_part_off(h);      // after the column kernel's partials
  // a launch over [l0, l0+nl) writes its reduction rows from row 0; chunked launches of one range (rsg_ram_run_host) place
  // their rows behind those of the pitch angles part_l0 .. l0-1
  if (part_l0 >= 0) c.cfg.part_off += (l0 - part_l0) * KG;
  c.cfg.l0 = l0;
  c.cfg.anisch = (rev && h->sp[s0].d_aE2) ? 1 : 0;
  c.cfg.tma = (h->planeTma && h->NR * 8 >= 512) ? 1 : 0;   // rows shorter than 512 B: the per-copy overhead shows (measured)
  if (rev) { RET(opt_in_smem(k_plane_rp<true>, c.smem)); k_plane_rp<true><<<g, c.T, c.smem, st>>>(dv, pk, s0, c.cfg, pv); }
  else if (peer) { RET(opt_in_smem(k_plane_rp<false, true>, c.smem)); k_plane_rp<false, true><<<g, c.T, c.smem, st>>>(dv, pk, s0, c.cfg, pv); }
  else { RET(opt_in_smem(k_plane_rp<false>, c.smem)); k_plane_rp<false><<<g, c.T, c.smem, st>>>(dv, pk, s0, c.cfg, pv); }
  CKL();
  h->launches++;
  return RSG_OK;
}
// CFL limits of the fused path: evaluated when the coefficient set changed, else cached
int L_cfl(rsg_ram* h, int s0, int ns, cudaStream_t st) {
  bool all_ok = true;
  for (int s = s0; s < s0 + ns; ++s) all_ok = all_ok && h->cfl_ok[s];
  if (all_ok) return RSG_OK;
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  CK(cudaMemcpy2DAsync(h->d_cfl_all + (size_t)s0 * 4, 4 * sizeof(unsigned long long), h->d_res_init + (size_t)s0 * RES_N,
                       RES_N * sizeof(unsigned long long), 4 * sizeof(unsigned long long), ns, cudaMemcpyDeviceToDevice, st));
  const int KCH = 5;
  k_cfl_fast<<<dim3(h->NPA, (h->NE + KCH - 1) / KCH, ns), 256, 0, st>>>(devfor(h, h->sp[s0].DTs), pk, s0, h->d_cfl_all, KCH);
  CKL();
  h->launches++;
  for (int s = s0; s < s0 + ns; ++s) h->cfl_ok[s] = true;
  return RSG_OK;
}
// elimination factors of the fused WPADIF for the species of `mask`: tabulated when DTs, the
// diffusion coefficients or the fields changed, else cached.  Allocates: call outside stream capture.
int L_wtab(rsg_ram* h, int mask, double DTs, cudaStream_t st) {
  const size_t n = h->specStride;
  for (int s = 0; s < h->nS; ++s) {
    if (!((mask >> s) & 1)) continue;
    Spec& sp = h->sp[s];
    const bool el = h->kind[s] == RSG_KIND_E;
    const double* DA = el ? h->d_diff[0] : h->d_diff[2];
    const double* DB = el ? h->d_diff[1] : h->d_diff[3];
    if (!DA && !DB) return fail(RSG_ERR_STATE, "WPADIF before set_diffcoef");
    if (!sp.d_wtab) RET(h->dalloc(&sp.d_wtab, 3 * n));
    sp.sd.DA = sp.d_wtab;              // the column kernel reads (cA,cB) through DA and RL through DB
    sp.sd.DB = sp.d_wtab + 2 * n;
    if (sp.wtab_DTs == DTs) continue;
    CK(cudaMemsetAsync(h->d_wviol + s, 0, sizeof(unsigned long long), st));
    k_wpadif_tables<<<nblk((long long)h->NE * h->Pp, 128), 128, 0, st>>>(devfor(h, DTs), DA ? DA : h->d_zero4, DB ? DB : h->d_zero4,
                                                                        (double2*)sp.d_wtab, sp.d_wtab + 2 * n, h->d_wviol + s);
    CKL();
    h->launches++;
    sp.wtab_DTs = DTs;
  }
  return RSG_OK;
}
template <int EXT, bool PEER>
int L_col_t(rsg_ram* h, int s0, int ns, int doA, int doW, double DTs, cudaStream_t st, int b0, int nb, const PeerView& pv, int doC, int pb0) {
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  ColPlan c = col_plan(h, EXT >= 1);
  c.cfg.doA = doA;
  c.cfg.b0 = b0;
  c.cfg.pb0 = pb0;
  c.cfg.doW = doW;
  c.cfg.wpart_off = fused_wpart_off(h);
  c.cfg.doC = doC;
  c.cfg.tpos = h->T_elapsed > 0.0 ? 1 : 0;
  c.cfg.cpart_off = fused_cpart_off(h);
  c.cfg.NECR = h->d_NECR;
  if (nb < 0) nb = (h->P + COL_PG - 1) / COL_PG - b0;
  const dim3 g(nb, ns);
  const RamDev dv = devfor(h, DTs);
  if (c.T <= 320) {          // register budget follows the CTA size
    RET(opt_in_smem(k_col_fused<COL_PG, 320, EXT, PEER>, c.smem));
    k_col_fused<COL_PG, 320, EXT, PEER><<<g, c.T, c.smem, st>>>(dv, pk, s0, c.cfg, pv);
  } else if (c.T <= 576) {   // 576 = 2 x 288: the configs[2] shape; 112 registers, no spill in the WPI instantiation
    RET(opt_in_smem(k_col_fused<COL_PG, 576, EXT, PEER>, c.smem));
    k_col_fused<COL_PG, 576, EXT, PEER><<<g, c.T, c.smem, st>>>(dv, pk, s0, c.cfg, pv);
  } else if (c.T <= 896) {
    RET(opt_in_smem(k_col_fused<COL_PG, 896, EXT, PEER>, c.smem));
    k_col_fused<COL_PG, 896, EXT, PEER><<<g, c.T, c.smem, st>>>(dv, pk, s0, c.cfg, pv);
  } else {
    RET(opt_in_smem(k_col_fused<COL_PG, 1024, EXT, PEER>, c.smem));
    k_col_fused<COL_PG, 1024, EXT, PEER><<<g, c.T, c.smem, st>>>(dv, pk, s0, c.cfg, pv);
  }
  CKL();
  h->launches++;
  return RSG_OK;
}
int L_col(rsg_ram* h, int s0, int ns, int doA, double DTs, cudaStream_t st, int b0 = 0, int nb = -1, int doW = 0,
          const PeerView* peer = nullptr, int doC = 0, int pb0 = 0) {
  for (int s = s0; s < s0 + ns; ++s)
    if (!((doW >> s) & 1)) { h->sp[s].sd.DA = h->d_zero4; h->sp[s].sd.DB = h->d_zero4; }
  static const PeerView kNoPeer{};
  const bool ext = doW || doC;          // the extended instantiation: WPADIF and / or the Coulomb operators as extra stages
  if (peer)
    return doC ? L_col_t<2, true>(h, s0, ns, doA, doW, DTs, st, b0, nb, *peer, doC, pb0)
         : ext ? L_col_t<1, true>(h, s0, ns, doA, doW, DTs, st, b0, nb, *peer, 0, pb0) : L_col_t<0, true>(h, s0, ns, doA, 0, DTs, st, b0, nb, *peer, 0, pb0);
  return doC ? L_col_t<2, false>(h, s0, ns, doA, doW, DTs, st, b0, nb, kNoPeer, doC, pb0)
       : ext ? L_col_t<1, false>(h, s0, ns, doA, doW, DTs, st, b0, nb, kNoPeer, 0, pb0) : L_col_t<0, false>(h, s0, ns, doA, 0, DTs, st, b0, nb, kNoPeer, 0, pb0);
}
// elimination factors of the fused COULMU (k_coulmu_tables) for species [s0, s0+ns): tabulated when DTs (through COULPARA's
// rate tables), the fields or the plasmasphere changed, else cached.  Allocates: call outside stream capture.
int L_ctab(rsg_ram* h, int s0, int ns, double DTs, cudaStream_t st) {
  const size_t n = h->specStride;
  for (int s = s0; s < s0 + ns; ++s) {
    Spec& sp = h->sp[s];
    if (sp.DTs_coul < 0) return fail(RSG_ERR_STATE, "COULMU before COULPARA");
    if (!sp.d_ctab) RET(h->dalloc(&sp.d_ctab, 3 * n));
    sp.sd.CA = sp.d_ctab;
    sp.sd.CB = sp.d_ctab + 2 * n;
    if (sp.ctab_DTs == DTs) continue;
    const size_t nt = (size_t)h->NE * h->NPA;
    k_coulmu_tables<<<nblk((long long)h->NE * h->Pp, 128), 128, 0, st>>>(devfor(h, DTs), sp.d_coul + 2 * nt, sp.d_coul + 3 * nt, h->d_NECR,
                                                                        (double2*)sp.d_ctab, sp.d_ctab + 2 * n);
    CKL();
    h->launches++;
    sp.ctab_DTs = DTs;
  }
  return RSG_OK;
}
// pressures of ANISCH in one pass + the result block of the step, both also written to the
// host-mapped copies (no memcpy nodes in the fused step)
int L_finish_fused(rsg_ram* h, int s0, int ns, cudaStream_t st, int l0 = 0, int nl = -1, int nb_col = -1) {
  if (nl < 0) nl = h->NPA - l0;
  if (nb_col < 0) nb_col = (h->P + COL_PG - 1) / COL_PG;
  SpecPack pk;
  make_pack(h, pk, s0, ns);
  const double RFAC = 4 * kPI / (kCS * 100);
  const bool folded = h->sp[s0].d_aE2 != nullptr;     // ANISCH sums written by k_plane_rp<REV>
  if (!folded) {
    RET(prof_mark(h, "k_anisch", st));
    const int LCH = h->anischLch;
    const int nch = std::max(1, std::min(16, (nl + LCH - 1) / LCH));
    const int lch = (nl + nch - 1) / nch;
    k_anisch_pa_fast<<<dim3(nblk(h->Pp, 32), h->NE, ns), dim3(32, nch), 0, st>>>(h->dev, pk, s0, l0, nl, lch);
    CKL();
    h->launches++;
  }
  RET(prof_mark(h, "k_finalize", st));
  const PlanePlan c = plane_plan(h);
  const int KG = (h->NE + c.cfg.KC - 1) / c.cfg.KC;
  k_finalize<<<dim3(6 + nblk(h->P, 32), ns), 256, 0, st>>>(h->dev, pk, s0, nb_col, fused_part_off(h), KG * nl, h->d_cfl_all, RES_N, NSUM,
                                                           h->hd_res_all, RFAC, h->hd_pp_all, folded ? KG * nl : 0, l0);
  CKL();
  h->launches++;
  return RSG_OK;
}

int fetch_res(rsg_ram* h, int s, cudaStream_t st) {
  Spec& sp = h->sp[s];
  CK(cudaMemcpyAsync(sp.h_res, sp.d_res, sizeof(unsigned long long) * RES_N, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return RSG_OK;
}

// ---- host tables of one species (reference operation order) ------------------------
// DRIFTPARA (src/ModRamDrift.f90:36-88) + the energy-only prefixes of the sweeps
int tables_drift(rsg_ram* h, int s, double DTs, cudaStream_t st) {
  const int nS = h->nS, NR = h->NR, NE = h->NE, NPA = h->NPA;
  Spec& sp = h->sp[s];
  if (sp.DTs == DTs) return RSG_OK;   // pure function of DTs (and the grids)
  const double MDR = h->dev.MDR, DPHI = h->dev.DPHI, FracCFL = h->dev.FracCFL;
  const double QS = (double)h->QS[s];
  double* t = sp.h_tab;
  double *P4 = t, *eK = t + NE, *epK = t + 2 * NE, *aE = t + 3 * NE;
  double* P2 = t + 4 * NE;
  double* EDOT = P2 + (size_t)NE * NR;
  double* aMU = EDOT + (size_t)NE * NR;
  double* fast = aMU + NPA;  // w2[NE], wM[NE], tabE[NE][4]
  fast += (4 - ((fast - t) & 3)) & 3;  // tabE records are 32-byte aligned (d_tab is 256-byte aligned)
#define GRELs(K) h->GREL[s + (size_t)nS * ((K)-1)]
#define GRBNDs(K) h->GRBND[s + (size_t)nS * ((K)-1)]
  for (int K = 1; K <= NE; ++K) {
    // DRIFTR :131, DRIFTE :337 (energy-only prefix), DRIFTMU :424 (prefix)
    P4[K - 1] = DTs * h->EKEV[K - 1] * 1000.0 * (GRELs(K) + 1) / GRELs(K) / DPHI / MDR / QS;
    eK[K - 1] = h->EBND[K - 1] * 1e3 * (GRBNDs(K) + 1) / 2 / GRBNDs(K);
    epK[K - 1] = h->EKEV[K - 1] * 1e3 * (GRELs(K) + 1) / 2 / GRELs(K);
    aE[K - 1] = FracCFL * DTs * h->DE[K - 1];
    // FAST-mode energy factors
    {
      double* tabE = fast + 4 * (K - 1);
      const double uE = h->EBND[K - 1] * DTs * (GRBNDs(K) + 1) / GRBNDs(K) / 2.;               // EDOT = uE/RLZ
      tabE[0] = uE;
      tabE[1] = uE * eK[K - 1] / QS;                                                            // vE
      tabE[2] = 1.0 / h->DE[K - 1];
      tabE[3] = 1.0 / h->WE[K - 1];
      fast[4 * NE + K - 1] = DTs * h->EKEV[K - 1] * 1000 * (GRELs(K) + 1) / GRELs(K) / DPHI / QS;  // w2: P2 = w2/RLZ**2
      fast[5 * NE + K - 1] = epK[K - 1] / QS;                                                   // wM
    }
    for (int I = 1; I <= NR; ++I) {
      // DRIFTPARA :71, :83
      P2[(size_t)(K - 1) * NR + (I - 1)] =
          DTs * h->EKEV[K - 1] * 1000 * (GRELs(K) + 1) / GRELs(K) / (h->RLZ[I - 1] * h->RLZ[I - 1]) / DPHI / QS;
      EDOT[(size_t)(K - 1) * NR + (I - 1)] = h->EBND[K - 1] * DTs / h->RLZ[I - 1] * (GRBNDs(K) + 1) / GRBNDs(K) / 2.;
    }
  }
  for (int L = 1; L <= NPA; ++L) aMU[L - 1] = FracCFL * DTs * h->DMU[L - 1];
  CK(cudaMemcpyAsync(sp.d_tab, sp.h_tab, sp.n_tab * sizeof(double), cudaMemcpyHostToDevice, st));
  SpecDev& sd = sp.sd;
  // DRIFTE :310-311, :334-335
  const double EZERO = h->EKEV[0] - h->WE[0];
  sd.GRZERO = 1. + EZERO * 1000. * kQ / h->RMAS[s] / kCS / kCS;
  sd.GREL1 = GRELs(1);
  sd.GREL2 = GRELs(2);
  sd.sqrtA = std::sqrt((sd.GREL2 * sd.GREL2 - 1) / (sd.GREL1 * sd.GREL1 - 1));
  sd.sqrtB = std::sqrt((sd.GREL1 * sd.GREL1 - 1) / (sd.GRZERO * sd.GRZERO - 1));
  sd.aRP = FracCFL * DTs;
  sd.OMEt = OME_EARTH * DTs / DPHI;
  sp.DTs = DTs;
  h->inflow_ok[s] = h->cfl_ok[s] = false;   // P4 changed
  return RSG_OK;
}

// CEPARA (src/ModRamLoss.f90:19-170): energy-only factors
int tables_cepara(rsg_ram* h, int s, double DTs, cudaStream_t st) {
  const int nS = h->nS, NR = h->NR, NE = h->NE;
  Spec& sp = h->sp[s];
  if (sp.DTs_ce == DTs) return RSG_OK;   // pure function of DTs
  double* sv = sp.h_ce;
  double* ATLOS = sp.h_ce + NE;
  double* xATL = ATLOS + (size_t)NE * NR;
  const int kind = h->kind[s];
  for (int K = 1; K <= NE; ++K) {
    const double Vk = h->V[s + (size_t)nS * (K - 1)];
    double v = 0.0;
    if (K >= 2 && kind != RSG_KIND_E) {
      // :44-47, :59-62, :74-77
      double X = std::log10(h->EKEV[K - 1]);
      if (X < -2.) X = -2.;
      double Y;
      if (kind == RSG_KIND_H)
        Y = -18.767 - 0.11017 * X - 3.8173e-2 * (X * X) - 0.1232 * (X * X * X) - 5.0488e-2 * ((X * X) * (X * X));
      else if (kind == RSG_KIND_HE)
        Y = -20.789 + 0.92316 * X - 0.68017 * (X * X) + 0.66153 * (X * X * X) - 0.20998 * ((X * X) * (X * X));
      else
        Y = -18.987 - 0.10613 * X - 5.4841E-3 * (X * X) - 1.6262E-2 * (X * X * X) - 7.0554E-3 * ((X * X) * (X * X));
      v = std::pow(10., Y) * Vk;
    }
    sv[K - 1] = v;
    for (int I = 1; I <= NR; ++I) {
      double a = 1.0, xa = 0.0;
      if (K >= 2 && I >= 2) {
        const double TAUB = 2 * h->RLZ[I - 1] / Vk;  // :163-166
        xa = -DTs / TAUB;
        a = std::exp(xa);
      }
      ATLOS[(size_t)(K - 1) * NR + (I - 1)] = a;
      xATL[(size_t)(K - 1) * NR + (I - 1)] = xa;
    }
  }
  CK(cudaMemcpyAsync(sp.d_ce, sp.h_ce, ((size_t)NE + 2 * (size_t)NE * NR) * sizeof(double), cudaMemcpyHostToDevice, st));
  sp.DTs_ce = DTs;
  return RSG_OK;
}

// WAVELO factor table exp(-DTs/TAU_LIF) (src/ModRamWPI.f90:599-632, DoUsePlasmasphere=.false.)
int tables_wavelo(rsg_ram* h, int s, double DTs, cudaStream_t st) {
  const int NR = h->NR, NT = h->NT, NE = h->NE;
  Spec& sp = h->sp[s];
  if (sp.DTs_wl == DTs) return RSG_OK;
  if (h->WALOS1.empty()) return fail(RSG_ERR_STATE, "WAVELO before set_wavelo");
  double Bw = 30.;
  if (h->Kp >= 4.0) Bw = 100.;
  const double RLpp = 5.39 - 0.382 * h->Kpmax12;
  std::memset(sp.h_wfac, 0, (size_t)NE * h->Pp * sizeof(double));
  for (int K = 2; K <= NE; ++K)
    for (int I = 2; I <= NR; ++I) {
      const double W1 = h->WALOS1[(I - 1) + (size_t)NR * (K - 1)], W2 = h->WALOS2[(I - 1) + (size_t)NR * (K - 1)],
                   W3 = h->WALOS3[(I - 1) + (size_t)NR * (K - 1)];
      const double E = h->EKEV[K - 1];
      double TAU_LIF = 0.0;
      if (h->LZ[I - 1] <= RLpp) {
        TAU_LIF = W1 * ((10. / Bw) * (10. / Bw));
      } else {
        if (E <= 1000.) {
          TAU_LIF = W2 * (1 + W3 / W2);
          if (E <= 1.1) TAU_LIF = TAU_LIF * 37.5813 * std::exp(-1.81255 * E);
          else if (E > 1.1 && E <= 5.) TAU_LIF = TAU_LIF * (7.5 - 1.15 * E);
        } else {
          TAU_LIF = 5. * 3600 * 24 / h->Kp;
        }
      }
      const double fac = std::exp(-DTs / TAU_LIF);
      for (int J = 1; J <= NT; ++J) sp.h_wfac[(size_t)(K - 1) * h->Pp + (size_t)(J - 1) * NR + (I - 1)] = fac;
    }
  CK(cudaMemcpyAsync(sp.d_wfac, sp.h_wfac, (size_t)NE * h->Pp * sizeof(double), cudaMemcpyHostToDevice, st));
  sp.DTs_wl = DTs;
  return RSG_OK;
}

}  // namespace

// =============================================================================
extern "C" {

const char* rsg_last_error(void) { return g_err.c_str(); }

int rsg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int rsg_device_info(char* name, int name_len, int* sm_count, long long* l2_bytes, long long* hbm_bytes) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, dev));
  if (name && name_len > 0) {
    std::strncpy(name, p.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (l2_bytes) *l2_bytes = p.l2CacheSize;
  if (hbm_bytes) *hbm_bytes = (long long)p.totalGlobalMem;
  return RSG_OK;
}

int rsg_ram_create(rsg_ram** out, int nS, int NR, int NT, int NE, int NPA, int device) {
  if (!out) return fail(RSG_ERR_ARG, "null out");
  if (nS < 1 || nS > RSG_MAX_SPECIES || NR < 4 || NT < 4 || NE < 4 || NPA < 6) return fail(RSG_ERR_ARG, "bad dimensions");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return fail(RSG_ERR_CUDA, "no CUDA device");
  if (device >= 0) CK(cudaSetDevice(device));
  rsg_ram* h = new rsg_ram();
  CK(cudaGetDevice(&h->device));
  h->nS = nS; h->NR = NR; h->NT = NT; h->NE = NE; h->NPA = NPA;
  h->NR1 = NR + 1;
  h->P = NR * NT;
  h->Pp = (h->P + 15) / 16 * 16;
  if (const char* e = getenv("RSG_SEG")) h->segE = h->segMU = h->segP = std::max(2, atoi(e));
  if (const char* e = getenv("RSG_SEG_E")) h->segE = std::max(2, atoi(e));
  if (const char* e = getenv("RSG_SEG_MU")) h->segMU = std::max(2, atoi(e));
  if (const char* e = getenv("RSG_SEG_P")) h->segP = std::max(2, atoi(e));
  if (const char* e = getenv("RSG_KC_R")) h->kcR = std::max(1, atoi(e));
  if (getenv("RSG_NO_FUSE")) h->use_fused = false;
  if (getenv("RSG_NO_FUSE_WPI")) h->use_fused_wpi = false;
  if (const char* e = getenv("RSG_KC_PLANE")) h->kcPlane = std::max(1, atoi(e));
  if (const char* e = getenv("RSG_COL_T")) h->colT = std::max(32, atoi(e));
  if (const char* e = getenv("RSG_PLANE_T")) h->planeT = std::max(32, atoi(e));
  if (const char* e = getenv("RSG_PLANE_ODD")) h->planeOdd = atoi(e) != 0;
  if (const char* e = getenv("RSG_PLANE_TMA")) h->planeTma = atoi(e) != 0;
  // measured (profiles/r2/anisch_lch_sweep.txt): 12 at the default grid; at the configs[2] grid longer chunks keep more loads in
  // flight per thread (0.199 -> 0.171 ms with 18)
  h->anischLch = ((double)NR * NT * NE * NPA > 8e6) ? 18 : 12;
  if (const char* e = getenv("RSG_ANISCH_LCH")) h->anischLch = std::max(1, atoi(e));
  if (getenv("RSG_NO_GRAPH")) h->use_graph = false;   // kernel-by-kernel launches (profilers)
  RamDev& d = h->dev;
  d.nS = nS; d.NR = NR; d.NT = NT; d.NE = NE; d.NPA = NPA; d.NR1 = h->NR1; d.P = h->P; d.Pp = h->Pp;
  const size_t n2 = (size_t)h->NR1 * NT, n3 = n2 * NPA, np = h->Pp, n3p = (size_t)NPA * h->Pp;
  double** g1[] = {(double**)&d.RLZ, (double**)&d.EKEV, (double**)&d.WE, (double**)&d.DE, (double**)&d.MU, (double**)&d.WMU, (double**)&d.DMU};
  const size_t g1n[] = {(size_t)NR + 1, (size_t)NE, (size_t)NE, (size_t)NE, (size_t)NPA, (size_t)NPA, (size_t)NPA};
  for (int q = 0; q < 7; ++q) RET(h->dalloc(g1[q], g1n[q]));
  RET(h->dalloc((double**)&d.rDMU, NPA));
  RET(h->dalloc((double**)&d.rWMU, NPA));
  RET(h->dalloc((double**)&d.exp2tab, 64));
  RET(h->dalloc((double**)&d.wPE, NPA));
  RET(h->dalloc((double**)&d.wPA, NPA));
  RET(h->dalloc((int**)&d.UPA, NR));
  double** f2d[] = {(double**)&d.BNES, (double**)&d.dBdt, (double**)&d.VT, (double**)&d.EIR, (double**)&d.EIP};
  for (auto p : f2d) RET(h->dalloc(p, n2));
  double** f3d[] = {(double**)&d.FNHS, (double**)&d.FNIS, (double**)&d.BOUNHS, (double**)&d.BOUNIS, (double**)&d.HDNS, (double**)&d.dIdt, (double**)&d.dIbndt};
  for (auto p : f3d) RET(h->dalloc(p, n3));
  RET(h->dalloc((int**)&d.outside, (size_t)NR * NT));
  double** p2d[] = {&d.CR, &d.sB, &d.pT1, &d.pT3, &d.sBp, &d.DRD1, &d.DPD1, &d.BNESc, &d.dBdt2, &d.RLZp, &d.fPa};
  for (auto p : p2d) RET(h->dalloc(p, np));
  RET(h->dalloc(&d.outp, np));
  RET(h->dalloc(&d.CRt, np));
  RET(h->dalloc(&d.fRbt, n3p));
  double** p3d[] = {&d.t1, &d.G, &d.sFp, &d.Gr, &d.Gp, &d.DRD2, &d.DPD2, &d.dBdt1, &d.dIdt1, &d.FNHSc,
                    &d.CMUDOT, &d.Gmr, &d.Gmp, &d.DRM2, &d.DPM2, &d.dIbndt2, &d.BOUNHSc, &d.HDNSc,
                    &d.fRb, &d.fPb, &d.fEa, &d.fEb, &d.fMa, &d.fMb, &d.rFNHS};
  for (auto p : p3d) RET(h->dalloc(p, n3p));
  h->specStride = (size_t)NPA * NE * h->Pp;
  RET(h->dalloc(&h->d_F2[0], h->specStride * nS));
  RET(h->dalloc(&h->d_F2[1], h->specStride * nS));
  RET(h->dalloc(&h->d_stage, (size_t)nS * h->P * NE * NPA));
  RET(h->dalloc(&h->d_zero4, h->specStride));
  RET(h->dalloc(&h->d_NECR, (size_t)NR * NT));
  RET(h->dalloc(&h->d_dtinit, 4));
  RET(h->dalloc(&h->d_outlist, (size_t)NR * NT));
  {
    const double init[4] = {100000.0, 100000.0, 10000.0, 10000.0};  // :115,223,308,404
    RET(up(h->d_dtinit, init, 4));
  }
  h->ntiles = (NE * NPA * NT + SCAN_TILE - 1) / SCAN_TILE;
  RET(h->dalloc(&h->d_tilemax, (size_t)nS * h->ntiles));
  RET(h->dalloc(&h->d_res_all, (size_t)nS * RES_N));
  RET(h->dalloc(&h->d_res_init, (size_t)nS * RES_N));
  RET(h->dalloc(&h->d_cfl_all, (size_t)nS * 4));
  RET(h->dalloc(&h->d_wviol, (size_t)nS));
  {
    std::vector<unsigned long long> init((size_t)nS * RES_N, 0ull);
    const double dflt[4] = {100000.0, 100000.0, 10000.0, 10000.0};  // :115,223,308,404
    for (int s = 0; s < nS; ++s)
      for (int q = 0; q < 4; ++q) {
        std::memcpy(&init[(size_t)s * RES_N + q], &dflt[q], 8);
        std::memcpy(&init[(size_t)s * RES_N + DTF_OFF + q], &dflt[q], 8);
      }
    RET(up(h->d_res_init, init.data(), init.size()));
  }
  CK(cudaMallocHost((void**)&h->h_res_all, (size_t)nS * RES_N * sizeof(unsigned long long)));
  RET(h->dalloc(&h->d_pp_all, (size_t)nS * 2 * h->Pp));
  CK(cudaMallocHost((void**)&h->h_pp_all, (size_t)nS * 2 * h->Pp * sizeof(double)));
  CK(cudaHostGetDevicePointer((void**)&h->hd_res_all, h->h_res_all, 0));
  CK(cudaHostGetDevicePointer((void**)&h->hd_pp_all, h->h_pp_all, 0));
  CK(cudaStreamCreateWithFlags(&h->prepst, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->prepev, cudaEventDisableTiming));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
  // grid-stride reductions: a fixed grid => a fixed summation tree
  h->nblk_sum = NPA * NE;                       // one CTA per (K,L) plane
  h->sum_threads = h->P >= 1024 ? 256 : 128;
  for (int s = 0; s < nS; ++s) {
    Spec& sp = h->sp[s];
    CK(cudaStreamCreateWithFlags(&sp.own, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&sp.ev, cudaEventDisableTiming));
    sp.n_tab = (size_t)NE * 10 + (size_t)NE * NR * 2 + NPA + 4;
    RET(h->dalloc(&sp.d_tab, sp.n_tab));
    CK(cudaMallocHost((void**)&sp.h_tab, sp.n_tab * sizeof(double)));
    RET(h->dalloc(&sp.d_ce, (size_t)NE + 2 * (size_t)NE * NR));
    CK(cudaMallocHost((void**)&sp.h_ce, ((size_t)NE + 2 * (size_t)NE * NR) * sizeof(double)));
    RET(h->dalloc(&sp.d_wfac, (size_t)NE * h->Pp));
    CK(cudaMallocHost((void**)&sp.h_wfac, (size_t)NE * h->Pp * sizeof(double)));
    RET(h->dalloc(&sp.d_FF, (size_t)NPA * NE * NR));
    RET(h->dalloc(&sp.d_EPP, (size_t)NE));
    RET(h->dalloc(&sp.d_FGEOS, (size_t)NPA * NE * NT));
    RET(h->dalloc(&sp.d_last, (size_t)NE * NPA * NT));
    RET(h->dalloc(&sp.d_ghost, (size_t)2 * NE * NPA * NT));
    {
      // per-plane partials of the reductions, or per-warp partials of the sweeps with a fused SUMRC
      const size_t wR = (size_t)nblk(h->P, 248) * NPA * NE * 8;                       // KC = 1 worst case
      const size_t wM = (size_t)nblk(h->P, 128) * NE * ((NPA - 2 + 1) / 2) * 4;       // SEG = 2 worst case
      const size_t wF = (size_t)(nblk(h->P, 4) + 16) * 11 + (size_t)NE * NPA + 128;    // fused step: column + plane + WPADIF + Coulomb partials
      RET(h->dalloc(&sp.d_part, std::max({(size_t)h->nblk_sum * RSG_NMOM, wR, wM, wF})));
    }
    RET(h->dalloc(&sp.d_tE, (size_t)2 * NE * h->Pp));
    RET(h->dalloc(&sp.d_rFFA, (size_t)NE * NR));
    RET(h->dalloc(&sp.d_coul, (size_t)4 * NE * NPA + NE));
    sp.d_res = h->d_res_all + (size_t)s * RES_N;
    sp.h_res = h->h_res_all + (size_t)s * RES_N;
    sp.d_pp = h->d_pp_all + (size_t)s * 2 * h->Pp;
    SpecDev& sd = sp.sd;
    sd.S = s;
    sd.last = sp.d_last;
    sd.ghost = sp.d_ghost;
    sd.part = sp.d_part;
    sd.FGEOS = sp.d_FGEOS;
    double* t = sp.d_tab;
    sd.P4 = t; t += NE;
    sd.eK = t; t += NE;
    sd.epK = t; t += NE;
    sd.aE = t; t += NE;
    sd.sv = sp.d_ce;
    sd.ATLOS = sp.d_ce + NE;
    sd.xATL = sp.d_ce + NE + (size_t)NE * NR;
    sd.P2 = t; t += (size_t)NE * NR;
    sd.EDOT = t; t += (size_t)NE * NR;
    sd.aMU = t; t += NPA;
    t += (4 - ((t - sp.d_tab) & 3)) & 3;
    sd.tabE = t; t += 4 * (size_t)NE;
    sd.w2 = t; t += NE;
    sd.wM = t; t += NE;
    sd.FF = sp.d_FF;
    sd.EPP = sp.d_EPP;
    sd.wfac = sp.d_wfac;
    sd.DA = sd.DB = h->d_zero4;
    sd.tE = sp.d_tE;
    sd.tA = sp.d_tE + (size_t)NE * h->Pp;
    sd.rFFA = sp.d_rFFA;
    sd.pper = sp.d_pp;
    sd.ppar = sp.d_pp + h->Pp;
    sd.dt = sp.d_res;
    sd.dtw = sp.d_res;
  }
  *out = h;
  return RSG_OK;
}

int rsg_ram_destroy(rsg_ram* h) {
  if (!h) return RSG_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  shard_release(h);
  for (void* p : h->allocs) cudaFree(p);
  for (int s = 0; s < h->nS; ++s) {
    Spec& sp = h->sp[s];
    if (sp.own) cudaStreamDestroy(sp.own);
    if (sp.ev) cudaEventDestroy(sp.ev);
    if (sp.h_tab) cudaFreeHost(sp.h_tab);
    if (sp.h_ce) cudaFreeHost(sp.h_ce);
    if (sp.h_wfac) cudaFreeHost(sp.h_wfac);
  }
  if (h->h_res_all) cudaFreeHost(h->h_res_all);
  if (h->h_pp_all) cudaFreeHost(h->h_pp_all);
  if (h->prepst) cudaStreamDestroy(h->prepst);
  if (h->copyst) cudaStreamDestroy(h->copyst);
  for (auto& e : h->evh) if (e) cudaEventDestroy(e);
  if (h->prepev) cudaEventDestroy(h->prepev);
  if (h->t0) cudaEventDestroy(h->t0);
  if (h->t1) cudaEventDestroy(h->t1);
  if (h->gexec) cudaGraphExecDestroy(h->gexec);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  delete h;
  return RSG_OK;
}

int rsg_ram_set_mode(rsg_ram* h, int mode) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (mode != RSG_MODE_EXACT && mode != RSG_MODE_FAST) return fail(RSG_ERR_ARG, "unknown mode");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  h->mode = mode;
  for (int s = 0; s < h->nS; ++s) { h->sp[s].DTs = -1.0; h->inflow_ok[s] = h->cfl_ok[s] = false; }  // the inflow pre-pass depends on the mode
  if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
  return RSG_OK;
}

int rsg_ram_set_stream(rsg_ram* h, void* stream) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  CK(cudaDeviceSynchronize());
  h->ext = (cudaStream_t)stream;
  return RSG_OK;
}

int rsg_ram_use_graph(rsg_ram* h, int on) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  h->use_graph = on != 0;
  if (!on && h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
  return RSG_OK;
}

int rsg_ram_use_fused(rsg_ram* h, int on) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  h->use_fused = on != 0;
  h->use_fused_wpi = !(on & 2);        // on = 3: fused kernels for the default operators only, WPADIF as its own kernel
  if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
  return RSG_OK;
}

int rsg_ram_sync(rsg_ram* h) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (h->ext) CK(cudaStreamSynchronize(h->ext));
  for (int s = 0; s < h->nS; ++s) CK(cudaStreamSynchronize(h->sp[s].own));
  CK(cudaStreamSynchronize(h->prepst));
  return RSG_OK;
}

int rsg_ram_set_grids(rsg_ram* h, const double* RLZ, const double* LZ, const double* EKEV, const double* WE,
                      const double* DE, const double* EBND, const double* MU, const double* WMU, const double* DMU,
                      const double* UPA, const double* GREL, const double* GRBND, const double* V,
                      const double* VBND, const double* EPP, const double* ERNH, const double* RMAS,
                      const double* FFACTOR, const int* QS, const int* kind, const int* khi, double MDR,
                      double DPHI, double CONF1, double CONF2, double BetaLim, double FracCFL) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (!RLZ || !LZ || !EKEV || !WE || !DE || !EBND || !MU || !WMU || !DMU || !UPA || !GREL || !GRBND || !V || !VBND || !EPP ||
      !ERNH || !RMAS || !FFACTOR || !QS || !kind || !khi)
    return fail(RSG_ERR_ARG, "null grid pointer");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const int nS = h->nS, NR = h->NR, NE = h->NE, NPA = h->NPA;
  h->RLZ.assign(RLZ, RLZ + NR + 1); h->LZ.assign(LZ, LZ + NR + 1);
  h->EKEV.assign(EKEV, EKEV + NE); h->WE.assign(WE, WE + NE); h->DE.assign(DE, DE + NE); h->EBND.assign(EBND, EBND + NE);
  h->MU.assign(MU, MU + NPA); h->WMU.assign(WMU, WMU + NPA); h->DMU.assign(DMU, DMU + NPA);
  h->UPA.assign(UPA, UPA + NR);
  h->GREL.assign(GREL, GREL + (size_t)nS * NE); h->GRBND.assign(GRBND, GRBND + (size_t)nS * NE);
  h->V.assign(V, V + (size_t)nS * NE); h->VBND.assign(VBND, VBND + (size_t)nS * NE);
  h->EPP.assign(EPP, EPP + (size_t)nS * NE); h->ERNH.assign(ERNH, ERNH + (size_t)nS * NE);
  h->RMAS.assign(RMAS, RMAS + nS);
  h->QS.assign(QS, QS + nS); h->kind.assign(kind, kind + nS); h->khi.assign(khi, khi + 5);
  RamDev& d = h->dev;
  d.MDR = MDR; d.DPHI = DPHI; d.CONF1 = CONF1; d.CONF2 = CONF2; d.BetaLim = BetaLim; d.FracCFL = FracCFL;
  RET(up((double*)d.RLZ, RLZ, NR + 1)); RET(up((double*)d.EKEV, EKEV, NE)); RET(up((double*)d.WE, WE, NE));
  RET(up((double*)d.DE, DE, NE)); RET(up((double*)d.MU, MU, NPA)); RET(up((double*)d.WMU, WMU, NPA));
  RET(up((double*)d.DMU, DMU, NPA));
  {
    std::vector<double> r1(NPA), r2(NPA), w1(NPA), w2(NPA);
    for (int l = 0; l < NPA; ++l) {
      r1[l] = 1.0 / DMU[l];
      r2[l] = 1.0 / WMU[l];
      // ANISCH :372-374 with FFACTOR(S,I,K,L) = A(S,I,K)*MU(L), FFACTOR(..,1) = FFACTOR(..,2) (src/ModRamInit.f90:561-569)
      const double mueff = (l == 0) ? MU[1] : MU[l];
      w1[l] = WMU[l] / mueff * (1.0 - MU[l] * MU[l]);
      w2[l] = WMU[l] / mueff * (MU[l] * MU[l]);
    }
    double e2[64];
    for (int j = 0; j < 64; ++j) e2[j] = std::exp2((double)j / 64.0);
    RET(up((double*)d.exp2tab, e2, 64));
    RET(up((double*)d.rDMU, r1.data(), NPA));
    RET(up((double*)d.rWMU, r2.data(), NPA));
    RET(up((double*)d.wPE, w1.data(), NPA));
    RET(up((double*)d.wPA, w2.data(), NPA));
  }
  std::vector<int> upa(NR);
  for (int i = 0; i < NR; ++i) upa[i] = (int)UPA[i];
  RET(up((int*)d.UPA, upa.data(), NR));
  std::vector<double> ff((size_t)NPA * NE * NR), epp(NE);
  for (int s = 0; s < nS; ++s) {
    for (int l = 0; l < NPA; ++l)
      for (int k = 0; k < NE; ++k)
        for (int i = 0; i < NR; ++i)
          ff[((size_t)l * NE + k) * NR + i] = FFACTOR[s + (size_t)nS * (i + (size_t)NR * (k + (size_t)NE * l))];
    RET(up(h->sp[s].d_FF, ff.data(), ff.size()));
    {
      std::vector<double> ra((size_t)NE * NR);
      for (int k = 0; k < NE; ++k)
        for (int i = 0; i < NR; ++i) ra[(size_t)k * NR + i] = MU[1] / ff[((size_t)1 * NE + k) * NR + i];
      RET(up(h->sp[s].d_rFFA, ra.data(), ra.size()));
    }
    for (int k = 0; k < NE; ++k) epp[k] = EPP[s + (size_t)nS * k];
    RET(up(h->sp[s].d_EPP, epp.data(), NE));
    h->sp[s].sd.kind = kind[s];
    h->sp[s].sd.QS = (double)QS[s];
    h->sp[s].DTs = h->sp[s].DTs_ce = h->sp[s].DTs_wl = -1.0;
  }
  h->grids_set = true;
  h->step_dirty = true;
  return RSG_OK;
}

int rsg_ram_set_fields(rsg_ram* h, const double* BNES, const double* dBdt, const double* FNHS, const double* FNIS,
                       const double* BOUNHS, const double* BOUNIS, const double* HDNS, const double* dIdt,
                       const double* dIbndt, const int* outsideMGNP) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (!BNES || !dBdt || !FNHS || !FNIS || !BOUNHS || !BOUNIS || !HDNS || !dIdt || !dIbndt || !outsideMGNP)
    return fail(RSG_ERR_ARG, "null field pointer");
  if (!h->grids_set) return fail(RSG_ERR_STATE, "set_fields before set_grids");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  RamDev& d = h->dev;
  const size_t n2 = (size_t)h->NR1 * h->NT, n3 = n2 * h->NPA;
  RET(up((double*)d.BNES, BNES, n2)); RET(up((double*)d.dBdt, dBdt, n2));
  RET(up((double*)d.FNHS, FNHS, n3)); RET(up((double*)d.FNIS, FNIS, n3)); RET(up((double*)d.BOUNHS, BOUNHS, n3));
  RET(up((double*)d.BOUNIS, BOUNIS, n3)); RET(up((double*)d.HDNS, HDNS, n3)); RET(up((double*)d.dIdt, dIdt, n3));
  RET(up((double*)d.dIbndt, dIbndt, n3));
  RET(up((int*)d.outside, outsideMGNP, (size_t)h->NR * h->NT));
  {
    std::vector<int> lst;
    for (int j = 1; j < h->NT - 1; ++j)
      for (int i = 0; i < h->NR; ++i)
        if (outsideMGNP[(size_t)j * h->NR + i] != 0) lst.push_back(j * h->NR + i);
    h->nout = (int)lst.size();
    if (h->nout) RET(up(h->d_outlist, lst.data(), lst.size()));
  }
  k_prep_fields<<<nblk((long long)h->NPA * h->Pp, 256), 256, 0, h->pst()>>>(d);
  CKL();
  h->launches++;
  CK(cudaStreamSynchronize(h->pst()));
  h->fields_set = true;
  h->step_dirty = true;
  for (int s = 0; s < h->nS; ++s) h->sp[s].wtab_DTs = h->sp[s].ctab_DTs = -1.0;   // FACMU = FNHS*MU enters the WPADIF / COULMU factors
  return RSG_OK;
}

// rsg_ram_set_fields with DEVICE pointers (what rsg_hi_device_fields returns: the resident computehI, same process and
// device): nine device-to-device copies instead of the host round trip.  Order: BNES, dBdt, FNHS, FNIS, BOUNHS, BOUNIS,
// HDNS, dIdt, dIbndt.  The list of lines beyond the magnetopause (the step's DRIFTP repair) is rebuilt from the flags.
int rsg_ram_set_fields_device(rsg_ram* h, const double* const* ptrs9, const int* d_outsideMGNP) {
  if (!h || !ptrs9 || !d_outsideMGNP) return fail(RSG_ERR_ARG, "null argument");
  for (int q = 0; q < 9; ++q)
    if (!ptrs9[q]) return fail(RSG_ERR_ARG, "null field pointer");
  if (!h->grids_set) return fail(RSG_ERR_STATE, "set_fields before set_grids");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  RamDev& d = h->dev;
  const size_t n2 = (size_t)h->NR1 * h->NT, n3 = n2 * h->NPA;
  double* dst[9] = {(double*)d.BNES, (double*)d.dBdt, (double*)d.FNHS, (double*)d.FNIS, (double*)d.BOUNHS, (double*)d.BOUNIS,
                    (double*)d.HDNS, (double*)d.dIdt, (double*)d.dIbndt};
  cudaStream_t st = h->pst();
  for (int q = 0; q < 9; ++q) CK(cudaMemcpyAsync(dst[q], ptrs9[q], (q < 2 ? n2 : n3) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync((int*)d.outside, d_outsideMGNP, (size_t)h->NR * h->NT * sizeof(int), cudaMemcpyDeviceToDevice, st));
  std::vector<int> om((size_t)h->NR * h->NT);
  CK(cudaMemcpyAsync(om.data(), d_outsideMGNP, om.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  {
    std::vector<int> lst;
    for (int j = 1; j < h->NT - 1; ++j)
      for (int i = 0; i < h->NR; ++i)
        if (om[(size_t)j * h->NR + i] != 0) lst.push_back(j * h->NR + i);
    h->nout = (int)lst.size();
    if (h->nout) RET(up(h->d_outlist, lst.data(), lst.size()));
  }
  k_prep_fields<<<nblk((long long)h->NPA * h->Pp, 256), 256, 0, st>>>(d);
  CKL();
  h->launches++;
  CK(cudaStreamSynchronize(st));
  h->fields_set = true;
  h->step_dirty = true;
  for (int s = 0; s < h->nS; ++s) h->sp[s].wtab_DTs = h->sp[s].ctab_DTs = -1.0;
  return RSG_OK;
}

int rsg_ram_set_efield(rsg_ram* h, const double* VT, const double* EIR, const double* EIP) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (!VT || !EIR || !EIP) return fail(RSG_ERR_ARG, "null e-field pointer");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const size_t n2 = (size_t)h->NR1 * h->NT;
  RET(up((double*)h->dev.VT, VT, n2)); RET(up((double*)h->dev.EIR, EIR, n2)); RET(up((double*)h->dev.EIP, EIP, n2));
  h->efield_set = true;
  h->step_dirty = true;
  return RSG_OK;
}

int rsg_ram_set_boundary(rsg_ram* h, const double* FGEOS) {
  if (!h || !FGEOS) return fail(RSG_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const int nS = h->nS, NT = h->NT, NE = h->NE, NPA = h->NPA;
  std::vector<double> b((size_t)NPA * NE * NT);
  for (int s = 0; s < nS; ++s) {
    for (int l = 0; l < NPA; ++l)
      for (int k = 0; k < NE; ++k)
        for (int j = 0; j < NT; ++j) b[((size_t)l * NE + k) * NT + j] = FGEOS[s + (size_t)nS * (j + (size_t)NT * (k + (size_t)NE * l))];
    RET(up(h->sp[s].d_FGEOS, b.data(), b.size()));
    h->inflow_ok[s] = h->cfl_ok[s] = false;
  }
  return RSG_OK;
}

int rsg_ram_set_wavelo(rsg_ram* h, const double* W1, const double* W2, const double* W3, double Kp, double Kpmax12) {
  if (!h || !W1 || !W2 || !W3) return fail(RSG_ERR_ARG, "null argument");
  const size_t n = (size_t)h->NR * h->NE;
  h->WALOS1.assign(W1, W1 + n); h->WALOS2.assign(W2, W2 + n); h->WALOS3.assign(W3, W3 + n);
  h->Kp = Kp; h->Kpmax12 = Kpmax12;
  for (int s = 0; s < h->nS; ++s) h->sp[s].DTs_wl = -1.0;
  return RSG_OK;
}

int rsg_ram_set_plasmasphere(rsg_ram* h, const double* NECR) {
  if (!h || !NECR) return fail(RSG_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  RET(up(h->d_NECR, NECR, (size_t)h->NR * h->NT));
  for (int s = 0; s < h->nS; ++s) h->sp[s].ctab_DTs = -1.0;
  return RSG_OK;
}

int rsg_ram_set_diffcoef(rsg_ram* h, int which, const double* D) {
  if (!h || !D) return fail(RSG_ERR_ARG, "null argument");
  if (which < 0 || which > 3) return fail(RSG_ERR_ARG, "which must be 0..3");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  if (!h->d_diff[which]) {
    RET(h->dalloc(&h->d_diff[which], h->specStride));
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }   // a kernel argument changes
  }
  std::vector<double> b;
  to_planes4(h, D, b);
  RET(up(h->d_diff[which], b.data(), b.size()));
  for (int s = 0; s < h->nS; ++s) h->sp[s].wtab_DTs = -1.0;
  return RSG_OK;
}

// FLC_coef(S,:,:,:,:) of one species, as a contiguous (NR,NT,NE,NPA) array (the output of PARA_FLC)
int rsg_ram_set_flc_coef(rsg_ram* h, int S, const double* D) {
  RET(check_S(h, S));
  if (!D) return fail(RSG_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  Spec& sp = h->sp[S - 1];
  if (!sp.d_flc) RET(h->dalloc(&sp.d_flc, h->specStride));
  std::vector<double> b;
  to_planes4(h, D, b);
  RET(up(sp.d_flc, b.data(), b.size()));
  return RSG_OK;
}

// PARA_FLC(S) (src/ModRamLoss.f90:342-455) on the device: r_curvEq, zeta1Eq, zeta2Eq are the (NR,NT)
// outputs of FLC_Radius (host: SCB geometry + 2-D interpolation, once per Dt_bc); the species'
// FLC_coef is built where FLCscatter reads it.  The caller keeps the reference's "every Dt_bc" gate (:371).
int rsg_para_flc(rsg_ram* h, int S, const double* r_curvEq, const double* zeta1Eq, const double* zeta2Eq) {
  RET(check_S(h, S));
  if (!r_curvEq || !zeta1Eq || !zeta2Eq) return fail(RSG_ERR_ARG, "null argument");
  if (!h->grids_set || !h->fields_set) return fail(RSG_ERR_STATE, "PARA_FLC before set_grids / set_fields");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const int s = S - 1, P = h->P, NE = h->NE, NR = h->NR;
  Spec& sp = h->sp[s];
  if (!sp.d_flc) RET(h->dalloc(&sp.d_flc, h->specStride));
  if (!h->d_flctab) RET(h->dalloc(&h->d_flctab, (size_t)3 * P + NE + NR));
  std::vector<double> tab((size_t)3 * P + NE + NR);
  std::memcpy(&tab[0], r_curvEq, sizeof(double) * P);
  std::memcpy(&tab[P], zeta1Eq, sizeof(double) * P);
  std::memcpy(&tab[2 * (size_t)P], zeta2Eq, sizeof(double) * P);
  for (int k = 0; k < NE; ++k) tab[3 * (size_t)P + k] = h->V[s + (size_t)h->nS * k];
  for (int i = 0; i < NR; ++i) tab[3 * (size_t)P + NE + i] = h->LZ[i];
  RET(up(h->d_flctab, tab.data(), tab.size()));
  cudaStream_t st = h->st(s);
  CK(cudaMemsetAsync(sp.d_flc, 0, sizeof(double) * h->specStride, st));
  k_para_flc<<<nblk((long long)P * NE, 128), 128, 0, st>>>(h->dev, h->d_flctab, h->RMAS[s], sp.d_flc);
  CKL();
  h->launches++;
  CK(cudaStreamSynchronize(st));
  return RSG_OK;
}
// the species' FLC_coef as the reference holds it: contiguous (NR,NT,NE,NPA) (diagnostics / tests)
int rsg_ram_get_flc_coef(rsg_ram* h, int S, double* D) {
  RET(check_S(h, S));
  if (!D) return fail(RSG_ERR_ARG, "null argument");
  Spec& sp = h->sp[S - 1];
  if (!sp.d_flc) return fail(RSG_ERR_STATE, "no FLC_coef on the device");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  std::vector<double> b(h->specStride);
  CK(cudaMemcpy(b.data(), sp.d_flc, sizeof(double) * h->specStride, cudaMemcpyDeviceToHost));
  for (int l = 0; l < h->NPA; ++l)
    for (int k = 0; k < h->NE; ++k)
      std::memcpy(&D[((size_t)l * h->NE + k) * h->P], &b[((size_t)l * h->NE + k) * h->Pp], sizeof(double) * h->P);
  return RSG_OK;
}

// ---- the boundary / E-field producers ------------------------------------------------------------
// GEOSB(S) (src/ModRamBoundary.f90:241-319, boundary 'LANL'): FluxLanl(NT,NE) of get_geomlt_flux and species%s_comp ->
// the species' FGEOS on the device (replaces the nS x NT x NE x NPA upload of rsg_ram_set_boundary for that species).
int rsg_geosb(rsg_ram* h, int S, const double* FluxLanl, double s_comp) {
  RET(check_S(h, S));
  if (!FluxLanl) return fail(RSG_ERR_ARG, "null argument");
  if (!h->grids_set) return fail(RSG_ERR_STATE, "GEOSB before set_grids");
  CK(cudaSetDevice(h->device));
  const int s = S - 1;
  cudaStream_t st = h->st(s);
  double* d_flux = nullptr;
  const size_t n = (size_t)h->NT * h->NE;
  CK(cudaMalloc((void**)&d_flux, n * sizeof(double)));
  CK(cudaMemcpyAsync(d_flux, FluxLanl, n * sizeof(double), cudaMemcpyHostToDevice, st));
  k_geosb<<<dim3(nblk((long long)n, 128), h->NPA), 128, 0, st>>>(h->dev, d_flux, s_comp, h->sp[s].d_FF, h->sp[s].d_FGEOS);
  CKL();
  h->launches++;
  CK(cudaStreamSynchronize(st));
  cudaFree(d_flux);
  h->inflow_ok[s] = h->cfl_ok[s] = false;
  return RSG_OK;
}
int rsg_ram_get_boundary(rsg_ram* h, int S, double* FGEOS_S) {          // (NT,NE,NPA) of species S (diagnostics / tests)
  RET(check_S(h, S));
  if (!FGEOS_S) return fail(RSG_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const int NT = h->NT, NE = h->NE, NPA = h->NPA;
  std::vector<double> b((size_t)NPA * NE * NT);
  CK(cudaMemcpy(b.data(), h->sp[S - 1].d_FGEOS, b.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int l = 0; l < NPA; ++l)
    for (int k = 0; k < NE; ++k)
      for (int j = 0; j < NT; ++j) FGEOS_S[j + (size_t)NT * (k + (size_t)NE * l)] = b[((size_t)l * NE + k) * NT + j];
  return RSG_OK;
}
// get_electric_field (src/ModRamEField.f90:14-63) on the device: mode 0 interpolates VT in time between two potential
// maps VTOL, VTN (NR+1,NT); mode 1 is the Volland-Stern potential from Kp, LZ(NR+1), PHI(NT), PHIOFS.  EIR / EIP keep the
// values of the last rsg_ram_set_efield.  VT_out (may be NULL) returns VT.
int rsg_get_electric_field(rsg_ram* h, int vols, const double* VTOL, const double* VTN, double TimeRamElapsed, double TOLV,
                           double DtEfi, double Kp, const double* PHI, double PHIOFS, double* VT_out) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (!h->grids_set) return fail(RSG_ERR_STATE, "get_electric_field before set_grids");
  if (vols ? !PHI : (!VTOL || !VTN)) return fail(RSG_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const size_t n2 = (size_t)h->NR1 * h->NT;
  double *da = nullptr, *db = nullptr;
  cudaStream_t st = h->pst();
  double p0, p1, p2 = 1.0;
  if (!vols) {
    CK(cudaMalloc((void**)&da, n2 * sizeof(double)));
    CK(cudaMalloc((void**)&db, n2 * sizeof(double)));
    CK(cudaMemcpyAsync(da, VTOL, n2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(db, VTN, n2 * sizeof(double), cudaMemcpyHostToDevice, st));
    p0 = TimeRamElapsed; p1 = TOLV; p2 = DtEfi;
  } else {
    const double RE = 6.371E6;
    const double q = 1. - 0.159 * Kp + 0.0093 * (Kp * Kp);
    std::vector<double> sn(h->NT);
    for (int j = 0; j < h->NT; ++j) sn[j] = std::sin(PHI[j] - PHIOFS);
    CK(cudaMalloc((void**)&da, (h->NR1) * sizeof(double)));
    CK(cudaMalloc((void**)&db, h->NT * sizeof(double)));
    CK(cudaMemcpyAsync(da, h->LZ.data(), h->NR1 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(db, sn.data(), h->NT * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    p0 = 7.05E-6 / (q * q * q) / RE; p1 = RE;
  }
  k_efield<<<nblk((long long)n2, 128), 128, 0, st>>>(h->dev, vols, da, db, p0, p1, p2, (double*)h->dev.VT);
  CKL();
  h->launches++;
  if (VT_out) CK(cudaMemcpyAsync(VT_out, h->dev.VT, n2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(da); cudaFree(db);
  if (!h->efield_set) {                       // EIR, EIP stay zero until rsg_ram_set_efield provides them
    h->efield_set = true;
  }
  h->step_dirty = true;
  return RSG_OK;
}

// ---- ANISCH, second half: rebuild of the WPADIF diffusion coefficients on the device (src/ModRamRun.f90:422-605) ----
int rsg_ram_set_wave_tables(rsg_ram* h, int ENG, int NCF, const double* ENOR, const double* fpofc, const double* NDAAJ,
                            const double* DAAR, int use_bas, int ENG_emic, int NCF_emic, const double* EKEV_emic,
                            const double* fp2c_emic, const double* Daa_emic_h, const double* Daa_emic_he, const double* Ihs_emic,
                            const double* Ihes_emic, const double* PAbn) {
  if (!h || !PAbn) return fail(RSG_ERR_ARG, "null argument");
  if (!h->grids_set) return fail(RSG_ERR_STATE, "set_wave_tables before set_grids");
  const bool wpi = ENOR && fpofc && NDAAJ && DAAR, emic = EKEV_emic && fp2c_emic && Daa_emic_h && Daa_emic_he && Ihs_emic && Ihes_emic;
  if (!wpi && !emic) return fail(RSG_ERR_ARG, "neither the WPI (ENOR, fpofc, NDAAJ, C/BDAAR) nor the EMIC tables are complete");
  if ((wpi && (ENG < 2 || NCF < 2)) || (emic && (ENG_emic < 2 || NCF_emic < 2))) return fail(RSG_ERR_ARG, "tables need >= 2 nodes per axis");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const int NR = h->NR, NT = h->NT, NE = h->NE, NPA = h->NPA;
  DiffTabs& t = h->dtab;
  t = DiffTabs{};
  auto upv = [&](const double** dst, const std::vector<double>& v) -> int {
    double* p = nullptr;
    RET(h->dalloc(&p, v.size()));
    RET(up(p, v.data(), v.size()));
    *dst = p;
    return RSG_OK;
  };
  auto upa = [&](const double** dst, const double* src, size_t n) -> int { return upv(dst, std::vector<double>(src, src + n)); };
  // abscissae of the chorus interpolation: PA(L) = ACOSD(MU(NPA-L+1)) (:437), through the wrapper's monotonicity filter
  {
    std::vector<double> xa;
    std::vector<int> idx;
    for (int L = 1; L <= NPA; ++L) {
      const double pa = 180.0 / kPI * std::acos(h->MU[NPA - L]);
      if (xa.empty() || pa > xa.back()) { xa.push_back(pa); idx.push_back(L - 1); }
    }
    t.n1 = (int)xa.size();
    RET(upv(&t.PAx, xa));
    int* pi = nullptr;
    RET(h->dalloc(&pi, idx.size()));
    RET(up(pi, idx.data(), idx.size()));
    t.PAidx = pi;
  }
  RET(upa(&t.PAbn, PAbn, NPA));
  if (wpi) {
    t.ENG = ENG; t.NCF = NCF; t.use_bas = use_bas ? 1 : 0;
    std::vector<double> al(ENG);
    for (int q = 0; q < ENG; ++q) al[q] = std::log10(ENOR[q]);
    RET(upv(&t.ALENOR, al));
    RET(upa(&t.fpofc, fpofc, NCF));
    RET(upa(&t.NDAAJ, NDAAJ, (size_t)NR * ENG * NPA * NCF));
    RET(upa(&t.DAAR, DAAR, (size_t)NR * NT * NE * NPA));
  }
  if (emic) {
    t.ENGe = ENG_emic; t.NCFe = NCF_emic;
    std::vector<double> al(ENG_emic);
    for (int q = 0; q < ENG_emic; ++q) al[q] = std::log10(EKEV_emic[q]);
    RET(upv(&t.logEe, al));
    RET(upa(&t.fp2c, fp2c_emic, NCF_emic));
    RET(upa(&t.DH, Daa_emic_h, (size_t)NR * ENG_emic * NPA * NCF_emic));
    RET(upa(&t.DHE, Daa_emic_he, (size_t)NR * ENG_emic * NPA * NCF_emic));
    RET(upa(&t.Ihs, Ihs_emic, (size_t)4 * NR * NT));
    RET(upa(&t.Ihes, Ihes_emic, (size_t)4 * NR * NT));
  }
  if (!h->d_XNE) RET(h->dalloc(&h->d_XNE, (size_t)NR * NT));
  if (!h->d_dcerr) RET(h->dalloc(&h->d_dcerr, 1));
  h->dtab_set = true;
  return RSG_OK;
}

// S: 1-based species; flags: RSG_F_WPI (electrons: ATAW + ATAC) | RSG_F_EMIC (H+: ATAW_emic_h + ATAW_emic_he); XNE(NR,NT):
// plasmaspheric electron density; AE index (I_emic).  Kp comes from rsg_ram_set_wavelo.  The caller keeps the reference's
// "every Dt_bc" gate (:422, :520).  *gslerr: lines whose 1-D interpolation failed (GSLerr of the reference).
int rsg_anisch_diffcoef(rsg_ram* h, int S, int flags, const double* XNE, int AE, int* gslerr) {
  RET(check_S(h, S));
  if (!XNE) return fail(RSG_ERR_ARG, "null argument");
  if (!h->dtab_set) return fail(RSG_ERR_STATE, "ANISCH diffusion coefficients before rsg_ram_set_wave_tables");
  if (!h->fields_set) return fail(RSG_ERR_STATE, "ANISCH diffusion coefficients before set_fields");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const int s = S - 1, NR = h->NR, NT = h->NT, NE = h->NE, NPA = h->NPA;
  RET(up(h->d_XNE, XNE, (size_t)NR * NT));
  DiffTabs t = h->dtab;
  t.XNE = h->d_XNE;
  t.RMASs = h->RMAS[s];
  t.RMASe = h->RMAS[s];
  for (int q = 0; q < h->nS; ++q)
    if (h->kind[q] == RSG_KIND_E) t.RMASe = h->RMAS[q];       // RMAS(4) of the reference: the electrons
  t.Bw = (h->Kp >= 4.0) ? 100. : 30.;
  t.cls = (AE >= 0 && AE < 100) ? 1 : ((AE >= 100 && AE < 300) ? 2 : ((AE >= 300 && AE < 400) ? 3 : (AE >= 400 ? 4 : 0)));
  {
    std::vector<double> g(NE);
    for (int k = 0; k < NE; ++k) g[k] = h->GREL[s + (size_t)h->nS * k];
    if (!h->d_dcgrel) RET(h->dalloc(&h->d_dcgrel, NE));
    RET(up(h->d_dcgrel, g.data(), NE));
    t.GRELs = h->d_dcgrel;
  }
  cudaStream_t st = h->st(s);
  CK(cudaMemsetAsync(h->d_dcerr, 0, sizeof(int), st));
  auto ensure = [&](int which) -> int {
    if (!h->d_diff[which]) {
      RET(h->dalloc(&h->d_diff[which], h->specStride));
      if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    }
    CK(cudaMemsetAsync(h->d_diff[which], 0, sizeof(double) * h->specStride, st));
    return RSG_OK;
  };
  bool did = false;
  if ((flags & RSG_F_WPI) && h->kind[s] == RSG_KIND_E) {
    if (!t.DAAR) return fail(RSG_ERR_STATE, "the WPI tables were not set");
    RET(ensure(0)); RET(ensure(1));
    const long long nlines = (long long)(NR - 1) * NT * (NE - 1);
    k_diffcoef_chorus<<<nblk(nlines, 4), 128, sizeof(double) * 4 * 2 * NPA, st>>>(h->dev, t, h->d_diff[1], h->d_dcerr);
    CKL();
    k_diffcoef_bilinear<<<nblk(nlines * NPA, 128), 128, 0, st>>>(h->dev, t, 0, h->d_diff[0], nullptr);
    CKL();
    h->launches += 2;
    did = true;
  }
  if ((flags & RSG_F_EMIC) && h->kind[s] == RSG_KIND_H) {
    if (!t.DH) return fail(RSG_ERR_STATE, "the EMIC tables were not set");
    RET(ensure(2)); RET(ensure(3));
    const long long n = (long long)(NR - 1) * NT * (NE - 1) * NPA;
    k_diffcoef_bilinear<<<nblk(n, 128), 128, 0, st>>>(h->dev, t, 1, h->d_diff[2], h->d_diff[3]);
    CKL();
    h->launches++;
    did = true;
  }
  int err = 0;
  CK(cudaMemcpyAsync(&err, h->d_dcerr, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (gslerr) *gslerr = err;
  if (did)
    for (int q = 0; q < h->nS; ++q) h->sp[q].wtab_DTs = -1.0;     // the fused WPADIF factors are stale
  return RSG_OK;
}

// which = 0 ATAW, 1 ATAC, 2 ATAW_emic_h, 3 ATAW_emic_he as the reference holds them: (NR,NT,NE,NPA)
int rsg_ram_get_diffcoef(rsg_ram* h, int which, double* D) {
  if (!h || !D) return fail(RSG_ERR_ARG, "null argument");
  if (which < 0 || which > 3) return fail(RSG_ERR_ARG, "which must be 0..3");
  if (!h->d_diff[which]) return fail(RSG_ERR_STATE, "this diffusion coefficient is not on the device");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  std::vector<double> b(h->specStride);
  CK(cudaMemcpy(b.data(), h->d_diff[which], sizeof(double) * h->specStride, cudaMemcpyDeviceToHost));
  for (int l = 0; l < h->NPA; ++l)
    for (int k = 0; k < h->NE; ++k)
      std::memcpy(&D[((size_t)l * h->NE + k) * h->P], &b[((size_t)l * h->NE + k) * h->Pp], sizeof(double) * h->P);
  return RSG_OK;
}

// ---- F2 transfers -------------------------------------------------------------
int rsg_ram_f2_h2d(rsg_ram* h, const double* F2, int S) {
  if (!h || !F2) return fail(RSG_ERR_ARG, "null argument");
  if (S < 0 || S > h->nS) return fail(RSG_ERR_ARG, "species index out of range");
  CK(cudaSetDevice(h->device));
  const size_t n = (size_t)h->nS * h->P * h->NE * h->NPA;
  if (S == 0) RET(rsg_ram_sync(h));
  cudaStream_t st = S ? h->st(S - 1) : h->pst();
  CK(cudaMemcpyAsync(h->d_stage, F2, n * sizeof(double), cudaMemcpyHostToDevice, st));
  for (int s = 0; s < h->nS; ++s) {
    if (S != 0 && s != S - 1) continue;
    k_f2_from_host<<<dim3(nblk(h->Pp, 256), h->NPA * h->NE), 256, 0, st>>>(h->dev, h->d_stage, h->d_F2[h->sp[s].cur] + h->specStride * s, s);
    CKL();
    h->launches++;
  }
  CK(cudaStreamSynchronize(st));
  return RSG_OK;
}

int rsg_ram_f2_d2h(rsg_ram* h, double* F2, int S) {
  if (!h || !F2) return fail(RSG_ERR_ARG, "null argument");
  if (S < 0 || S > h->nS) return fail(RSG_ERR_ARG, "species index out of range");
  CK(cudaSetDevice(h->device));
  const size_t n = (size_t)h->nS * h->P * h->NE * h->NPA;
  if (S == 0) RET(rsg_ram_sync(h));
  cudaStream_t st = S ? h->st(S - 1) : h->pst();
  if (S != 0) {
    // keep the other species' host values: start from the host image
    CK(cudaMemcpyAsync(h->d_stage, F2, n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  for (int s = 0; s < h->nS; ++s) {
    if (S != 0 && s != S - 1) continue;
    k_f2_to_host<<<dim3(nblk(h->Pp, 256), h->NPA * h->NE), 256, 0, st>>>(h->dev, h->d_stage, h->d_F2[h->sp[s].cur] + h->specStride * s, s);
    CKL();
    h->launches++;
  }
  CK(cudaMemcpyAsync(F2, h->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return RSG_OK;
}

int rsg_ram_f2_device(rsg_ram* h, int S, void** ptr, long long* n_doubles, int* Pp) {
  RET(check_S(h, S));
  if (ptr) *ptr = h->d_F2[h->sp[S - 1].cur] + h->specStride * (S - 1);
  if (n_doubles) *n_doubles = (long long)h->specStride;
  if (Pp) *Pp = h->Pp;
  return RSG_OK;
}

// ---- ModRamDrift ----------------------------------------------------------------
int rsg_driftpara(rsg_ram* h, int S, double DTs) {
  RET(check_S(h, S));
  if (!h->grids_set) return fail(RSG_ERR_STATE, "DRIFTPARA before set_grids");
  CK(cudaSetDevice(h->device));
  const int s = S - 1;
  RET(ensure_step(h, DTs));
  // the previous contents of h_tab may still be in flight on the species stream
  if (h->sp[s].DTs != DTs) CK(cudaStreamSynchronize(h->st(s)));
  RET(tables_drift(h, s, DTs, h->st(s)));
  return L_inflow(h, s, 1, h->st(s));
}

#define SWEEP_ENTRY(name, fn)                                                      \
  int name(rsg_ram* h, int S) {                                                    \
    RET(check_S(h, S));                                                            \
    if (h->sp[S - 1].DTs < 0) return fail(RSG_ERR_STATE, #name " before DRIFTPARA"); \
    CK(cudaSetDevice(h->device));                                                  \
    return fn(h, S - 1, 1, h->st(S - 1));                                          \
  }
SWEEP_ENTRY(rsg_driftr, L_driftr)
SWEEP_ENTRY(rsg_driftp, L_driftp)
SWEEP_ENTRY(rsg_drifte, L_drifte)
SWEEP_ENTRY(rsg_driftmu, L_driftmu)

int rsg_driftend(rsg_ram* h) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  return RSG_OK;
}

int rsg_get_dtdrift(rsg_ram* h, int S, double out4[4]) {
  RET(check_S(h, S));
  if (!out4) return fail(RSG_ERR_ARG, "null out");
  CK(cudaSetDevice(h->device));
  RET(fetch_res(h, S - 1, h->st(S - 1)));
  std::memcpy(out4, h->sp[S - 1].h_res, 4 * sizeof(double));
  return RSG_OK;
}

// ---- ModRamLoss -------------------------------------------------------------------
int rsg_cepara(rsg_ram* h, int S, double DTs) {
  RET(check_S(h, S));
  if (!h->grids_set) return fail(RSG_ERR_STATE, "CEPARA before set_grids");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->st(S - 1)));
  return tables_cepara(h, S - 1, DTs, h->st(S - 1));
}

int rsg_charexchange(rsg_ram* h, int S) {
  RET(check_S(h, S));
  Spec& sp = h->sp[S - 1];
  if (sp.DTs_ce < 0) return fail(RSG_ERR_STATE, "CHAREXCHANGE before CEPARA");
  if (!h->fields_set) return fail(RSG_ERR_STATE, "CHAREXCHANGE before set_fields");
  CK(cudaSetDevice(h->device));
  if (h->kind[S - 1] == RSG_KIND_E) return RSG_OK;  // CHARGE == 1 for electrons
  return L_loss(h, S - 1, 0, sp.DTs_ce, h->st(S - 1));
}

int rsg_atmol(rsg_ram* h, int S) {
  RET(check_S(h, S));
  Spec& sp = h->sp[S - 1];
  if (sp.DTs_ce < 0) return fail(RSG_ERR_STATE, "ATMOL before CEPARA");
  if (!h->fields_set) return fail(RSG_ERR_STATE, "ATMOL before set_fields");
  CK(cudaSetDevice(h->device));
  return L_loss(h, S - 1, 1, sp.DTs_ce, h->st(S - 1));
}

// ---- ModRamWPI --------------------------------------------------------------------
int rsg_wavelo(rsg_ram* h, int S, double DTs) {
  RET(check_S(h, S));
  CK(cudaSetDevice(h->device));
  const int s = S - 1;
  if (h->sp[s].DTs_wl != DTs) CK(cudaStreamSynchronize(h->st(s)));
  RET(tables_wavelo(h, s, DTs, h->st(s)));
  return L_loss(h, s, 2, DTs, h->st(s));
}

int rsg_wpadif(rsg_ram* h, int S, double DTs, long long* nviolation) {
  RET(check_S(h, S));
  if (!h->fields_set) return fail(RSG_ERR_STATE, "WPADIF before set_fields");
  CK(cudaSetDevice(h->device));
  const int s = S - 1;
  RET(L_wpadif(h, s, DTs, h->st(s)));
  if (nviolation) {
    RET(fetch_res(h, s, h->st(s)));
    *nviolation = (long long)h->sp[s].h_res[4 + NSUM];
  }
  return RSG_OK;
}

// FLCscatter (src/ModRamLoss.f90:513-575).  The reference skips it during the first boundary
// cycle (TimeRamElapsed < Dt_bc): the caller passes both times.
int rsg_flcscatter(rsg_ram* h, int S, double DTs, double T, double Dt_bc, long long* nviolation) {
  RET(check_S(h, S));
  if (!h->fields_set) return fail(RSG_ERR_STATE, "FLCscatter before set_fields");
  if (nviolation) *nviolation = 0;
  if (T < Dt_bc) return RSG_OK;
  CK(cudaSetDevice(h->device));
  const int s = S - 1;
  RET(L_wpadif(h, s, DTs, h->st(s), 0, -1, true));
  if (nviolation) {
    RET(fetch_res(h, s, h->st(s)));
    *nviolation = (long long)h->sp[s].h_res[4 + NSUM];
  }
  return RSG_OK;
}

// ---- ModRamCoul -----------------------------------------------------------------------
namespace {
// src/ModRamFunctions.f90:72-143
double gcoul_h(double x) {
  const double g1 = std::erf(x) - 2. * x / std::sqrt(kPI) * std::exp(-x * x);
  return g1 / 2. / x / x;
}
double funt_h(double x) {
  const double y = std::sqrt(1 - x * x);
  const double alpha = 1. + std::log(2. + std::sqrt(3.)) / 2. / std::sqrt(3.);
  const double beta = alpha / 2. - kPI * std::sqrt(2.) / 12.;
  return alpha - beta * (y + std::sqrt(y)) + 0.055 * std::pow(y, 1. / 3.) + -0.037 * std::pow(y, 2. / 3.) + -0.074 * y +
         0.056 * std::pow(y, 4. / 3.);
}
double funi_h(double x) {
  const double y = std::sqrt(1 - x * x);
  const double ylog = (y > 0) ? std::log(y) : 0.0;
  const double alpha = 1. + std::log(2. + std::sqrt(3.)) / 2. / std::sqrt(3.);
  const double beta = alpha / 2. - kPI * std::sqrt(2.) / 12.;
  const double a1 = 0.055, a2 = -0.037, a3 = -0.074, a4 = 0.056;
  return 2. * alpha * (1. - y) + 2. * beta * y * ylog + 4. * beta * (y - std::sqrt(y)) + 3. * a1 * (std::pow(y, 1. / 3.) - y) +
         6. * a2 * (std::pow(y, 2. / 3.) - y) + 6. * a4 * (y - std::pow(y, 4. / 3.)) - 2. * a3 * y * ylog;
}
// COULPARA (src/ModRamCoul.f90:17-125): energy / pitch-angle tables of the Coulomb drag and
// scattering rates against the plasmasphere population (RAMSpecies(1:6): e-, H+, He+, O+, N+, Sr+
// with plasmasphereRatio 1, 0.77, 0.2, 0.03, 0, 0; src/ModRamSpecies.f90:42-133).  As in the
// reference the collision sums are NOT reset inside the energy loop (they accumulate over K).
// Tables are laid out [k][l]; entries the reference never assigns stay 0.
int tables_coulomb(rsg_ram* h, int s, double DTs) {
  Spec& sp = h->sp[s];
  if (sp.DTs_coul == DTs) return RSG_OK;
  const int nS = h->nS, NE = h->NE, NPA = h->NPA;
  const double MP = 1.673E-27, RE = 6.371E6, EPS = 8.854E-12, DLN = 21.5;
  static const double ps_mass[6] = {5.4462E-4, 1.0, 4.0, 16.0, 14.0, 87.62};
  static const double ps_charge[6] = {-1, 1, 1, 1, 1, 1};
  static const double ps_ratio[6] = {1.0, 0.77, 0.2, 0.03, 0.0, 0.0};
  std::vector<double> tab((size_t)4 * NE * NPA, 0.0), de(NPA, 0.0), di(NPA, 0.0);
  double* COULE = tab.data();
  double* COULI = COULE + (size_t)NE * NPA;
  double* ATA = COULI + (size_t)NE * NPA;
  double* GTA = ATA + (size_t)NE * NPA;
  const double Zt = (double)h->QS[s];
  const double QE = (kQ * kQ / EPS);
  const double GAMA = Zt * Zt * DLN / 4. / kPI * QE * 1E6 * QE;
  const double CCO = GAMA / kQ * DTs / kQ / 1E3;
  const double CCD = GAMA * DTs / (h->RMAS[s] * h->RMAS[s]) / (kCS * kCS * kCS);
  double CCE = 0, CDE = 0, CCI = 0, CDI = 0;
  const double *MU = h->MU.data(), *WMU = h->WMU.data(), *DMU = h->DMU.data();
  for (int k = 0; k < NE; ++k) {
    const double Vk = h->V[s + (size_t)nS * k], VBk = h->VBND[s + (size_t)nS * k];
    const double GRk = h->GREL[s + (size_t)nS * k], GRBk = h->GRBND[s + (size_t)nS * k];
    for (int b = 0; b < 6; ++b) {
      const double RA = ps_ratio[b];
      if (RA < 1e-9) continue;
      const double VF = std::sqrt(2. * kQ / (MP * ps_mass[b]));
      const double Zb = ps_charge[b];
      const double X = VBk / VF, XD = Vk / VF;
      if (Zb < 0.0) {
        CCE = CCE + RA * gcoul_h(X);
        CDE = CDE + RA * (std::erf(XD) - gcoul_h(XD));
      } else {
        CCI = CCI + RA * (Zb * Zb) * gcoul_h(X);
        CDI = CDI + RA * (Zb * Zb) * (std::erf(XD) - gcoul_h(XD));
      }
    }
    const double ce1 = -CCE * VBk * CCO * (GRBk * GRBk);
    const double ci1 = -CCI * VBk * CCO * (GRBk * GRBk);
    COULE[(size_t)k * NPA] = ce1;
    COULI[(size_t)k * NPA] = ci1;
    const double CCDE = CCD * CDE * GRk / std::pow(GRk * GRk - 1, 1.5);
    const double CCDI = CCD * CDI * GRk / std::pow(GRk * GRk - 1, 1.5);
    for (int l = 1; l <= NPA - 2; ++l) {      // L = 2..NPA-1
      COULE[(size_t)k * NPA + l] = ce1;
      COULI[(size_t)k * NPA + l] = ci1;
      const double MUBOUN = MU[l] + 0.5 * WMU[l];
      const double BADIF = (1. - MUBOUN * MUBOUN) / MUBOUN / 2.;
      de[l] = CCDE * BADIF;
      const double AFER = de[l] / MU[l] / DMU[l] / WMU[l];
      const double ASEC = de[l - 1] / MU[l] / DMU[l - 1] / WMU[l];
      di[l] = CCDI * BADIF;
      const double AFIR = di[l] / MU[l] / DMU[l] / WMU[l];
      const double ASIC = di[l - 1] / MU[l] / DMU[l - 1] / WMU[l];
      ATA[(size_t)k * NPA + l] = AFIR + AFER;
      GTA[(size_t)k * NPA + l] = ASIC + ASEC;
    }
  }
  (void)RE; (void)funt_h; (void)funi_h;
  CK(cudaStreamSynchronize(h->st(s)));
  RET(up(sp.d_coul, tab.data(), tab.size()));
  {
    std::vector<double> ck(NE);
    for (int k = 0; k < NE; ++k) ck[k] = COULE[(size_t)k * NPA] + COULI[(size_t)k * NPA];
    RET(up(sp.d_coul + (size_t)4 * NE * NPA, ck.data(), NE));
    sp.sd.cK = sp.d_coul + (size_t)4 * NE * NPA;
    const double EZERO = h->EKEV[0] - h->WE[0];
    const double GRZERO = 1. + EZERO * 1000. * kQ / h->RMAS[s] / kCS / kCS;
    const double GREL1 = h->GREL[s], GREL2 = h->GREL[s + (size_t)h->nS];
    sp.sd.cg1 = std::sqrt((GREL1 * GREL1 - 1) / (GREL2 * GREL2 - 1));
    sp.sd.cg0 = std::sqrt((GRZERO * GRZERO - 1) / (GREL1 * GREL1 - 1));
  }
  sp.DTs_coul = DTs;
  sp.ctab_DTs = -1.0;
  return RSG_OK;
}
int L_coulen(rsg_ram* h, int s, cudaStream_t st) {
  Spec& sp = h->sp[s];
  if (sp.DTs_coul < 0) return fail(RSG_ERR_STATE, "COULEN before COULPARA");
  SpecPack pk;
  make_pack(h, pk, s, 1);
  // ghost cells F(1), F(0) (:154-156, :182-183)
  const double EZERO = h->EKEV[0] - h->WE[0];
  const double GRZERO = 1. + EZERO * 1000. * kQ / h->RMAS[s] / kCS / kCS;
  const double GREL1 = h->GREL[s], GREL2 = h->GREL[s + (size_t)h->nS];
  const double g1 = std::sqrt((GREL1 * GREL1 - 1) / (GREL2 * GREL2 - 1));
  const double g0 = std::sqrt((GRZERO * GRZERO - 1) / (GREL1 * GREL1 - 1));
  const size_t n = (size_t)h->NE * h->NPA;
  k_coulen<<<dim3(nblk(h->P, 128), h->NPA - 1), 128, 0, st>>>(h->dev, pk.s[s], sp.d_coul, sp.d_coul + n, h->d_NECR, GREL1, GREL2,
                                                              GRZERO, g1, g0);
  CKL();
  h->launches++;
  return RSG_OK;
}
int L_coulmu(rsg_ram* h, int s, double T, cudaStream_t st) {
  Spec& sp = h->sp[s];
  if (sp.DTs_coul < 0) return fail(RSG_ERR_STATE, "COULMU before COULPARA");
  SpecPack pk;
  make_pack(h, pk, s, 1);
  const int TB = 64;
  const size_t smem = sizeof(double) * 2 * h->NPA * TB;
  static thread_local size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    CK(cudaFuncSetAttribute(k_coulmu, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const size_t n = (size_t)h->NE * h->NPA;
  k_coulmu<<<nblk((long long)h->NE * h->Pp, TB), TB, smem, st>>>(h->dev, pk.s[s], sp.d_coul + 2 * n, sp.d_coul + 3 * n, h->d_NECR, T);
  CKL();
  h->launches++;
  return RSG_OK;
}
}  // namespace

int rsg_coulpara(rsg_ram* h, int S, double DTs) {
  RET(check_S(h, S));
  if (!h->grids_set) return fail(RSG_ERR_STATE, "COULPARA before set_grids");
  CK(cudaSetDevice(h->device));
  return tables_coulomb(h, S - 1, DTs);
}
int rsg_coulen(rsg_ram* h, int S) {
  RET(check_S(h, S));
  if (!h->fields_set) return fail(RSG_ERR_STATE, "COULEN before set_fields");
  CK(cudaSetDevice(h->device));
  return L_coulen(h, S - 1, h->st(S - 1));
}
int rsg_coulmu(rsg_ram* h, int S, double T) {
  RET(check_S(h, S));
  if (!h->fields_set) return fail(RSG_ERR_STATE, "COULMU before set_fields");
  CK(cudaSetDevice(h->device));
  return L_coulmu(h, S - 1, T, h->st(S - 1));
}

// ---- ModRamRun --------------------------------------------------------------------
int rsg_sumrc(rsg_ram* h, int S, double* setrc, double* elorc) {
  RET(check_S(h, S));
  CK(cudaSetDevice(h->device));
  const int s = S - 1;
  Spec& sp = h->sp[s];
  RET(L_sumrc(h, s, 1, 0, h->st(s)));
  RET(fetch_res(h, s, h->st(s)));
  double v;
  std::memcpy(&v, sp.h_res + 4, 8);
  const double old = sp.setrc;
  sp.setrc = v;
  if (setrc) *setrc = v;
  if (elorc) *elorc = old - v;
  return RSG_OK;
}

// (NR,NT) planes of species s out of the pinned image; stride = 1 for a
// per-species slice, nS (offset s) for the full (nS,NR,NT) arrays
static void scatter_anisch(rsg_ram* h, int s, double* PPERT, double* PPART, int stride, int off) {
  const double* b = h->h_pp_all + (size_t)s * 2 * h->Pp;
  for (int p = 0; p < h->P; ++p) {
    if (PPERT) PPERT[(size_t)p * stride + off] = b[p];
    if (PPART) PPART[(size_t)p * stride + off] = b[(size_t)h->Pp + p];
  }
}

int rsg_anisch(rsg_ram* h, int S, double* PPERT_S, double* PPART_S) {
  RET(check_S(h, S));
  if (!h->fields_set) return fail(RSG_ERR_STATE, "ANISCH before set_fields");
  CK(cudaSetDevice(h->device));
  const int s = S - 1;
  cudaStream_t st = h->st(s);
  RET(L_anisch(h, s, 1, st));
  CK(cudaMemcpyAsync(h->h_pp_all + (size_t)s * 2 * h->Pp, h->sp[s].d_pp, (size_t)2 * h->Pp * sizeof(double),
                     cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  scatter_anisch(h, s, PPERT_S, PPART_S, 1, 0);
  return RSG_OK;
}

// ---- ram_run split at its two exchange points ------------------------------------------
// SUMRC slots (global numbering): 0 fwd drifts | 1 WPI diffusion | 2 EMIC | 3-6 fused loss
// block | 7 EMIC | 8 WPI diffusion | 9 reverse drifts.  cat: -1 unused, 0 LSDR 1 LSCHA 2 LSATM 3 LSWAE
namespace {
constexpr int NSLOT = 14;
// SUMRC slots in the order ram_run takes them (src/ModRamRun.f90:77-174); slots 10..13 are the
// Coulomb ones (COULEN, COULMU | COULMU, COULEN), appended so that the others keep their index
constexpr int kSlotOrder[NSLOT] = {0, 10, 11, 1, 2, 3, 4, 5, 6, 7, 8, 12, 13, 9};
void slot_cats(rsg_ram* h, int flags, int cat[RSG_MAX_SPECIES][NSLOT], int* doA, bool* wavelo_sp) {
  const bool DoUseWPI = flags & RSG_F_WPI, DoUseEMIC = flags & RSG_F_EMIC, DoUseCoulomb = flags & RSG_F_COULOMB;
  *doA = 0;
  for (int s = 0; s < h->nS; ++s) {
    const int kind = h->kind[s];
    const bool sWPI = (kind == RSG_KIND_E), sCEX = (kind != RSG_KIND_E), sEMIC = (kind == RSG_KIND_H);
    const bool wavelo = sWPI && !DoUseWPI;
    for (int q = 0; q < NSLOT; ++q) cat[s][q] = -1;
    cat[s][0] = 0; cat[s][9] = 0;
    if (sWPI && DoUseWPI) cat[s][1] = cat[s][8] = 3;
    if (sEMIC && DoUseEMIC) cat[s][2] = cat[s][7] = 3;
    if (sCEX) { cat[s][3] = cat[s][6] = 1; *doA |= 1 << s; }
    if (wavelo) { cat[s][3] = cat[s][6] = 3; *doA |= 1 << s; }
    cat[s][4] = cat[s][5] = 2;
    if (DoUseCoulomb) { cat[s][10] = cat[s][13] = 4; cat[s][11] = cat[s][12] = 5; }   // LSCOE, LSCSC
    if (wavelo_sp) wavelo_sp[s] = wavelo;
  }
}
int check_part(rsg_ram* h, int s0, int ns, int a0, int na, int amax) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (!h->grids_set || !h->fields_set || !h->efield_set) return fail(RSG_ERR_STATE, "ram_run before set_grids/fields/efield");
  if (s0 < 0 || ns < 1 || s0 + ns > h->nS) return fail(RSG_ERR_ARG, "species range out of bounds");
  if (a0 < 0 || na < 1 || a0 + na > amax) return fail(RSG_ERR_ARG, "slab range out of bounds");
  return RSG_OK;
}
}  // namespace

// part 1: tables, coefficient planes, DRIFTR inflow scan, forward DRIFTR/P/E on the
// pitch-angle slab [l0, l0+nl) of species [s0, s0+ns)        (src/ModRamRun.f90:67-75)
namespace {
// host part of a step: tables (pure functions of DTs), coefficient planes, DRIFTR inflow scan.
// Contains synchronisation, so it stays outside any stream capture.
int step_prepare(rsg_ram* h, double DTs, int flags, int s0, int ns) {
  RET(rsg_ram_sync(h));  // per-species streams idle; host staging tables free
  cudaStream_t st = h->pst();
  int cat[RSG_MAX_SPECIES][NSLOT], doA;
  bool wl[RSG_MAX_SPECIES];
  slot_cats(h, flags, cat, &doA, wl);
  h->prof_n = 0;
  RET(prof_mark(h, "prep_step", st));
  for (int s = s0; s < s0 + ns; ++s) {
    if (flags & RSG_F_COULOMB) RET(tables_coulomb(h, s, DTs));
    RET(tables_cepara(h, s, DTs, st));
    RET(tables_drift(h, s, DTs, st));
    if (wl[s]) RET(tables_wavelo(h, s, DTs, st));
  }
  RET(ensure_step(h, DTs, st));
  RET(prof_mark(h, "driftr_inflow", st));
  RET(L_inflow(h, s0, ns, st));
  if (fused_ok(h, flags)) {
    // measured on the B200 (profiles/r2): the fold trades k_anisch_pa_fast's read of F2 for 2 x rows x Pp doubles written by
    // the plane kernel and read back by k_finalize -- a wash at the default grid, a loss at the 4x grid (3 energies per
    // CTA => rows ~ 0.7 of F2).  Kept as an option.
    if (getenv("RSG_ANISCH_FOLD")) {
      const PlanePlan c = plane_plan(h);
      const size_t rows = (size_t)h->NPA * ((h->NE + c.cfg.KC - 1) / c.cfg.KC);
      for (int s = s0; s < s0 + ns; ++s)
        if (!h->sp[s].d_aE2) {
          RET(h->dalloc(&h->sp[s].d_aE2, 2 * rows * h->Pp));
          h->sp[s].sd.aE2 = h->sp[s].d_aE2;
          h->sp[s].sd.aA2 = h->sp[s].d_aE2 + rows * h->Pp;
          if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
        }
    }
    RET(L_cfl(h, s0, ns, st));
    int wm = wpadif_mask(h, flags);
    for (int s = 0; s < h->nS; ++s)
      if (s < s0 || s >= s0 + ns) wm &= ~(1 << s);
    if (wm) RET(L_wtab(h, wm, DTs, st));
    if (flags & RSG_F_COULOMB) RET(L_ctab(h, s0, ns, DTs, st));
  }
  return RSG_OK;
}
int enqueue_fwd(rsg_ram* h, int s0, int ns, int l0, int nl);
int enqueue_fused(rsg_ram* h, double DTs, int flags, int s0, int ns);
int enqueue_tail(rsg_ram* h, int s0, int ns, int l0, int nl, cudaStream_t st);
}  // namespace

int rsg_ram_part_fwd(rsg_ram* h, double DTs, int flags, int s0, int ns, int l0, int nl) {
  RET(check_part(h, s0, ns, l0, nl, h ? h->NPA : 0));
  CK(cudaSetDevice(h->device));
  RET(step_prepare(h, DTs, flags, s0, ns));
  return enqueue_fwd(h, s0, ns, l0, nl);
}

namespace {
int enqueue_fwd(rsg_ram* h, int s0, int ns, int l0, int nl) {
  cudaStream_t st = h->pst();
  // CFL minima, moment slots and counters of the owned species: one reset per step
  CK(cudaMemcpyAsync(h->d_res_all + (size_t)s0 * RES_N, h->d_res_init + (size_t)s0 * RES_N,
                     (size_t)ns * RES_N * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
  h->in_step = true;
  h->fwd_half = true;
  int rc = RSG_OK;
  do {
    if ((rc = prof_mark(h, "k_driftr", st)) != RSG_OK) break;
    if ((rc = L_driftr(h, s0, ns, st, l0, nl)) != RSG_OK) break;
    if ((rc = prof_mark(h, "k_driftp", st)) != RSG_OK) break;
    if ((rc = L_driftp(h, s0, ns, st, l0, nl)) != RSG_OK) break;
    if ((rc = prof_mark(h, "k_drifte", st)) != RSG_OK) break;
    if ((rc = L_drifte(h, s0, ns, st, l0, nl)) != RSG_OK) break;
    rc = prof_mark(h, "exchange", st);
  } while (0);
  h->in_step = false;
  return rc;
}
// epilogue of ram_run (src/ModRamRun.f90:186-209) and the pitch-angle sums of ANISCH
int enqueue_tail(rsg_ram* h, int s0, int ns, int l0, int nl, cudaStream_t st) {
  const PlaneRange pr{l0, nl, 0, h->NE};
  RET(prof_mark(h, "k_epilogue", st));
  {
    SpecPack pk;
    make_pack(h, pk, s0, ns);
    k_epilogue<<<dim3(nblk(h->NR + h->nout, 128), nl * h->NE, ns), 128, 0, st>>>(h->dev, pk, s0, h->d_outlist, h->nout, pr);
    CKL();
    h->launches++;
  }
  RET(prof_mark(h, "k_anisch", st));
  RET(L_anisch(h, s0, ns, st, l0, nl));
  RET(prof_mark(h, "d2h_results", st));
  return RSG_OK;
}
// the whole step of all species on the fused FAST kernels: F2 makes three round trips
int enqueue_fused(rsg_ram* h, double DTs, int flags, int s0, int ns) {
  cudaStream_t st = h->pst();
  int cat[RSG_MAX_SPECIES][NSLOT], doA;
  slot_cats(h, flags, cat, &doA, nullptr);
  h->in_step = false;
  RET(prof_mark(h, "k_plane_rp", st));
  RET(L_plane_rp(h, s0, ns, st, false));
  int doW = wpadif_mask(h, flags);
  for (int s = 0; s < h->nS; ++s)
    if (s < s0 || s >= s0 + ns) doW &= ~(1 << s);
  const int doC = (flags & RSG_F_COULOMB) ? 1 : 0;
  RET(prof_mark(h, "k_col_fused", st));
  RET(L_col(h, s0, ns, doA, DTs, st, 0, -1, doW, nullptr, doC));
  RET(prof_mark(h, "k_plane_rp", st));
  RET(L_plane_rp(h, s0, ns, st, true));          // ends with the epilogue of ram_run
  RET(L_finish_fused(h, s0, ns, st));
  if (doC) {                                     // SUMRC moments after COULEN, COULMU | COULMU, COULEN
    SpecPack pk;
    make_pack(h, pk, s0, ns);
    k_finalize_coul<<<dim3(4, ns), 256, 0, st>>>(pk, s0, (h->P + COL_PG - 1) / COL_PG, fused_cpart_off(h), RES_N, h->hd_res_all);
    CKL();
    h->launches++;
  }
  if (doW) {                                     // SUMRC moments after the two WPADIFs, violation count
    SpecPack pk;
    make_pack(h, pk, s0, ns);
    k_finalize_wpi<<<dim3(2, ns), 256, 0, st>>>(pk, s0, doW, (h->P + COL_PG - 1) / COL_PG, fused_wpart_off(h), RES_N, NSUM, h->d_wviol,
                                                h->hd_res_all);
    CKL();
    h->launches++;
  }
  return prof_mark(h, "d2h_results", st);
}
}  // namespace

// part 2: the pitch-angle block on the energy slab [k0, k0+nk): DRIFTMU, SUMRC, [WPADIF],
// fused losses, [WPADIF], DRIFTMU                                         (:76-170)
int rsg_ram_part_mid(rsg_ram* h, double DTs, int flags, int s0, int ns, int k0, int nk) {
  RET(check_part(h, s0, ns, k0, nk, h ? h->NE : 0));
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->pst();
  int cat[RSG_MAX_SPECIES][NSLOT], doA;
  slot_cats(h, flags, cat, &doA, nullptr);
  const PlaneRange pr{0, h->NPA, k0, nk};
  h->in_step = true;
  h->fwd_half = true;
  RET(prof_mark(h, "k_driftmu", st));
  {
    const int rc = L_driftmu(h, s0, ns, st, k0, nk, 0);   // SUMRC of :77 fused into the sweep
    h->in_step = false;
    if (rc != RSG_OK) return rc;
  }
  const bool coul = flags & RSG_F_COULOMB;
  if (coul && (k0 != 0 || nk != h->NE))
    return fail(RSG_ERR_UNSUPPORTED, "Coulomb operators need every energy and pitch angle local (no slab sharding)");
  RET(prof_mark(h, "wpadif+sumrc", st));
  if (coul)                                        // COULEN, SUMRC, COULMU, SUMRC  (:79-84)
    for (int s = s0; s < s0 + ns; ++s) {
      RET(L_coulen(h, s, st)); RET(L_sumrc(h, s, 1, 10, st, pr));
      RET(L_coulmu(h, s, h->T_elapsed, st)); RET(L_sumrc(h, s, 1, 11, st, pr));
    }
  for (int s = s0; s < s0 + ns; ++s)
    if (cat[s][1] >= 0) { RET(L_wpadif(h, s, DTs, st, k0, nk)); RET(L_sumrc(h, s, 1, 1, st, pr)); }
  for (int s = s0; s < s0 + ns; ++s)
    if (cat[s][2] >= 0) { RET(L_wpadif(h, s, DTs, st, k0, nk)); RET(L_sumrc(h, s, 1, 2, st, pr)); }
  RET(prof_mark(h, "k_loss_mid", st));
  RET(L_loss_mid(h, s0, ns, doA, DTs, 3, st, pr));
  RET(prof_mark(h, "wpadif+sumrc", st));
  for (int s = s0; s < s0 + ns; ++s)
    if (cat[s][7] >= 0) { RET(L_wpadif(h, s, DTs, st, k0, nk)); RET(L_sumrc(h, s, 1, 7, st, pr)); }
  for (int s = s0; s < s0 + ns; ++s)
    if (cat[s][8] >= 0) { RET(L_wpadif(h, s, DTs, st, k0, nk)); RET(L_sumrc(h, s, 1, 8, st, pr)); }
  if (coul)                                        // COULMU, SUMRC, COULEN, SUMRC  (:161-166)
    for (int s = s0; s < s0 + ns; ++s) {
      RET(L_coulmu(h, s, h->T_elapsed, st)); RET(L_sumrc(h, s, 1, 12, st, pr));
      RET(L_coulen(h, s, st)); RET(L_sumrc(h, s, 1, 13, st, pr));
    }
  RET(prof_mark(h, "k_driftmu", st));
  h->in_step = true;
  h->fwd_half = false;
  {
    const int rc = L_driftmu(h, s0, ns, st, k0, nk);
    h->in_step = false;
    if (rc != RSG_OK) return rc;
  }
  RET(prof_mark(h, "exchange", st));
  return RSG_OK;
}

// part 3: reverse DRIFTE/P/R, SUMRC, epilogue and the pitch-angle sums of ANISCH on the
// pitch-angle slab                                                      (:171-175, :186-209)
int rsg_ram_part_rev(rsg_ram* h, int s0, int ns, int l0, int nl) {
  RET(check_part(h, s0, ns, l0, nl, h ? h->NPA : 0));
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->pst();
  h->in_step = true;
  h->fwd_half = false;
  {
    int rc = RSG_OK;
    do {
      if ((rc = prof_mark(h, "k_drifte", st)) != RSG_OK) break;
      if ((rc = L_drifte(h, s0, ns, st, l0, nl)) != RSG_OK) break;
      if ((rc = prof_mark(h, "k_driftp", st)) != RSG_OK) break;
      if ((rc = L_driftp(h, s0, ns, st, l0, nl)) != RSG_OK) break;
      if ((rc = prof_mark(h, "k_driftr", st)) != RSG_OK) break;
      rc = L_driftr(h, s0, ns, st, l0, nl, 9);             // SUMRC of :174 fused into the sweep
    } while (0);
    h->in_step = false;
    if (rc != RSG_OK) return rc;
  }
  return enqueue_tail(h, s0, ns, l0, nl, st);
}

// ---- fused step for ranks that share a species (FAST mode, default operators) ---------------
// The plane kernels shard by pitch angle, the column kernel by blocks of 4 plane positions:
//   planes_fwd(l-slab) | exchange | columns(block range) | exchange | planes_rev(l-slab) + finish
// Moments and pressures are partial sums over the rank's slab / block range.
int rsg_ram_fused_available(rsg_ram* h, int flags) { return (h && flags == 0 && fused_ok(h, flags)) ? 1 : 0; }
int rsg_ram_fpart_planes_fwd(rsg_ram* h, double DTs, int flags, int s0, int ns, int l0, int nl) {
  RET(check_part(h, s0, ns, l0, nl, h ? h->NPA : 0));
  if (flags != 0 || !fused_ok(h, flags)) return fail(RSG_ERR_UNSUPPORTED, "fused kernels not available for this mode / flags / grid");
  CK(cudaSetDevice(h->device));
  RET(step_prepare(h, DTs, flags, s0, ns));
  h->in_step = false;
  return L_plane_rp(h, s0, ns, h->pst(), false, l0, nl);
}
int rsg_ram_fpart_columns(rsg_ram* h, double DTs, int flags, int s0, int ns, int b0, int nb) {
  const int nbtot = h ? (h->P + COL_PG - 1) / COL_PG : 0;
  RET(check_part(h, s0, ns, b0, nb, nbtot));
  if (flags != 0 || !fused_ok(h, flags)) return fail(RSG_ERR_UNSUPPORTED, "fused kernels not available for this mode / flags / grid");
  CK(cudaSetDevice(h->device));
  int cat[RSG_MAX_SPECIES][NSLOT], doA;
  slot_cats(h, flags, cat, &doA, nullptr);
  h->fp_nb = nb;
  return L_col(h, s0, ns, doA, DTs, h->pst(), b0, nb);
}
int rsg_ram_fpart_planes_rev(rsg_ram* h, int s0, int ns, int l0, int nl) {
  RET(check_part(h, s0, ns, l0, nl, h ? h->NPA : 0));
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->pst();
  RET(L_plane_rp(h, s0, ns, st, true, l0, nl));
  return L_finish_fused(h, s0, ns, st, l0, nl, h->fp_nb);
}
int rsg_ram_col_blocks(rsg_ram* h, int* nblocks, int* positions_per_block) {
  if (!h || !nblocks || !positions_per_block) return fail(RSG_ERR_ARG, "null argument");
  *nblocks = (h->P + COL_PG - 1) / COL_PG;
  *positions_per_block = COL_PG;
  return RSG_OK;
}

// raw per-rank results of the three parts: DtDrift(4,ns) minima, SUMRC partial sums
// moments(14,ns) over the local slab (unused slots are 0), partial PPERT/PPART(NR,NT,ns).
// Ranks sharing a species add moments and pressures and take the min of DtDrift.
namespace {
int enqueue_results(rsg_ram* h, int s0, int ns, bool pressures) {
  cudaStream_t st = h->pst();
  CK(cudaMemcpyAsync(h->h_res_all + (size_t)s0 * RES_N, h->d_res_all + (size_t)s0 * RES_N, (size_t)ns * RES_N * sizeof(unsigned long long),
                     cudaMemcpyDeviceToHost, st));
  if (pressures)
    CK(cudaMemcpyAsync(h->h_pp_all + (size_t)s0 * 2 * h->Pp, h->d_pp_all + (size_t)s0 * 2 * h->Pp,
                       (size_t)ns * 2 * h->Pp * sizeof(double), cudaMemcpyDeviceToHost, st));
  return RSG_OK;
}
int collect_results(rsg_ram* h, int s0, int ns, double* DtDrift, double* moments, double* PPER, double* PPAR);
}  // namespace

int rsg_ram_part_results(rsg_ram* h, int s0, int ns, double* DtDrift, double* moments, double* PPER, double* PPAR) {
  RET(check_part(h, s0, ns, 0, 1, 1));
  CK(cudaSetDevice(h->device));
  RET(enqueue_results(h, s0, ns, PPER || PPAR));
  return collect_results(h, s0, ns, DtDrift, moments, PPER, PPAR);
}

namespace {
int collect_results(rsg_ram* h, int s0, int ns, double* DtDrift, double* moments, double* PPER, double* PPAR) {
  cudaStream_t st = h->pst();
  RET(prof_mark(h, "end", st));
  CK(cudaStreamSynchronize(st));
  RET(prof_fold(h));
  for (int s = s0; s < s0 + ns; ++s) {
    Spec& sp = h->sp[s];
    if (DtDrift) std::memcpy(DtDrift + 4 * (s - s0), sp.h_res, 32);
    if (moments) std::memcpy(moments + NSLOT * (s - s0), sp.h_res + 4, NSLOT * 8);
    const double* b = h->h_pp_all + (size_t)s * 2 * h->Pp;
    if (PPER) std::memcpy(PPER + (size_t)h->P * (s - s0), b, h->P * sizeof(double));
    if (PPAR) std::memcpy(PPAR + (size_t)h->P * (s - s0), b + h->Pp, h->P * sizeof(double));
  }
  return RSG_OK;
}
}  // namespace

namespace {
// One full step of species [s0, s0+ns) on this GPU, results left in the pinned result block.
// The launch sequence of a step is fixed for a given (DTs, flags, mode, species range): kernel
// arguments carry DTs by value and the ping-pong buffers return to their start after 8 sweeps.
// It is captured once into a CUDA graph and replayed (no per-launch gaps); any change of the key
// re-captures.  Stage profiling needs the events between launches: no graph then.
int run_core(rsg_ram* h, double DTs, int flags, int s0, int ns) {
  RET(check_part(h, s0, ns, 0, h ? h->NPA : 0, h ? h->NPA : 0));
  CK(cudaSetDevice(h->device));
  RET(step_prepare(h, DTs, flags, s0, ns));
  cudaStream_t st = h->pst();
  const bool graph_ok = h->use_graph && !h->prof_on;
  if (graph_ok && h->gexec && h->g_DTs == DTs && h->g_flags == flags && h->g_mode == h->mode && h->g_s0 == s0 && h->g_ns == ns && h->g_tpos == (h->T_elapsed > 0.0) &&
      h->g_stream == st) {
    CK(cudaGraphLaunch(h->gexec, st));
    h->launches += h->g_launches;
    return RSG_OK;
  }
  if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
  const long long l0 = h->launches;
  if (graph_ok) CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc;
  if (fused_ok(h, flags)) rc = enqueue_fused(h, DTs, flags, s0, ns);   // results land in the pinned blocks directly
  else {
    rc = enqueue_fwd(h, s0, ns, 0, h->NPA);
    if (rc == RSG_OK) rc = rsg_ram_part_mid(h, DTs, flags, s0, ns, 0, h->NE);
    if (rc == RSG_OK) rc = rsg_ram_part_rev(h, s0, ns, 0, h->NPA);
    if (rc == RSG_OK) rc = enqueue_results(h, s0, ns, true);
  }
  if (graph_ok) {
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != RSG_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(RSG_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->gexec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { h->gexec = nullptr; return fail(RSG_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    h->g_DTs = DTs; h->g_flags = flags; h->g_mode = h->mode; h->g_s0 = s0; h->g_ns = ns; h->g_stream = st; h->g_tpos = (h->T_elapsed > 0.0);
    h->g_launches = h->launches - l0;
    CK(cudaGraphLaunch(h->gexec, st));
  }
  return rc;
}
}  // namespace

// Device result blocks (species-major): res = nS x res_n 8-byte words (CFL minima as ordered bit
// patterns, moments, counters), pp = nS x pp_n doubles (PPER plane, PPAR plane).  Species-sharded
// ranks all-gather these in place and then decode every species with rsg_ram_part_results.
int rsg_ram_results_device(rsg_ram* h, void** res, long long* res_n, void** pp, long long* pp_n) {
  if (!h || !res || !res_n || !pp || !pp_n) return fail(RSG_ERR_ARG, "null argument");
  *res = h->d_res_all; *res_n = RES_N;
  *pp = h->d_pp_all; *pp_n = 2 * (long long)h->Pp;
  return RSG_OK;
}

// All three parts of a step for species [s0, s0+ns) with every pitch angle and energy local
// (species-sharded ranks): same fused kernels and graph replay as rsg_ram_run.  Results through
// rsg_ram_part_results.
int rsg_ram_part_all(rsg_ram* h, double DTs, int flags, int s0, int ns) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  return run_core(h, DTs, flags, s0, ns);
}

// The whole species loop of ram_run (src/ModRamRun.f90:64-185) + epilogue (:186-222) on one
// GPU: the three parts back to back, all species advanced by each launch, on one stream.
namespace {
// decode the result blocks of all species (pinned copies) into the reference's step outputs: DtsNext (:202-205),
// DtDrift, the loss increments ELORC = ENOLD - SETRC chained through the SUMRC calls in ram_run's order (:77-174)
int decode_step(rsg_ram* h, int flags, double DtsMin, double* dts_next, double* DtDrift, double* losses, double* SETRC, double* PPERT,
                double* PPART) {
  const int nS = h->nS;
  std::vector<double> dt((size_t)4 * nS), mom((size_t)NSLOT * nS), pe((size_t)h->P * nS), pa((size_t)h->P * nS);
  RET(collect_results(h, 0, nS, dt.data(), mom.data(), pe.data(), pa.data()));
  int cat[RSG_MAX_SPECIES][NSLOT], doA;
  slot_cats(h, flags, cat, &doA, nullptr);
  double dtn = 1e300;
  for (int s = 0; s < nS; ++s) {
    Spec& sp = h->sp[s];
    for (int q = 0; q < 4; ++q) {
      dtn = std::min(dtn, dt[q + 4 * s]);
      if (DtDrift) DtDrift[q + 4 * s] = dt[q + 4 * s];
    }
    double ls[6] = {0, 0, 0, 0, 0, 0};
    double prev = sp.setrc;
    for (int qi = 0; qi < NSLOT; ++qi) {
      const int q = kSlotOrder[qi];
      if (cat[s][q] < 0) continue;
      const double v = mom[q + (size_t)NSLOT * s];
      ls[cat[s][q]] += prev - v;  // ELORC = ENOLD - SETRC (:256)
      prev = v;
    }
    sp.setrc = prev;
    if (losses)
      for (int q = 0; q < 6; ++q) losses[q + 6 * s] = ls[q];
    if (SETRC) SETRC[s] = prev;
    for (int p = 0; p < h->P; ++p) {
      if (PPERT) PPERT[(size_t)p * nS + s] = pe[(size_t)h->P * s + p];
      if (PPART) PPART[(size_t)p * nS + s] = pa[(size_t)h->P * s + p];
    }
  }
  if (dts_next) *dts_next = std::max(dtn, DtsMin);
  return RSG_OK;
}
}  // namespace

int rsg_ram_run(rsg_ram* h, double DTs, double DtsMin, double T, int flags, double* dts_next, double* DtDrift,
                double* losses, double* SETRC, double* PPERT, double* PPART) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  h->T_elapsed = T;                    // COULMU clamps negatives only once T > 0 (src/ModRamCoul.f90:289)
  RET(run_core(h, DTs, flags, 0, h->nS));
  return decode_step(h, flags, DtsMin, dts_next, DtDrift, losses, SETRC, PPERT, PPART);
}

// ---- the step with F2 coming from and returning to the HOST array (the routine-level drop-in's every step) ----------
// rsg_ram_f2_h2d + rsg_ram_run + rsg_ram_f2_d2h in one call, pipelined over chunks of pitch angles (L is the slowest index of
// the host array, so a chunk is one contiguous run): the upload of chunk c+1 runs beside the layout conversion and the
// forward plane kernel (DRIFTR, DRIFTP) of chunk c; after the column kernel the reverse plane kernel of a chunk is followed
// at once by its conversion and download, beside the next chunk's kernel.  Same kernels, same per-cell arithmetic as
// rsg_ram_run (F2 and the CFL limits bit-identical; moments summed in the same order).  What cannot overlap: upload and
// download themselves -- the column kernel needs every pitch angle of a position, the plane kernels every position of a
// pitch angle.  Steps the fused kernels do not cover (EXACT mode, RSG_NO_FUSE) run the three calls one after the other.
int rsg_ram_run_host(rsg_ram* h, double* F2, double DTs, double DtsMin, double T, int flags, double* dts_next, double* DtDrift,
                     double* losses, double* SETRC, double* PPERT, double* PPART) {
  if (!h || !F2) return fail(RSG_ERR_ARG, "null argument");
  const int NC = std::min(12, h->NPA / 4);
  bool same_buf = true;
  for (int s = 0; s < h->nS; ++s) same_buf = same_buf && h->sp[s].cur == 0;
  if (!fused_ok(h, flags) || !same_buf || h->ext || NC < 2 || h->sp[0].d_aE2 || getenv("RSG_NO_HOST_PIPELINE")) {
    RET(rsg_ram_f2_h2d(h, F2, 0));
    RET(rsg_ram_run(h, DTs, DtsMin, T, flags, dts_next, DtDrift, losses, SETRC, PPERT, PPART));
    return rsg_ram_f2_d2h(h, F2, 0);
  }
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  h->T_elapsed = T;
  const int nS = h->nS;
  RET(step_prepare(h, DTs, flags, 0, nS));
  if (!h->copyst) {
    CK(cudaStreamCreateWithFlags(&h->copyst, cudaStreamNonBlocking));
    for (auto& e : h->evh) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaStream_t st = h->pst(), cs = h->copyst;
  const size_t per_l = (size_t)nS * h->P * h->NE;
  SpecPtrs sp;
  for (int s = 0; s < RSG_MAX_SPECIES; ++s) sp.F[s] = s < nS ? h->d_F2[0] + h->specStride * s : nullptr;
  const dim3 tb(nblk(h->Pp, 256));
  int cat[RSG_MAX_SPECIES][NSLOT], doA;
  slot_cats(h, flags, cat, &doA, nullptr);
  const int doW = wpadif_mask(h, flags), doC = (flags & RSG_F_COULOMB) ? 1 : 0;
  h->in_step = false;
  CK(cudaEventRecord(h->evh[0], st));                     // the copies start after whatever the run stream still holds
  CK(cudaStreamWaitEvent(cs, h->evh[0], 0));
  for (int c = 0; c < NC; ++c) {                          // up: copy | convert + DRIFTR, DRIFTP
    int a, n;
    split_range(h->NPA, NC, c, &a, &n);
    CK(cudaMemcpyAsync(h->d_stage + a * per_l, F2 + a * per_l, n * per_l * sizeof(double), cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(h->evh[1 + c], cs));
    CK(cudaStreamWaitEvent(st, h->evh[1 + c], 0));
    k_f2_host_all<false><<<dim3(tb.x, n * h->NE), 256, 0, st>>>(h->dev, h->d_stage, sp, a * h->NE);
    CKL();
    RET(L_plane_rp(h, 0, nS, st, false, a, n));
    h->launches++;
  }
  RET(L_col(h, 0, nS, doA, DTs, st, 0, -1, doW, nullptr, doC));
  auto down = [&](int c, int a, int n) -> int {          // convert + copy of a finished chunk
    k_f2_host_all<true><<<dim3(tb.x, n * h->NE), 256, 0, st>>>(h->dev, h->d_stage, sp, a * h->NE);
    CKL();
    h->launches++;
    CK(cudaEventRecord(h->evh[16 + c], st));
    CK(cudaStreamWaitEvent(cs, h->evh[16 + c], 0));
    CK(cudaMemcpyAsync(F2 + a * per_l, h->d_stage + a * per_l, n * per_l * sizeof(double), cudaMemcpyDeviceToHost, cs));
    return RSG_OK;
  };
  for (int c = NC - 1; c >= 0; --c) {                     // down: DRIFTP, DRIFTR, epilogue | convert + copy
    int a, n;
    split_range(h->NPA, NC, c, &a, &n);
    RET(L_plane_rp(h, 0, nS, st, true, a, n, nullptr, 0));
    if (c > 0) RET(down(c, a, n));
  }
  RET(L_finish_fused(h, 0, nS, st));                     // ANISCH sums, second stage of the reductions; sets F2(L=1) = F2(L=2)
  {
    int a, n;
    split_range(h->NPA, NC, 0, &a, &n);
    RET(down(0, a, n));
  }
  if (doC) {
    SpecPack pk;
    make_pack(h, pk, 0, nS);
    k_finalize_coul<<<dim3(4, nS), 256, 0, st>>>(pk, 0, (h->P + COL_PG - 1) / COL_PG, fused_cpart_off(h), RES_N, h->hd_res_all);
    CKL();
    h->launches++;
  }
  if (doW) {
    SpecPack pk;
    make_pack(h, pk, 0, nS);
    k_finalize_wpi<<<dim3(2, nS), 256, 0, st>>>(pk, 0, doW, (h->P + COL_PG - 1) / COL_PG, fused_wpart_off(h), RES_N, NSUM, h->d_wviol,
                                                h->hd_res_all);
    CKL();
    h->launches++;
  }
  CK(cudaStreamSynchronize(st));
  CK(cudaStreamSynchronize(cs));
  return decode_step(h, flags, DtsMin, dts_next, DtDrift, losses, SETRC, PPERT, PPART);
}

int rsg_ram_flux_d2h(rsg_ram* h, double* FLUX) {
  if (!h || !FLUX) return fail(RSG_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  const size_t n = (size_t)h->nS * h->P * h->NE * h->NPA;
  cudaStream_t st = h->pst();
  SpecPack pk;
  make_pack(h, pk);
  for (int s = 0; s < h->nS; ++s) {
    k_flux_to_host<<<dim3(nblk(h->P, 256), h->NPA * h->NE), 256, 0, st>>>(h->dev, pk.s[s], h->d_stage);
    CKL();
    h->launches++;
  }
  CK(cudaMemcpyAsync(FLUX, h->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return RSG_OK;
}

long long rsg_ram_launch_count(rsg_ram* h) { return h ? h->launches : 0; }

int rsg_ram_profile(rsg_ram* h, int on) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  h->prof_on = on != 0;
  h->prof_acc.clear();
  h->prof_n = 0;
  return RSG_OK;
}
int rsg_ram_profile_get(rsg_ram* h, int idx, char* name, int name_len, double* ms_total, long long* count) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (idx < 0 || idx >= (int)h->prof_acc.size()) return RSG_ERR_ARG;
  const auto& a = h->prof_acc[idx];
  if (name && name_len > 0) {
    std::strncpy(name, a.first.c_str(), name_len - 1);
    name[name_len - 1] = 0;
  }
  if (ms_total) *ms_total = a.second.first;
  if (count) *count = a.second.second;
  return RSG_OK;
}

// device-side timing across the library's streams: begin() puts a start event in
// front of every species stream, end() joins them all and returns the elapsed ms
int rsg_ram_timer_begin(rsg_ram* h) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  CK(cudaSetDevice(h->device));
  if (!h->t0) {
    CK(cudaEventCreate(&h->t0));
    CK(cudaEventCreate(&h->t1));
  }
  cudaStream_t ps = h->pst();
  if (!h->ext)
    for (int s = 0; s < h->nS; ++s) {
      CK(cudaEventRecord(h->sp[s].ev, h->sp[s].own));
      CK(cudaStreamWaitEvent(ps, h->sp[s].ev, 0));
    }
  CK(cudaEventRecord(h->t0, ps));
  if (!h->ext)
    for (int s = 0; s < h->nS; ++s) CK(cudaStreamWaitEvent(h->sp[s].own, h->t0, 0));
  return RSG_OK;
}
int rsg_ram_timer_end(rsg_ram* h, double* ms) {
  if (!h || !ms) return fail(RSG_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  cudaStream_t ps = h->pst();
  if (!h->ext)
    for (int s = 0; s < h->nS; ++s) {
      CK(cudaEventRecord(h->sp[s].ev, h->sp[s].own));
      CK(cudaStreamWaitEvent(ps, h->sp[s].ev, 0));
    }
  CK(cudaEventRecord(h->t1, ps));
  CK(cudaEventSynchronize(h->t1));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, h->t0, h->t1));
  *ms = (double)f;
  return RSG_OK;
}

// pin / unpin an existing host array (e.g. the Fortran allocatable F2) so that
// the F2 transfers run at full PCIe speed and asynchronously
int rsg_host_register(void* p, long long bytes) {
  if (!p || bytes <= 0) return fail(RSG_ERR_ARG, "bad argument");
  CK(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
  return RSG_OK;
}
int rsg_host_unregister(void* p) {
  if (!p) return fail(RSG_ERR_ARG, "bad argument");
  CK(cudaHostUnregister(p));
  return RSG_OK;
}

}  // extern "C"

// ---- multi-GPU step over peer memory (device-side exchange and barriers) ----------------------
namespace {
// the decomposition of SURVEY 8(e) as a pure host function (include/ramscb_gpu.h: rsg_shard_plan_t)
void split_range(int n, int parts, int idx, int* start, int* count) {
  const int base = n / parts, rem = n % parts;
  *start = idx * base + std::min(idx, rem);
  *count = base + (idx < rem ? 1 : 0);
}

int plan_for(int world, int rank, int policy, int nS, int NPA, int P, rsg_shard_plan_t* p) {
  if (!p) return fail(RSG_ERR_ARG, "null plan");
  if (world < 1 || world > RSG_MAX_PEERS || rank < 0 || rank >= world) return fail(RSG_ERR_ARG, "world must be 1..8 and 0 <= rank < world");
  if (nS < 1 || nS > RSG_MAX_SPECIES) return fail(RSG_ERR_ARG, "bad species count");
  std::memset(p, 0, sizeof(*p));
  p->world = world; p->rank = rank; p->policy = policy;
  p->per = COL_PG;
  const int nblocks = (P + COL_PG - 1) / COL_PG;
  if (policy == RSG_SHARD_SLABS) {                 // every rank: a slab of ALL species
    p->s0 = 0; p->ns = nS; p->G = world; p->gidx = rank; p->g0 = 0;
  } else if (policy == RSG_SHARD_SPECIES) {        // species first, slabs inside a species beyond nS ranks
    if (world <= nS) {
      if (nS % world) return fail(RSG_ERR_ARG, "the species cannot be split evenly over the ranks");
      p->ns = nS / world; p->s0 = rank * p->ns; p->G = 1; p->gidx = 0; p->g0 = rank;
    } else {
      if (world % nS) return fail(RSG_ERR_ARG, "the rank count must be a multiple of the species count");
      p->G = world / nS; p->s0 = rank / p->G; p->ns = 1; p->gidx = rank % p->G; p->g0 = p->s0 * p->G;
    }
  } else return fail(RSG_ERR_ARG, "unknown sharding policy");
  if (p->G > 1 && (NPA / p->G < 2 || nblocks / p->G < 1)) return fail(RSG_ERR_ARG, "too many ranks per species for this grid");
  split_range(NPA, p->G, p->gidx, &p->l0, &p->nl);
  split_range(nblocks, p->G, p->gidx, &p->b0, &p->nb);
  return RSG_OK;
}

}  // namespace
extern "C" int rsg_shard_plan(int world, int rank, int policy, int nS, int NPA, int P, rsg_shard_plan_t* out) {
  return plan_for(world, rank, policy, nS, NPA, P, out);
}
#ifdef __CUDACC__
#include "ram_shard.inl"
#else
// host-CPU emulator build of the test-suite (tests/emu): no peer memory, no device-side barriers
namespace { void shard_release(rsg_ram*) {} }
extern "C" {
#define RSG_NO_PEER(name, ...) int name(__VA_ARGS__) { return fail(RSG_ERR_UNSUPPORTED, #name ": needs the sm_100a build"); }
RSG_NO_PEER(rsg_ram_shard_info, rsg_ram*, rsg_shard_plan_t*)
RSG_NO_PEER(rsg_ram_peer_export, rsg_ram*, void*)
RSG_NO_PEER(rsg_ram_peer_attach, rsg_ram*, int, int, int, const void*)
RSG_NO_PEER(rsg_ram_peer_attach_local, rsg_ram*, int, int, int, rsg_ram* const*)
RSG_NO_PEER(rsg_ram_peer_detach, rsg_ram*)
RSG_NO_PEER(rsg_ram_run_sharded, rsg_ram*, double, double, double, int, double*, double*, double*, double*, double*, double*)
RSG_NO_PEER(rsg_ram_run_sharded_enqueue, rsg_ram*, double, double, int)
RSG_NO_PEER(rsg_ram_run_sharded_collect, rsg_ram*, double, double*, double*, double*, double*, double*, double*)
RSG_NO_PEER(rsg_ram_f2_h2d_shard, rsg_ram*, const double*)
RSG_NO_PEER(rsg_ram_f2_d2h_shard, rsg_ram*, double*)
#undef RSG_NO_PEER
}
#endif
