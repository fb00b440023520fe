// Multi-GPU RAM step inside the library (included at the end of ram_gpu.cu).
//
// One process per GPU.  A rank owns species [s0, s0+ns) and, when G ranks share them, the pitch-angle slab
// [l0, l0+nl) for the plane kernels (DRIFTR/DRIFTP) and the plane-position blocks [b0, b0+nb) for the column
// kernel (DRIFTE, DRIFTMU, losses, WPADIF).  The step is the palindrome of src/ModRamRun.f90:70-175, so there are
// exactly two re-shardings -- and neither is a separate pass: the ranks map each other's F2 buffer (CUDA IPC peer
// memory over NVLink / NVSwitch) and the producing kernel's write-back stores every finished value straight into
// the buffer of the rank that reads it next (k_plane_rp<fwd, PEER>, k_col_fused<.., PEER>).  What remains between
// the kernels is a device-side barrier (k_peer_barrier: release / acquire flags at system scope in a peer-mapped
// mailbox), and after the step the result blocks (CFL limits, SUMRC partial sums, partial pressures) are pushed
// into every rank's mailbox and reduced in rank order, so every rank returns the same full result as rsg_ram_run
// on one GPU.  The whole sequence is ONE CUDA graph per (DTs, flags, mode): no host code, no NCCL call and no
// pack / unpack copy between the kernels.
//
// The host (Fortran + MPI, or Python + torch.distributed) only moves an opaque 192-byte blob per rank once
// (rsg_ram_peer_export -> all-gather -> rsg_ram_peer_attach).  rsg_ram_peer_attach_local wires several handles of
// ONE process together by pointer: the same kernels, barriers and graphs on one device (tests on a 1-GPU box).

namespace {

constexpr unsigned long long kBarrierTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;   // RSG_BARRIER_TIMEOUT_MS overrides (tests)
constexpr int kBlobMagic = 0x52534732;   // "RSG2"

struct PeerBlob {
  int magic, pid, device, nS, NR, NT, NE, NPA;
  cudaIpcMemHandle_t f2, mbox;
  char pad[RSG_PEER_BLOB_BYTES - 8 * sizeof(int) - 2 * sizeof(cudaIpcMemHandle_t)];
};
static_assert(sizeof(PeerBlob) == RSG_PEER_BLOB_BYTES, "blob size is part of the ABI");

// mailbox of one rank (device memory, mapped by every peer)
struct PeerCtx {
  int W, rank, nS, res_n;
  long long pp_n;
  unsigned char* mb[RSG_MAX_PEERS];
  size_t off_flags, off_epoch, off_res, off_pp;   // [4][MAX] u64 | [4] u64 | [MAX][nS][res_n] u64 | [MAX][nS][pp_n] f64
  unsigned long long* err;                        // host-mapped: non-zero = a barrier timed out
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Barrier among the ranks of `mask` (bit r), all of which launch it the same number of times.  Everything this
// GPU stored before the kernel (kernel boundary + fence) is visible to a peer once it has seen the flag.
// which: 0 = the species group's barrier, 1 = the world's, 2 = the group's barrier of the second species pipeline
// (RSG_PEER_PIPES=2); separate epoch counters.  One CTA of 32 threads.
__global__ void k_peer_barrier(const __grid_constant__ PeerCtx pc, int which, unsigned mask) {
  __shared__ unsigned long long e;
  if (threadIdx.x == 0) {
    unsigned long long* ep = (unsigned long long*)(pc.mb[pc.rank] + pc.off_epoch) + which;
    e = *ep + 1;
    *ep = e;
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r < pc.W && ((mask >> r) & 1u)) {
    __threadfence_system();
    st_release_sys((unsigned long long*)(pc.mb[r] + pc.off_flags) + which * RSG_MAX_PEERS + pc.rank, e);
    const unsigned long long* src = (const unsigned long long*)(pc.mb[pc.rank] + pc.off_flags) + which * RSG_MAX_PEERS + r;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(src) < e) {
      if (global_ns() - t0 > pc.timeout_ns) { *pc.err = 1ull + (unsigned long long)r; break; }
      __nanosleep(64);
    }
  }
}

// The two re-shardings of the step in bulk form (RSG_PEER_PUSH): the producing kernel writes into the LOCAL buffer and
// this kernel sends the rows to their next owner as contiguous runs of 16-byte stores -- full NVLink packets -- while
// the producer already works on its next chunk (second stream).  The source was written a moment ago: an L2 read.
//   mode 0 (after the forward plane kernel): rows [row0, row0+nrows) are this rank's (l, k) rows; blockIdx.z = the
//           destination rank, which receives its block range of plane positions [ccut[z], ccut[z+1]) of every row;
//   mode 1 (after the column kernel): all (l, k) rows whose pitch angle belongs to another rank; that rank receives the
//           columns [c0, c1) this launch of the column kernel has produced (3.9 KB per row at 8 ranks on the 4x grid,
//           where the column kernel's own write-back can only send the 32 bytes a CTA holds of a row).
// Warp per row; grid: x = row tiles, y = species, z = destinations (mode 0) / 1.
__global__ void __launch_bounds__(256) k_peer_push(const __grid_constant__ PeerView pv, const __grid_constant__ SpecPack pk, int s0,
                                                   int NE, int Pp, int P, int row0, int nrows, int c0, int c1, int mode) {
  const SpecDev& sp = pk.s[s0 + blockIdx.y];
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const int row = row0 + r;
  int o;
  if (mode == 0) {
    o = blockIdx.z;
    c0 = pv.ccut[o];
    c1 = min(pv.ccut[o + 1], P);
  } else {
    o = peer_owner(pv.lcut, pv.G, row / NE);
  }
  if (o == pv.gidx || c1 <= c0) return;
  const double* src = sp.F + (size_t)row * Pp;
  double* dst = pv.F[o] + (sp.F - pv.F[pv.gidx]) + (size_t)row * Pp;
  int c = c0;
  if ((c & 1) && c < c1) { if (lane == 0) dst[c] = src[c]; ++c; }                  // Pp is even: rows start 16-byte aligned
  const int n2 = (c1 - c) >> 1;
  for (int t = lane; t < n2; t += 32) *(double2*)(dst + c + 2 * t) = *(const double2*)(src + c + 2 * t);
  if (((c1 - c) & 1) && lane == 0) dst[c1 - 1] = src[c1 - 1];
}

// this rank's result blocks of species [s0, s0+ns) into slot `rank` of every rank's mailbox.  grid: x = destination rank
__global__ void __launch_bounds__(256) k_push_results(const __grid_constant__ PeerCtx pc, int s0, int ns,
                                                      const unsigned long long* __restrict__ res, const double* __restrict__ pp) {
  const int r = blockIdx.x;
  unsigned long long* dr = (unsigned long long*)(pc.mb[r] + pc.off_res) + ((size_t)pc.rank * pc.nS + s0) * pc.res_n;
  double* dp = (double*)(pc.mb[r] + pc.off_pp) + ((size_t)pc.rank * pc.nS + s0) * pc.pp_n;
  const unsigned long long* sr = res + (size_t)s0 * pc.res_n;
  const double* sp = pp + (size_t)s0 * pc.pp_n;
  for (int t = threadIdx.x; t < ns * pc.res_n; t += blockDim.x) dr[t] = sr[t];
  for (long long t = threadIdx.x; t < (long long)ns * pc.pp_n; t += blockDim.x) dp[t] = sp[t];
}

// Full results of every species on every rank: species s was advanced by ranks first[s] .. first[s]+cnt[s]-1, each
// holding the CFL limits (identical), partial SUMRC sums and partial pressures of its slab / column range.  Sums run
// in rank order (the same on every rank => identical bits everywhere).  grid: x = tiles, y = species
struct OwnerTab { int first[RSG_MAX_SPECIES], cnt[RSG_MAX_SPECIES]; };
__global__ void __launch_bounds__(256) k_reduce_results(const __grid_constant__ PeerCtx pc, const __grid_constant__ OwnerTab ow,
                                                        int nsum, int dtf_off, unsigned long long* __restrict__ res,
                                                        double* __restrict__ pp, unsigned long long* __restrict__ host_res,
                                                        double* __restrict__ host_pp) {
  const int s = blockIdx.y;
  const int f = ow.first[s], c = ow.cnt[s];
  const unsigned long long* mr = (const unsigned long long*)(pc.mb[pc.rank] + pc.off_res);
  const double* mp = (const double*)(pc.mb[pc.rank] + pc.off_pp);
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < pc.res_n) {
    const int w = (int)t;
    unsigned long long v = mr[((size_t)f * pc.nS + s) * pc.res_n + w];
    const bool is_min = (w < 4) || (w >= dtf_off && w < dtf_off + 4);
    const bool is_sum = (w >= 4 && w < 4 + nsum);
    if (is_min)
      for (int q = 1; q < c; ++q) {
        const unsigned long long o = mr[((size_t)(f + q) * pc.nS + s) * pc.res_n + w];
        v = o < v ? o : v;
      }
    else if (is_sum) {
      double a = __longlong_as_double((long long)v);
      for (int q = 1; q < c; ++q) a += __longlong_as_double((long long)mr[((size_t)(f + q) * pc.nS + s) * pc.res_n + w]);
      v = (unsigned long long)__double_as_longlong(a);
    }
    res[(size_t)s * pc.res_n + w] = v;
    host_res[(size_t)s * pc.res_n + w] = v;
  }
  if (t < pc.pp_n) {
    double a = mp[((size_t)f * pc.nS + s) * pc.pp_n + t];
    for (int q = 1; q < c; ++q) a += mp[((size_t)(f + q) * pc.nS + s) * pc.pp_n + t];
    pp[(size_t)s * pc.pp_n + t] = a;
    host_pp[(size_t)s * pc.pp_n + t] = a;
  }
}

}  // namespace

struct rsg_shard {
  rsg_shard_plan_t plan{};
  bool attached = false, local = false;
  unsigned char* mbox = nullptr;      // own mailbox (device)
  size_t mbox_bytes = 0;
  void* opened[2 * RSG_MAX_PEERS] = {nullptr};   // IPC mappings to close
  int nopened = 0;
  PeerCtx pc{};
  PeerView pv{};
  OwnerTab ow{};
  unsigned long long* h_err = nullptr;   // pinned + mapped
  cudaGraphExec_t gexec = nullptr;
  double g_DTs = -1.0;
  int g_flags = -1, g_mode = -1, g_tpos = -1;
  long long g_launches = 0;
  cudaStream_t g_stream = nullptr;
  unsigned group_mask = 0, world_mask = 0;
  int pending_flags = 0;
  bool pending = false;
  // RSG_PEER_PUSH: 0 = the kernels' own write-backs go to the peers; 1 = the column kernel writes locally, one bulk push
  // follows; 2 = both re-shardings in `nchunk` chunks, the push of chunk c on a second stream beside the kernel of chunk c+1
  int push_mode = 0, nchunk = 4;
  // RSG_PEER_PIPES=2: the rank's species run as two independent pipelines (planes | barrier | columns | barrier | planes) on
  // two streams, so one pipeline computes while the other waits for its peers' stores, and fills the other's last wave
  int pipes = 1;
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev[2 * 8 + 2] = {nullptr};
};

namespace {

int shard_alloc_mbox(rsg_ram* h) {
  if (h->shard && h->shard->mbox) return RSG_OK;
  if (!h->shard) h->shard = new rsg_shard();
  rsg_shard& sh = *h->shard;
  PeerCtx& pc = sh.pc;
  pc.nS = h->nS; pc.res_n = RES_N; pc.pp_n = 2 * (long long)h->Pp;
  pc.off_flags = 0;
  pc.off_epoch = pc.off_flags + sizeof(unsigned long long) * 4 * RSG_MAX_PEERS;
  pc.off_res = (pc.off_epoch + 4 * sizeof(unsigned long long) + 255) & ~(size_t)255;
  pc.off_pp = (pc.off_res + sizeof(unsigned long long) * RSG_MAX_PEERS * h->nS * RES_N + 255) & ~(size_t)255;
  sh.mbox_bytes = pc.off_pp + sizeof(double) * RSG_MAX_PEERS * h->nS * (size_t)pc.pp_n;
  RET(h->dalloc(&sh.mbox, sh.mbox_bytes));
  CK(cudaHostAlloc((void**)&sh.h_err, sizeof(unsigned long long), cudaHostAllocMapped));
  *sh.h_err = 0;
  CK(cudaHostGetDevicePointer((void**)&pc.err, sh.h_err, 0));
  pc.timeout_ns = kBarrierTimeoutNs;
  if (const char* e = getenv("RSG_BARRIER_TIMEOUT_MS")) pc.timeout_ns = (unsigned long long)std::max(1, std::atoi(e)) * 1000000ull;
  return RSG_OK;
}

int shard_finish_attach(rsg_ram* h, int rank, int world, int policy) {
  rsg_shard& sh = *h->shard;
  RET(plan_for(world, rank, policy, h->nS, h->NPA, h->P, &sh.plan));
  const rsg_shard_plan_t& p = sh.plan;
  sh.pc.W = world; sh.pc.rank = rank;
  sh.pv.G = p.G; sh.pv.gidx = p.gidx;
  for (int g = 0; g <= p.G; ++g) {
    int a, n;
    if (g < p.G) { split_range(h->NPA, p.G, g, &a, &n); sh.pv.lcut[g] = a; } else sh.pv.lcut[g] = h->NPA;
    if (g < p.G) { split_range((h->P + COL_PG - 1) / COL_PG, p.G, g, &a, &n); sh.pv.ccut[g] = a * COL_PG; } else sh.pv.ccut[g] = h->Pp;
  }
  sh.group_mask = 0; sh.world_mask = 0;
  for (int g = 0; g < p.G; ++g) sh.group_mask |= 1u << (p.g0 + g);
  for (int r = 0; r < world; ++r) sh.world_mask |= 1u << r;
  for (int s = 0; s < h->nS; ++s) {
    rsg_shard_plan_t q;
    int first = -1, cnt = 0;
    for (int r = 0; r < world; ++r) {
      RET(plan_for(world, r, policy, h->nS, h->NPA, h->P, &q));
      if (s >= q.s0 && s < q.s0 + q.ns) { if (first < 0) first = r; ++cnt; }
    }
    sh.ow.first[s] = first; sh.ow.cnt[s] = cnt;
  }
  for (int s = 0; s < h->nS; ++s)
    if (h->sp[s].cur != 0) return fail(RSG_ERR_STATE, "peer attach needs F2 in buffer 0 (attach before running single operators)");
  if (sh.gexec) { cudaGraphExecDestroy(sh.gexec); sh.gexec = nullptr; }
  if (const char* e = getenv("RSG_PEER_PUSH")) sh.push_mode = std::max(0, std::min(2, std::atoi(e)));
  if (const char* e = getenv("RSG_PEER_CHUNKS")) sh.nchunk = std::max(1, std::min(8, std::atoi(e)));
  // two species pipelines by default with one process per GPU (measured: 0.849 -> 0.820 ms at 8 GPUs, 1.30 -> 1.22 at 4, 2.10 -> 2.03
  // at 2); several in-process "ranks" on ONE device (attach_local, the test harness) stall each other's barriers with forked graphs
  sh.pipes = sh.local ? 1 : 2;
  if (const char* e = getenv("RSG_PEER_PIPES")) sh.pipes = std::max(1, std::min(2, std::atoi(e)));
  if ((sh.push_mode == 2 || sh.pipes == 2) && !sh.st2) {
    CK(cudaStreamCreateWithFlags(&sh.st2, cudaStreamNonBlocking));
    for (auto& e : sh.ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  sh.attached = true;
  return RSG_OK;
}

void shard_release(rsg_ram* h) {
  if (!h->shard) return;
  rsg_shard& sh = *h->shard;
  if (sh.gexec) cudaGraphExecDestroy(sh.gexec);
  if (sh.st2) cudaStreamDestroy(sh.st2);
  for (auto& e : sh.ev) if (e) cudaEventDestroy(e);
  for (int q = 0; q < sh.nopened; ++q) cudaIpcCloseMemHandle(sh.opened[q]);
  if (sh.h_err) cudaFreeHost(sh.h_err);
  delete h->shard;
  h->shard = nullptr;
}

// the launch sequence of one sharded step on the run stream (captured into the step's graph)
int enqueue_sharded(rsg_ram* h, double DTs, int flags) {
  rsg_shard& sh = *h->shard;
  const rsg_shard_plan_t& p = sh.plan;
  cudaStream_t st = h->pst();
  const int s0 = p.s0, ns = p.ns;
  if (p.G > 1) {
    int cat[RSG_MAX_SPECIES][NSLOT], doA;
    slot_cats(h, flags, cat, &doA, nullptr);
    int doW = wpadif_mask(h, flags);
    for (int s = 0; s < h->nS; ++s)
      if (s < s0 || s >= s0 + ns) doW &= ~(1 << s);
    h->in_step = false;
    const int doC = (flags & RSG_F_COULOMB) ? 1 : 0;
    const int doW_all = doW;
    auto pipeline = [&](const int s0, const int ns, cudaStream_t st, const int bar) -> int {
      int doW = doW_all;
      for (int s = 0; s < h->nS; ++s)
        if (s < s0 || s >= s0 + ns) doW &= ~(1 << s);
      SpecPack pk;
      make_pack(h, pk, s0, ns);
      auto push = [&](cudaStream_t q, int row0, int nrows, int c0, int c1, int mode) {
        k_peer_push<<<dim3(nblk(nrows, 8), ns, mode == 0 ? p.G : 1), 256, 0, q>>>(sh.pv, pk, s0, h->NE, h->Pp, h->P, row0, nrows, c0, c1, mode);
        h->launches++;
        return cudaGetLastError();
      };
      if (sh.push_mode == 2) {
        // both re-shardings chunked: kernel of chunk c on the run stream, its bulk push on the second stream beside chunk c+1
        const int NC = std::max(1, std::min({sh.nchunk, p.nl, p.nb}));
        RET(prof_mark(h, "k_plane_rp_fwd(local) || k_peer_push", st));
        for (int c = 0; c < NC; ++c) {
          int a, n;
          split_range(p.nl, NC, c, &a, &n);
          RET(L_plane_rp(h, s0, ns, st, false, p.l0 + a, n));
          CK(cudaEventRecord(sh.ev[c], st));
          CK(cudaStreamWaitEvent(sh.st2, sh.ev[c], 0));
          CK(push(sh.st2, (p.l0 + a) * h->NE, n * h->NE, 0, 0, 0));
        }
        CK(cudaEventRecord(sh.ev[16], sh.st2));
        CK(cudaStreamWaitEvent(st, sh.ev[16], 0));
        RET(prof_mark(h, "barrier_1", st));
        k_peer_barrier<<<1, 32, 0, st>>>(sh.pc, bar, sh.group_mask);
        CKL();
        RET(prof_mark(h, "k_col_fused(local) || k_peer_push", st));
        for (int c = 0; c < NC; ++c) {
          int a, n;
          split_range(p.nb, NC, c, &a, &n);
          RET(L_col(h, s0, ns, doA, DTs, st, p.b0 + a, n, doW, nullptr, doC, a));
          CK(cudaEventRecord(sh.ev[8 + c], st));
          CK(cudaStreamWaitEvent(sh.st2, sh.ev[8 + c], 0));
          CK(push(sh.st2, 0, h->NE * h->NPA, (p.b0 + a) * COL_PG, std::min(h->P, (p.b0 + a + n) * COL_PG), 1));
        }
        CK(cudaEventRecord(sh.ev[17], sh.st2));
        CK(cudaStreamWaitEvent(st, sh.ev[17], 0));
      } else {
        RET(prof_mark(h, "k_plane_rp_fwd(peer stores)", st));
        RET(L_plane_rp(h, s0, ns, st, false, p.l0, p.nl, &sh.pv));     // DRIFTR, DRIFTP -> the column owners
        RET(prof_mark(h, "barrier_1", st));
        k_peer_barrier<<<1, 32, 0, st>>>(sh.pc, bar, sh.group_mask);
        CKL();
        if (sh.push_mode == 1) {
          RET(prof_mark(h, "k_col_fused(local)", st));
          RET(L_col(h, s0, ns, doA, DTs, st, p.b0, p.nb, doW, nullptr, doC));  // DRIFTE .. DRIFTE into the local buffer
          RET(prof_mark(h, "k_peer_push", st));
          CK(push(st, 0, h->NE * h->NPA, p.b0 * COL_PG, std::min(h->P, (p.b0 + p.nb) * COL_PG), 1));
        } else {
          RET(prof_mark(h, "k_col_fused(peer stores)", st));
          RET(L_col(h, s0, ns, doA, DTs, st, p.b0, p.nb, doW, &sh.pv, doC));   // DRIFTE .. DRIFTE -> the pitch-angle owners
        }
      }
      RET(prof_mark(h, "barrier_2", st));
      k_peer_barrier<<<1, 32, 0, st>>>(sh.pc, bar, sh.group_mask);
      CKL();
      RET(prof_mark(h, "k_plane_rp_rev", st));
      RET(L_plane_rp(h, s0, ns, st, true, p.l0, p.nl));               // DRIFTP, DRIFTR, epilogue (local slab)
      RET(prof_mark(h, "anisch+finalize", st));
      RET(L_finish_fused(h, s0, ns, st, p.l0, p.nl, p.nb));
      h->launches += 2;
      if (doW) {
        SpecPack pk;
        make_pack(h, pk, s0, ns);
        k_finalize_wpi<<<dim3(2, ns), 256, 0, st>>>(pk, s0, doW, p.nb, fused_wpart_off(h), RES_N, NSUM, h->d_wviol, h->hd_res_all);
        CKL();
        h->launches++;
      }
      if (doC) {
        SpecPack pk;
        make_pack(h, pk, s0, ns);
        k_finalize_coul<<<dim3(4, ns), 256, 0, st>>>(pk, s0, p.nb, fused_cpart_off(h), RES_N, h->hd_res_all);
        CKL();
        h->launches++;
      }
      return RSG_OK;
    };
    if (sh.pipes == 2 && ns >= 2 && sh.push_mode != 2 && !h->prof_on) {   // (per-stage timing needs one stream)
      const int na = ns / 2;
      CK(cudaEventRecord(sh.ev[0], st));
      CK(cudaStreamWaitEvent(sh.st2, sh.ev[0], 0));
      RET(pipeline(s0, na, st, 0));
      RET(pipeline(s0 + na, ns - na, sh.st2, 2));
      CK(cudaEventRecord(sh.ev[1], sh.st2));
      CK(cudaStreamWaitEvent(st, sh.ev[1], 0));
    } else {
      RET(pipeline(s0, ns, st, 0));
    }
  } else if (fused_ok(h, flags)) {
    RET(enqueue_fused(h, DTs, flags, s0, ns));
  } else {
    RET(enqueue_fwd(h, s0, ns, 0, h->NPA));
    RET(rsg_ram_part_mid(h, DTs, flags, s0, ns, 0, h->NE));
    RET(rsg_ram_part_rev(h, s0, ns, 0, h->NPA));
  }
  RET(prof_mark(h, "results: push, world barrier, reduce", st));
  k_push_results<<<p.world, 256, 0, st>>>(sh.pc, s0, ns, h->d_res_all, h->d_pp_all);
  CKL();
  k_peer_barrier<<<1, 32, 0, st>>>(sh.pc, 1, sh.world_mask);
  CKL();
  k_reduce_results<<<dim3(nblk(std::max<long long>(sh.pc.pp_n, RES_N), 256), h->nS), 256, 0, st>>>(
      sh.pc, sh.ow, NSUM, DTF_OFF, h->d_res_all, h->d_pp_all, h->hd_res_all, h->hd_pp_all);
  CKL();
  RET(prof_mark(h, "end", st));
  h->launches += 3;
  return RSG_OK;
}

}  // namespace

extern "C" {

int rsg_ram_shard_info(rsg_ram* h, rsg_shard_plan_t* out) {
  if (!h || !out) return fail(RSG_ERR_ARG, "null argument");
  if (!h->shard || !h->shard->attached) return fail(RSG_ERR_STATE, "no peers attached");
  *out = h->shard->plan;
  return RSG_OK;
}

int rsg_ram_peer_export(rsg_ram* h, void* blob) {
  if (!h || !blob) return fail(RSG_ERR_ARG, "null argument");
  CK(cudaSetDevice(h->device));
  RET(shard_alloc_mbox(h));
  PeerBlob b;
  std::memset(&b, 0, sizeof(b));
  b.magic = kBlobMagic; b.pid = (int)getpid(); b.device = h->device;
  b.nS = h->nS; b.NR = h->NR; b.NT = h->NT; b.NE = h->NE; b.NPA = h->NPA;
  CK(cudaIpcGetMemHandle(&b.f2, h->d_F2[0]));
  CK(cudaIpcGetMemHandle(&b.mbox, h->shard->mbox));
  std::memcpy(blob, &b, sizeof(b));
  return RSG_OK;
}

int rsg_ram_peer_attach(rsg_ram* h, int rank, int world, int policy, const void* blobs) {
  if (!h || !blobs) return fail(RSG_ERR_ARG, "null argument");
  if (world < 1 || world > RSG_MAX_PEERS || rank < 0 || rank >= world) return fail(RSG_ERR_ARG, "world must be 1..8 and 0 <= rank < world");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  RET(shard_alloc_mbox(h));
  rsg_shard& sh = *h->shard;
  if (sh.attached) return fail(RSG_ERR_STATE, "peers already attached");
  rsg_shard_plan_t p;
  RET(plan_for(world, rank, policy, h->nS, h->NPA, h->P, &p));
  const PeerBlob* bl = (const PeerBlob*)blobs;
  for (int r = 0; r < world; ++r) {
    const PeerBlob& b = bl[r];
    if (b.magic != kBlobMagic || b.nS != h->nS || b.NR != h->NR || b.NT != h->NT || b.NE != h->NE || b.NPA != h->NPA)
      return fail(RSG_ERR_ARG, "peer blob of rank " + std::to_string(r) + " does not describe the same grid");
    const bool in_group = p.G > 1 && r >= p.g0 && r < p.g0 + p.G;
    if (r == rank) {
      sh.pc.mb[r] = sh.mbox;
      if (in_group) sh.pv.F[r - p.g0] = h->d_F2[0];
      continue;
    }
    if (b.pid == (int)getpid()) return fail(RSG_ERR_ARG, "peer in the same process: use rsg_ram_peer_attach_local");
    void* q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, b.mbox, cudaIpcMemLazyEnablePeerAccess));
    sh.opened[sh.nopened++] = q;
    sh.pc.mb[r] = (unsigned char*)q;
    if (in_group) {
      CK(cudaIpcOpenMemHandle(&q, b.f2, cudaIpcMemLazyEnablePeerAccess));
      sh.opened[sh.nopened++] = q;
      sh.pv.F[r - p.g0] = (double*)q;
    }
  }
  sh.local = false;
  return shard_finish_attach(h, rank, world, policy);
}

int rsg_ram_peer_attach_local(rsg_ram* h, int rank, int world, int policy, rsg_ram* const* peers) {
  if (!h || !peers) return fail(RSG_ERR_ARG, "null argument");
  if (world < 1 || world > RSG_MAX_PEERS || rank < 0 || rank >= world || peers[rank] != h) return fail(RSG_ERR_ARG, "bad rank / world / peer list");
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  RET(shard_alloc_mbox(h));
  rsg_shard& sh = *h->shard;
  if (sh.attached) return fail(RSG_ERR_STATE, "peers already attached");
  rsg_shard_plan_t p;
  RET(plan_for(world, rank, policy, h->nS, h->NPA, h->P, &p));
  for (int r = 0; r < world; ++r) {
    rsg_ram* q = peers[r];
    if (!q || q->nS != h->nS || q->NR != h->NR || q->NT != h->NT || q->NE != h->NE || q->NPA != h->NPA || q->device != h->device)
      return fail(RSG_ERR_ARG, "local peers must be handles of the same grid on the same device");
    RET(shard_alloc_mbox(q));
    sh.pc.mb[r] = q->shard->mbox;
    if (p.G > 1 && r >= p.g0 && r < p.g0 + p.G) sh.pv.F[r - p.g0] = q->d_F2[0];
  }
  sh.local = true;
  return shard_finish_attach(h, rank, world, policy);
}

int rsg_ram_peer_detach(rsg_ram* h) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (!h->shard) return RSG_OK;
  CK(cudaSetDevice(h->device));
  RET(rsg_ram_sync(h));
  rsg_shard& sh = *h->shard;
  if (sh.gexec) { cudaGraphExecDestroy(sh.gexec); sh.gexec = nullptr; }
  for (int q = 0; q < sh.nopened; ++q) cudaIpcCloseMemHandle(sh.opened[q]);
  sh.nopened = 0;
  sh.attached = false;
  return RSG_OK;
}

// enqueue: everything of the step is put on the run stream (graph replay when DTs / flags / mode repeat) and the
// call returns without waiting -- ranks living in one process (attach_local) enqueue all their steps first and
// collect afterwards; rsg_ram_run_sharded does both.
int rsg_ram_run_sharded_enqueue(rsg_ram* h, double DTs, double T, int flags) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (!h->shard || !h->shard->attached) return fail(RSG_ERR_STATE, "rsg_ram_run_sharded before rsg_ram_peer_attach");
  rsg_shard& sh = *h->shard;
  const rsg_shard_plan_t& p = sh.plan;
  RET(check_part(h, p.s0, p.ns, 0, h->NPA, h->NPA));
  if (p.G > 1 && !fused_ok(h, flags))
    return fail(RSG_ERR_UNSUPPORTED, "ranks sharing a species need the fused FAST kernels (RSG_MODE_FAST)");
  for (int s = p.s0; s < p.s0 + p.ns; ++s)
    if (p.G > 1 && h->sp[s].cur != 0) return fail(RSG_ERR_STATE, "sharded step needs F2 in buffer 0");
  CK(cudaSetDevice(h->device));
  h->T_elapsed = T;
  RET(step_prepare(h, DTs, flags, p.s0, p.ns));
  cudaStream_t st = h->pst();
  sh.pending = true;
  sh.pending_flags = flags;
  if (h->use_graph && !h->prof_on && sh.gexec && sh.g_DTs == DTs && sh.g_flags == flags && sh.g_mode == h->mode && sh.g_stream == st && sh.g_tpos == (T > 0.0)) {
    CK(cudaGraphLaunch(sh.gexec, st));
    h->launches += sh.g_launches;
    return RSG_OK;
  }
  if (sh.gexec) { cudaGraphExecDestroy(sh.gexec); sh.gexec = nullptr; }
  const bool graph_ok = h->use_graph && !h->prof_on;   // COULMU's T > 0 switch is a launch argument: part of the graph's key
  const long long l0 = h->launches;
  if (graph_ok) CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue_sharded(h, DTs, flags);
  if (graph_ok) {
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != RSG_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(RSG_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&sh.gexec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { sh.gexec = nullptr; return fail(RSG_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    sh.g_DTs = DTs; sh.g_flags = flags; sh.g_mode = h->mode; sh.g_stream = st; sh.g_tpos = T > 0.0;
    sh.g_launches = h->launches - l0;
    CK(cudaGraphLaunch(sh.gexec, st));
  }
  return rc;
}

int rsg_ram_run_sharded_collect(rsg_ram* h, double DtsMin, double* dts_next, double* DtDrift, double* losses, double* SETRC,
                                double* PPERT, double* PPART) {
  if (!h) return fail(RSG_ERR_ARG, "null handle");
  if (!h->shard || !h->shard->pending) return fail(RSG_ERR_STATE, "collect without a pending sharded step");
  rsg_shard& sh = *h->shard;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->pst()));
  sh.pending = false;
  RET(prof_fold(h));
  if (*sh.h_err) {
    const unsigned long long e = *sh.h_err;
    *sh.h_err = 0;
    return fail(RSG_ERR_CUDA, "peer barrier timed out waiting for rank " + std::to_string((long long)e - 1));
  }
  return decode_step(h, sh.pending_flags, DtsMin, dts_next, DtDrift, losses, SETRC, PPERT, PPART);
}

int rsg_ram_run_sharded(rsg_ram* h, double DTs, double DtsMin, double T, int flags, double* dts_next, double* DtDrift,
                        double* losses, double* SETRC, double* PPERT, double* PPART) {
  RET(rsg_ram_run_sharded_enqueue(h, DTs, T, flags));
  return rsg_ram_run_sharded_collect(h, DtsMin, dts_next, DtDrift, losses, SETRC, PPERT, PPART);
}

// F2 of the host array (full shape, species fastest) <-> this rank's share: its pitch-angle slab of its species.
// The slab is one contiguous run of the host array (L is the slowest index); with RSG_SHARD_SLABS it is all the
// rank moves.  d2h leaves the entries of species the rank does not own untouched.
int rsg_ram_f2_h2d_shard(rsg_ram* h, const double* F2) {
  if (!h || !F2) return fail(RSG_ERR_ARG, "null argument");
  if (!h->shard || !h->shard->attached) return fail(RSG_ERR_STATE, "no peers attached");
  const rsg_shard_plan_t& p = h->shard->plan;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->pst();
  const size_t per_l = (size_t)h->nS * h->P * h->NE;
  CK(cudaMemcpyAsync(h->d_stage + p.l0 * per_l, F2 + p.l0 * per_l, p.nl * per_l * sizeof(double), cudaMemcpyHostToDevice, st));
  for (int s = p.s0; s < p.s0 + p.ns; ++s) {
    k_f2_from_host<<<dim3(nblk(h->Pp, 256), p.nl * h->NE), 256, 0, st>>>(h->dev, h->d_stage, h->d_F2[h->sp[s].cur] + h->specStride * s, s,
                                                                          p.l0 * h->NE);
    CKL();
    h->launches++;
  }
  CK(cudaStreamSynchronize(st));
  return RSG_OK;
}

int rsg_ram_f2_d2h_shard(rsg_ram* h, double* F2) {
  if (!h || !F2) return fail(RSG_ERR_ARG, "null argument");
  if (!h->shard || !h->shard->attached) return fail(RSG_ERR_STATE, "no peers attached");
  const rsg_shard_plan_t& p = h->shard->plan;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = h->pst();
  const size_t per_l = (size_t)h->nS * h->P * h->NE;
  if (p.ns < h->nS)   // keep the other species' host values: start from the host image of the slab
    CK(cudaMemcpyAsync(h->d_stage + p.l0 * per_l, F2 + p.l0 * per_l, p.nl * per_l * sizeof(double), cudaMemcpyHostToDevice, st));
  for (int s = p.s0; s < p.s0 + p.ns; ++s) {
    k_f2_to_host<<<dim3(nblk(h->Pp, 256), p.nl * h->NE), 256, 0, st>>>(h->dev, h->d_stage, h->d_F2[h->sp[s].cur] + h->specStride * s, s,
                                                                        p.l0 * h->NE);
    CKL();
    h->launches++;
  }
  CK(cudaMemcpyAsync(F2 + p.l0 * per_l, h->d_stage + p.l0 * per_l, p.nl * per_l * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return RSG_OK;
}

}  // extern "C"
