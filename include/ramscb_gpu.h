/* ramscb_gpu.h -- C ABI of the B200-native RAM-SCB hot-path library
 * (libramscb_gpu.so, built from ramscb_b200/csrc/ for sm_100a).
 *
 * This is the drop-in boundary.  The reference (lanl/RAM-SCB) has no FFI for
 * these routines: they are Fortran module procedures that take only the
 * species index and read/write module globals.  Each entry point below
 * replaces the body of one such procedure; the Fortran shim modules in
 * ramscb_b200/fortran/ keep the reference's names/signatures and forward
 * c_loc() of the module arrays here (binding style copied from the
 * reference's only existing ISO_C_BINDING boundary, src/ModRamGSL.f90:10-71
 * <-> src/RamGSL.c).  INTEGRATION.md shows the shim a maintainer would add.
 *
 * Conventions
 *  - plain C, pointers + sizes only; every function returns 0 on success and
 *    a non-zero rsg_status otherwise (rsg_last_error() gives the text);
 *    nothing throws, nothing calls exit().
 *  - all host arrays are Fortran column-major with the reference's shapes
 *    (src/ModRamInit.f90:68-151); species index S is 1-based as in Fortran.
 *  - host pointers are never retained beyond the call.
 *  - calls with different S may be issued concurrently from different host
 *    threads (the reference calls them from `!$OMP PARALLEL DO` over species,
 *    src/ModRamRun.f90:64); calls with the same S are ordered by the caller.
 *  - there is no CPU fallback: without a CUDA device every call fails with
 *    RSG_ERR_CUDA.
 */
#ifndef RAMSCB_GPU_H
#define RAMSCB_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsg_ram rsg_ram; /* opaque RAM state (device mirrors + streams) */

enum rsg_status {
  RSG_OK = 0,
  RSG_ERR_CUDA = 1,     /* CUDA runtime error (no device, launch failure, ...) */
  RSG_ERR_ARG = 2,      /* bad argument (NULL, species out of range, ...)      */
  RSG_ERR_STATE = 3,    /* call order violated (e.g. DRIFTR before DRIFTPARA)  */
  RSG_ERR_UNSUPPORTED = 4
};

/* arithmetic mode of the sweep kernels */
enum rsg_mode {
  RSG_MODE_EXACT = 0, /* reference operation order, no FMA contraction: drift
                         sweeps / WPADIF are bit-identical to the CPU oracle   */
  RSG_MODE_FAST = 1   /* separable coefficients a+w(K)*b, FMA, division-free limiter: same maths, other
                         rounding.  Per cell within 1e-12 relative of the EXACT result after a sweep, and
                         after whole ram_run steps within 1e-12 for all but a handful of 1e-20 .. 1e-65
                         cells (<= 3e-12 measured; PARITY.md) -- the reference's own FMA-contracted
                         build moves by up to 5e-10                                               */
};

/* species kinds: select the charge-exchange cross-section polynomial
 * (src/ModRamLoss.f90:39-83) and the WPADIF coefficient pair (:677-685) */
enum rsg_kind { RSG_KIND_H = 0, RSG_KIND_O = 1, RSG_KIND_HE = 2, RSG_KIND_E = 3 };

/* flags of rsg_ram_run (src/ModRamParams.f90: DoUseWPI, DoUseCoulomb, DoUseEMIC) */
enum { RSG_F_WPI = 1, RSG_F_COULOMB = 2, RSG_F_EMIC = 4 };

const char* rsg_last_error(void);
int rsg_device_count(void);
/* name/SM count/L2 bytes of the current device; any pointer may be NULL */
int rsg_device_info(char* name, int name_len, int* sm_count, long long* l2_bytes, long long* hbm_bytes);

/* ---- life cycle ----------------------------------------------------------
 * replaces ram_allocate / ram_deallocate for the device mirrors
 * (src/ModRamInit.f90:13-154).  device < 0 keeps the current device. */
int rsg_ram_create(rsg_ram** out, int nS, int NR, int NT, int NE, int NPA, int device);
int rsg_ram_destroy(rsg_ram* h);
int rsg_ram_set_mode(rsg_ram* h, int mode);
/* run every species on `stream` (a cudaStream_t) instead of the library's own
 * per-species streams; NULL restores the default.  Lets a host framework time
 * the work with its own events. */
int rsg_ram_set_stream(rsg_ram* h, void* stream);
int rsg_ram_sync(rsg_ram* h);
/* rsg_ram_run replays its launch sequence from a CUDA graph when DTs/flags/mode repeat
 * (default on); 0 disables (every call launches kernel by kernel). */
int rsg_ram_use_graph(rsg_ram* h, int on);
/* FAST mode, flags without RSG_F_COULOMB: rsg_ram_run advances F2 with the fused shared-memory
 * kernels (DRIFTR+DRIFTP per plane; DRIFTE, DRIFTMU, [WPADIF], losses, [WPADIF], DRIFTMU, DRIFTE per
 * column block), default on.  0 selects the one-kernel-per-operator FAST path; with flags == 0 F2
 * is bit-identical either way.  With RSG_F_WPI / RSG_F_EMIC the column kernel applies WPADIF
 * (src/ModRamWPI.f90:643-714) from tabulated elimination factors (k_wpadif_tables; same matrix,
 * different rounding: <= 1e-12 of the one-kernel-per-operator result); on = 3 keeps WPADIF as its
 * own bit-exact kernel, which takes such a step to the one-kernel-per-operator path. */
int rsg_ram_use_fused(rsg_ram* h, int on);

/* ---- static data ----------------------------------------------------------
 * 1-D grids and per-species tables built by ARRAYS (src/ModRamInit.f90:364-587)
 * RLZ(NR+1) LZ(NR+1) EKEV,WE,DE,EBND(NE) MU,WMU,DMU(NPA) UPA(NR)
 * GREL,GRBND,V,VBND,EPP,ERNH(nS,NE) RMAS(nS) FFACTOR(nS,NR,NE,NPA)
 * QS(nS)=species%s_charge, kind(nS)=rsg_kind, khi(5)=ANISCH band edges
 * (src/ModRamRun.f90:303,322). */
int rsg_ram_set_grids(rsg_ram* h, const double* RLZ, const double* LZ, const double* EKEV, const double* WE,
                      const double* DE, const double* EBND, const double* MU, const double* WMU, const double* DMU,
                      const double* UPA, const double* GREL, const double* GRBND, const double* V,
                      const double* VBND, const double* EPP, const double* ERNH, const double* RMAS,
                      const double* FFACTOR, const int* QS, const int* kind, const int* khi, double MDR,
                      double DPHI, double CONF1, double CONF2, double BetaLim, double FracCFL);

/* field-geometry arrays written by computehI (src/ModRamScb.f90:568-626):
 * BNES,dBdt(NR+1,NT); FNHS,FNIS,BOUNHS,BOUNIS,HDNS,dIdt,dIbndt(NR+1,NT,NPA);
 * outsideMGNP(NR,NT) integer. */
int rsg_ram_set_fields(rsg_ram* h, const double* BNES, const double* dBdt, const double* FNHS, const double* FNIS,
                       const double* BOUNHS, const double* BOUNIS, const double* HDNS, const double* dIdt,
                       const double* dIbndt, const int* outsideMGNP);
/* VT,EIR,EIP(NR+1,NT): src/ModRamRun.f90:45-54, src/ModRamEField.f90 */
int rsg_ram_set_efield(rsg_ram* h, const double* VT, const double* EIR, const double* EIP);
/* FGEOS(nS,NT,NE,NPA): src/ModRamBoundary.f90 (GEOSB) */
int rsg_ram_set_boundary(rsg_ram* h, const double* FGEOS);
/* WALOS1/2/3(NR,NE) (WAVEPARA1/2, src/ModRamWPI.f90:18-182), Kp, Kpmax12 */
int rsg_ram_set_wavelo(rsg_ram* h, const double* WALOS1, const double* WALOS2, const double* WALOS3, double Kp,
                       double Kpmax12);
/* NECR(NR,NT) plasmaspheric density (Coulomb operators) */
int rsg_ram_set_plasmasphere(rsg_ram* h, const double* NECR);
/* pitch-angle diffusion coefficients (NR,NT,NE,NPA), src/ModRamRun.f90:422-605:
 * which = 0 ATAW, 1 ATAC, 2 ATAW_emic_h, 3 ATAW_emic_he */
int rsg_ram_set_diffcoef(rsg_ram* h, int which, const double* D);

/* SURVEY 8(f)-4, the producers of two of those inputs on the device.  GEOSB(S) (src/ModRamBoundary.f90:241-319, boundary
 * 'LANL'): FluxLanl(NT,NE) of get_geomlt_flux (file I/O: host) and species%s_comp -> FGEOS of species S, built where DRIFTR
 * reads it.  get_electric_field (src/ModRamEField.f90:14-63): vols = 0 interpolates VT between the potential maps VTOL, VTN
 * (NR+1,NT) at TimeRamElapsed; vols = 1 is the Volland-Stern potential (Kp, PHI(NT), PHIOFS; LZ from set_grids).  EIR / EIP keep
 * the values of the last rsg_ram_set_efield (zero before).  The readers of the flux / potential files and the restart /
 * NetCDF formats stay on the host. */
int rsg_geosb(rsg_ram* h, int S, const double* FluxLanl, double s_comp);
int rsg_ram_get_boundary(rsg_ram* h, int S, double* FGEOS_S /* (NT,NE,NPA) */);
int rsg_get_electric_field(rsg_ram* h, int vols, const double* VTOL, const double* VTN, double TimeRamElapsed, double TOLV,
                           double DtEfi, double Kp, const double* PHI, double PHIOFS, double* VT_out);

/* ANISCH, second half (src/ModRamRun.f90:422-605): the rebuild of those coefficients on the device, so a
 * WPI / EMIC run needs no host-built ATAW / ATAC / ATAW_emic_*.  set_wave_tables, once: the tabulated bounce-averaged
 * diffusion coefficients the reference reads at start-up (src/ModRamWPI.f90:185-470) -- ENOR(ENG), fpofc(NCF),
 * NDAAJ(NR,ENG,NPA,NCF) [hiss], CDAAR(NR,NT,NE,NPA) with use_bas = 1 (DoUseBASdiff) or BDAAR with 0 [chorus];
 * EKEV_emic(ENG_emic), fp2c_emic(NCF_emic), Daa_emic_h / _he(NR,ENG_emic,NPA,NCF_emic), Ihs_emic / Ihes_emic(4,NR,NT)
 * [EMIC]; PAbn(NPA).  Either group may be NULL.  rsg_anisch_diffcoef(S): chorus -> ATAC by the Steffen spline of
 * GSL_Interpolation_1D, hiss -> ATAW and EMIC -> ATAW_emic_h / _he by the bilinear rule of GSL_Interpolation_2D, for the
 * species S the flags select (RSG_F_WPI: electrons, RSG_F_EMIC: H+); XNE(NR,NT) plasmaspheric density, AE index.  The
 * caller keeps the reference's MOD(INT(T),INT(Dt_bc)) == 0 gate.  get_diffcoef: the arrays as the reference holds them. */
int rsg_ram_set_wave_tables(rsg_ram* h, int ENG, int NCF, const double* ENOR, const double* fpofc, const double* NDAAJ,
                            const double* DAAR, int use_bas, int ENG_emic, int NCF_emic, const double* EKEV_emic,
                            const double* fp2c_emic, const double* Daa_emic_h, const double* Daa_emic_he, const double* Ihs_emic,
                            const double* Ihes_emic, const double* PAbn);
int rsg_anisch_diffcoef(rsg_ram* h, int S, int flags, const double* XNE, int AE, int* gslerr);
int rsg_ram_get_diffcoef(rsg_ram* h, int which, double* D);

/* ---- phase-space density ---------------------------------------------------
 * F2(nS,NR,NT,NE,NPA), species fastest (src/ModRamInit.f90:72).  S = 0 moves
 * all species, S >= 1 one species (the host array is always the full one). */
int rsg_ram_f2_h2d(rsg_ram* h, const double* F2, int S);
int rsg_ram_f2_d2h(rsg_ram* h, double* F2, int S);
/* device pointer of species S's block, layout [NPA][NE][Pp] with the (NT,NR)
 * plane padded to Pp doubles (for NCCL plumbing and tests). */
int rsg_ram_f2_device(rsg_ram* h, int S, void** ptr, long long* n_doubles, int* Pp);

/* ---- operators: one per reference subroutine -------------------------------
 * ModRamDrift (src/ModRamDrift.f90) */
int rsg_driftpara(rsg_ram* h, int S, double DTs); /* :36-88   */
int rsg_driftr(rsg_ram* h, int S);                /* :95-198  */
int rsg_driftp(rsg_ram* h, int S);                /* :204-279 */
int rsg_drifte(rsg_ram* h, int S);                /* :285-376 */
int rsg_driftmu(rsg_ram* h, int S);               /* :382-473 */
int rsg_driftend(rsg_ram* h);                     /* :23-30 (no-op: scratch is persistent) */
/* DtDriftR/P/E/Mu(S) after the sweeps (synchronises species S) */
int rsg_get_dtdrift(rsg_ram* h, int S, double out4[4]);
/* ModRamLoss (src/ModRamLoss.f90) */
int rsg_cepara(rsg_ram* h, int S, double DTs); /* :19-170  */
int rsg_charexchange(rsg_ram* h, int S);       /* :457-478 */
int rsg_atmol(rsg_ram* h, int S);              /* :485-507 */
/* FLCscatter :513-575 (skipped while T < Dt_bc, :523).  FLC_coef(S,:,:,:,:) -- the output of
 * PARA_FLC, which stays on the host -- is set per species as a contiguous (NR,NT,NE,NPA) array. */
int rsg_ram_set_flc_coef(rsg_ram* h, int S, const double* FLC_coef);
int rsg_flcscatter(rsg_ram* h, int S, double DTs, double T, double Dt_bc, long long* nviolation);
/* PARA_FLC(S) :342-455 on the device: builds the species' FLC_coef from r_curvEq, zeta1Eq, zeta2Eq (NR,NT) --
 * the 2-D output of FLC_Radius (:176-340; host: SCB geometry + 2-D interpolation) -- with BNES, BOUNHS of
 * rsg_ram_set_fields; replaces rsg_ram_set_flc_coef's (NR,NT,NE,NPA) upload per species.  The caller keeps the
 * reference's "every Dt_bc" gate (:371).  rsg_ram_get_flc_coef: the array as the reference holds it (diagnostics). */
int rsg_para_flc(rsg_ram* h, int S, const double* r_curvEq, const double* zeta1Eq, const double* zeta2Eq);
int rsg_ram_get_flc_coef(rsg_ram* h, int S, double* FLC_coef);
/* ModRamWPI (src/ModRamWPI.f90) */
int rsg_wavelo(rsg_ram* h, int S, double DTs);                       /* :580-636 */
int rsg_wpadif(rsg_ram* h, int S, double DTs, long long* nviolation); /* :643-714; nviolation may be NULL */
/* ModRamCoul (src/ModRamCoul.f90): COULPARA builds the rate tables from the grids of
 * rsg_ram_set_grids (host side, cached by DTs); COULEN / COULMU need rsg_ram_set_plasmasphere
 * (NECR) and the fields; T = TimeRamElapsed arms COULMU's negative clamp (:289). */
int rsg_coulpara(rsg_ram* h, int S, double DTs); /* :17-125  */
int rsg_coulen(rsg_ram* h, int S);               /* :133-221 */
int rsg_coulmu(rsg_ram* h, int S, double T);     /* :229-296 */
/* ModRamRun (src/ModRamRun.f90) */
int rsg_sumrc(rsg_ram* h, int S, double* setrc, double* elorc);          /* :231-259 */
int rsg_anisch(rsg_ram* h, int S, double* PPERT_S, double* PPART_S);     /* :343-415; (NR,NT) slices of species S */
/* The whole species loop of ram_run plus its epilogue (:64-222) in one call,
 * F2 resident on the device.  Outputs (any may be NULL): DtDrift(4,nS)
 * [R,P,E,Mu fastest], losses(6,nS) = increments of LSDR,LSCHA,LSATM,LSWAE,
 * LSCOE,LSCSC; SETRC(nS); PPERT,PPART(nS,NR,NT).  Returns DtsNext in *dts_next. */
int rsg_ram_run(rsg_ram* h, double DTs, double DtsMin, double T, int flags, double* dts_next, double* DtDrift,
                double* losses, double* SETRC, double* PPERT, double* PPART);
/* The same step (src/ModRamRun.f90:64-222) with F2 coming from and returning to the HOST array -- what the routine-level drop-in does every step
 * (rsg_ram_f2_h2d, rsg_ram_run, rsg_ram_f2_d2h) in one call, pipelined over chunks of pitch angles: the upload of a chunk
 * runs beside the layout conversion and DRIFTR / DRIFTP of the previous one, and after the column kernel each chunk is
 * downloaded as soon as its reverse DRIFTP / DRIFTR are done.  F2 (nS,NR,NT,NE,NPA) in / out, best page-locked
 * (rsg_host_register).  Results identical to the three calls. */
int rsg_ram_run_host(rsg_ram* h, double* F2, double DTs, double DtsMin, double T, int flags, double* dts_next, double* DtDrift,
                     double* losses, double* SETRC, double* PPERT, double* PPART);
/* Multi-GPU: rsg_ram_run split at its two exchange points.  A rank owns species
 * [s0, s0+ns) (0-based) and, inside them, the pitch-angle slab [l0, l0+nl) for the
 * R/P/E sweeps and the energy slab [k0, k0+nk) for the pitch-angle block; between the
 * parts the ranks sharing a species swap the complementary (L,K) blocks of the current
 * F2 buffer (rsg_ram_f2_device) with NCCL send/recv -- there is no other data-path
 * collective.  The calls only enqueue work (on the stream of rsg_ram_set_stream, or the
 * library's run stream); rsg_ram_part_results synchronises and returns the rank's raw
 * results: DtDrift(4,ns) [min over ranks], SUMRC partial sums moments(14,ns) and partial
 * pressures PPER/PPAR(NR,NT,ns) [sum over the ranks of a species].  One GPU:
 * rsg_ram_run == fwd(all) + mid(all) + rev(all) + results. */
int rsg_ram_part_fwd(rsg_ram* h, double DTs, int flags, int s0, int ns, int l0, int nl);
int rsg_ram_part_mid(rsg_ram* h, double DTs, int flags, int s0, int ns, int k0, int nk);
int rsg_ram_part_rev(rsg_ram* h, int s0, int ns, int l0, int nl);
/* All three parts at once for species [s0, s0+ns) when every pitch angle and energy is local
 * (species-sharded ranks, or one GPU): uses the fused kernels / graph replay of rsg_ram_run. */
int rsg_ram_part_all(rsg_ram* h, double DTs, int flags, int s0, int ns);
/* The fused FAST step for ranks that SHARE a species (flags == 0): the plane kernels shard by pitch
 * angle [l0, l0+nl), the column kernel by blocks [b0, b0+nb) of *positions_per_block consecutive
 * plane positions (rsg_ram_col_blocks).  Between the calls the caller re-shards F2 (rows l x
 * columns p of the [NPA*NE][Pp] species buffer).  planes_rev also runs the epilogue, ANISCH and
 * the result block; moments and pressures are partial sums over the slab / block range. */
int rsg_ram_fused_available(rsg_ram* h, int flags);
int rsg_ram_col_blocks(rsg_ram* h, int* nblocks, int* positions_per_block);
int rsg_ram_fpart_planes_fwd(rsg_ram* h, double DTs, int flags, int s0, int ns, int l0, int nl);
int rsg_ram_fpart_columns(rsg_ram* h, double DTs, int flags, int s0, int ns, int b0, int nb);
int rsg_ram_fpart_planes_rev(rsg_ram* h, int s0, int ns, int l0, int nl);
/* Device result blocks, species-major: res = nS x *res_n 8-byte words, pp = nS x *pp_n doubles.
 * Species-sharded ranks all-gather them in place (NCCL) and decode with rsg_ram_part_results(0, nS). */
int rsg_ram_results_device(rsg_ram* h, void** res, long long* res_n, void** pp, long long* pp_n);
int rsg_ram_part_results(rsg_ram* h, int s0, int ns, double* DtDrift, double* moments, double* PPER, double* PPAR);

/* ---- Multi-GPU inside the library: the sharded ram_run over NVLink peer memory ------------------
 * One process per GPU (<= 8 on one node).  The reference has no domain decomposition (its MPI is
 * init / finalize only, src/Main.f90:54-57); this replaces the species loop of ram_run
 * (src/ModRamRun.f90:64-185) the way SURVEY 8(e) shards it: species over ranks, and pitch-angle slabs /
 * plane-position blocks inside a species when several ranks share it, with exactly two re-shardings per
 * step.  Nothing of that is visible to the host: the ranks map each other's F2 buffer (CUDA IPC) and
 * the kernels store their results straight into the buffer of the rank that needs them next; barriers
 * and the reduction of the step's results run on the device too, in one CUDA graph per rank.
 *
 * Set-up, once: every rank calls rsg_ram_peer_export (an opaque blob of RSG_PEER_BLOB_BYTES), the host
 * all-gathers the blobs in rank order (MPI_Allgather in the Fortran host, torch.distributed in
 * ramscb_b200/parallel.py) and every rank calls rsg_ram_peer_attach.  Then rsg_ram_run_sharded is
 * rsg_ram_run: same arguments, same outputs on every rank (DtDrift, losses, SETRC, PPERT, PPART of ALL
 * species), F2 sharded.  It must be called by all ranks with the same DTs / flags (it contains barriers;
 * a rank that never arrives makes the others fail with a time-out after 20 s instead of hanging).
 * policy: RSG_SHARD_SPECIES = whole species per rank while world <= nS, G = world/nS ranks per species
 * beyond; RSG_SHARD_SLABS = every rank holds a pitch-angle slab of ALL species (a contiguous run of the
 * host array F2(nS,NR,NT,NE,NPA), which is what rsg_ram_f2_h2d_shard / d2h_shard move).
 * Ranks that share a species need RSG_MODE_FAST and no Coulomb flag (the fused kernels). */
#define RSG_PEER_BLOB_BYTES 192
enum { RSG_SHARD_SPECIES = 0, RSG_SHARD_SLABS = 1 };
typedef struct rsg_shard_plan_t {
  int world, rank, policy;
  int s0, ns;        /* owned species, 0-based [s0, s0+ns) */
  int G, gidx, g0;   /* ranks g0 .. g0+G-1 share them; this rank's index among them */
  int l0, nl;        /* pitch-angle slab of the plane kernels (DRIFTR, DRIFTP) */
  int b0, nb, per;   /* blocks of `per` plane positions of the column kernel (DRIFTE, DRIFTMU, losses, WPADIF) */
} rsg_shard_plan_t;
/* the decomposition as a pure function (no device needed): P = NR*NT plane positions */
int rsg_shard_plan(int world, int rank, int policy, int nS, int NPA, int P, rsg_shard_plan_t* out);
int rsg_ram_peer_export(rsg_ram* h, void* blob);
int rsg_ram_peer_attach(rsg_ram* h, int rank, int world, int policy, const void* blobs /* world x RSG_PEER_BLOB_BYTES */);
/* several handles of ONE process on one device wired together by pointer (tests on a 1-GPU box; a host
 * that drives several "ranks" itself): enqueue every rank's step, then collect every rank's results. */
int rsg_ram_peer_attach_local(rsg_ram* h, int rank, int world, int policy, rsg_ram* const* peers);
int rsg_ram_peer_detach(rsg_ram* h);
int rsg_ram_shard_info(rsg_ram* h, rsg_shard_plan_t* out);
int rsg_ram_run_sharded(rsg_ram* h, double DTs, double DtsMin, double T, int flags, double* dts_next, double* DtDrift,
                        double* losses, double* SETRC, double* PPERT, double* PPART);
int rsg_ram_run_sharded_enqueue(rsg_ram* h, double DTs, double T, int flags);
int rsg_ram_run_sharded_collect(rsg_ram* h, double DtsMin, double* dts_next, double* DtDrift, double* losses, double* SETRC,
                                double* PPERT, double* PPART);
/* this rank's share of the host array F2 (full shape): its pitch-angle slab of its species */
int rsg_ram_f2_h2d_shard(rsg_ram* h, const double* F2);
int rsg_ram_f2_d2h_shard(rsg_ram* h, double* F2);

/* FLUX = F2/FFACTOR/FNHS (src/ModRamRun.f90:210-221), host array like F2 */
int rsg_ram_flux_d2h(rsg_ram* h, double* FLUX);

/* number of kernel launches issued through this handle since creation */
long long rsg_ram_launch_count(rsg_ram* h);

/* Device-side timing across the library's internal streams (CUDA events):
 * begin() orders a start event before all later work of this handle, end()
 * joins every stream, synchronises and returns the elapsed milliseconds. */
int rsg_ram_timer_begin(rsg_ram* h);
int rsg_ram_timer_end(rsg_ram* h, double* ms);

/* Per-stage device timing of rsg_ram_run: when enabled, CUDA events are recorded
 * on the run stream between the stages (one stage = the launch(es) of one
 * kernel for all species) and accumulated; _get() enumerates the accumulators
 * (returns non-zero past the end).  ms_total/count: summed milliseconds and
 * number of intervals since the last rsg_ram_profile() call. */
int rsg_ram_profile(rsg_ram* h, int on);
int rsg_ram_profile_get(rsg_ram* h, int idx, char* name, int name_len, double* ms_total, long long* count);

/* Pin / unpin an existing host array (cudaHostRegister), e.g. the Fortran
 * allocatable F2, so rsg_ram_f2_h2d/d2h run at full PCIe speed. */
int rsg_host_register(void* p, long long bytes);
int rsg_host_unregister(void* p);

/* ===========================================================================
 * SCB: 3-D force-balance Euler-potential solve (second hot path)
 *
 * Replaces the bodies of the argument-less module procedures of ModScbCompute,
 * ModScbEquation and ModScbEuler, which work on the ModScbVariables globals
 * (src/ModScbVariables.f90:10-96, allocated src/ModScbInit.f90:22-119).  All 3-D
 * host arrays are Fortran (nthe,npsi,nzeta+1) ["+1" arrays: x,y,z,alfa,psi,bf,
 * bsq,pper,ppar,sigma] or (nthe,npsi,nzeta) [everything else].  The device keeps
 * a mirror of every array; rsg_scb_get_field/set_field move any of them by its
 * reference name ("jacobian", "vecd", "Bx", "GradRhoSq", ...), so the Fortran
 * host only downloads what it reads.
 * ========================================================================= */
typedef struct rsg_scb rsg_scb;

/* SOR sweep ordering */
enum rsg_sor_order {
  RSG_SOR_LEX = 0,    /* the reference's lexicographic Gauss-Seidel order, run as a skewed
                         wavefront: iterates, iteration counts and residuals bit-identical */
  RSG_SOR_COLOR4 = 1  /* 4-colour ordering (the 9-point stencil has corner couplings, so
                         red-black is not enough): same fixed point, far more parallel   */
};

const char* rsg_scb_last_error(void);
int rsg_scb_create(rsg_scb** out, int nthe, int npsi, int nzeta, int device); /* scb_allocate, src/ModScbInit.f90:13-155 */
int rsg_scb_destroy(rsg_scb* h);
/* thetaVal(nthe) rhoVal(npsi) zetaVal(nzeta) (src/ModScbInit.f90:278-290), f(npsi), fzet(nzeta+1) (src/ModScbIO.f90:179-205) */
int rsg_scb_set_grid(rsg_scb* h, const double* thetaVal, const double* rhoVal, const double* zetaVal, const double* f,
                     const double* fzet);
int rsg_scb_set_geometry(rsg_scb* h, const double* x, const double* y, const double* z);
/* outputs of `pressure` (src/ModScbRun.f90:753-1199) that the hot path reads; for isotropy = 1
 * only dPdAlpha, dPdPsi are required */
int rsg_scb_set_pressure(rsg_scb* h, int isotropy, const double* pper, const double* ppar, const double* sigma,
                         const double* dPPerdTheta, const double* dPPerdRho, const double* dPPerdZeta, const double* dBsqdTheta,
                         const double* dBsqdRho, const double* dBsqdZeta, const double* dPPerdPsi, const double* dPPerdAlpha,
                         const double* dBsqdPsi, const double* dBsqdAlpha, const double* dPdAlpha, const double* dPdPsi);
int rsg_scb_set_field(rsg_scb* h, const char* name, const double* src);
int rsg_scb_get_field(rsg_scb* h, const char* name, double* dst);
int rsg_scb_field_size(rsg_scb* h, const char* name, long long* n);

int rsg_scb_bandjacob(rsg_scb* h, int* sorfail); /* computeBandJacob, src/ModScbCompute.f90:412-496 */
int rsg_scb_metrica(rsg_scb* h);                 /* src/ModScbEquation.f90:18-280  */
int rsg_scb_metric(rsg_scb* h);                  /* :283-540 */
int rsg_scb_newk(rsg_scb* h);                    /* :546-604 (Picard) */
int rsg_scb_newj(rsg_scb* h);                    /* :607-665 (Picard) */
/* iterateAlpha / iteratePsi (src/ModScbEuler.f90:160-299 / :469-612) incl. the extap /
 * theta-end / periodic-wrap post-processing.  Outputs mirror the module scalars nisave,
 * sumb, sumdb, diffmx and the logical SORFail; ni (may be NULL) receives the per-surface
 * iteration counts ni(1:npsi) / ni(1:nzeta). */
int rsg_scb_iterate_alpha(rsg_scb* h, double InConAlpha, int nimax, int theChange, int psiChange, int ordering, int* nisave,
                          double* sumb, double* sumdb, double* diffmx, int* sorfail, int* ni);
int rsg_scb_iterate_psi(rsg_scb* h, double InConPsi, int nimax, int theChange, int psiChange, int ordering, int* nisave,
                        double* sumb, double* sumdb, double* diffmx, int* sorfail, int* ni);
/* Compute_convergence (src/ModScbCompute.f90:499-754): fills jGradRho/Zeta/Theta, Jx..Jz,
 * GradPx..GradPz, jCrossB, GradP on the device and returns the three norms */
int rsg_scb_convergence(rsg_scb* h, double* normDiff, double* normJxB, double* normGradP, int* sorfail);
/* GSL_Derivs with the Steffen spline (src/ModRamGSL.f90:794-869) of a host (nthe,npsi,nzeta) field;
 * any output may be NULL */
int rsg_scb_derivs(rsg_scb* h, const double* f, double* dfdTheta, double* dfdRho, double* dfdZeta);
/* mapAlpha, mapPsi, mapTheta (src/ModScbEuler.f90:97-147, :403-457, :15-75) -- the re-gridding
 * steps that follow iterateAlpha / iteratePsi inside the SCB outer iteration (src/ModScbRun.f90:
 * 232-250, 418-430): x, y, z are moved along zeta / rho / theta lines by Steffen interpolation
 * (GSL_Interpolation_1D, src/ModRamGSL.f90:240-311 + src/RamGSL.c:111-174) so that alfa / psi / the
 * arc-length coordinate take their prescribed node values again; alfa / psi are reset (alfges,
 * psiges) and the periodic planes refreshed.  Everything stays on the device for the next
 * computeBandJacob.  set_map_targets: alphaVal(nzeta+1), psiVal(npsi), chiVal(nthe), once.
 * *sorfail: a line could not be interpolated (GSLerr > 0 in the reference => SORFail). */
int rsg_scb_set_map_targets(rsg_scb* h, const double* alphaVal, const double* psiVal, const double* chiVal);
int rsg_scb_map_alpha(rsg_scb* h, int* sorfail);
int rsg_scb_map_psi(rsg_scb* h, int* sorfail);
int rsg_scb_map_theta(rsg_scb* h, int* sorfail);
/* `pressure`, anisotropic branch from the normalised equatorial pressures on (src/ModScbRun.f90:1087-1175):
 * pperEq, pparEq are (npsi, nzeta+1) host arrays (already divided by pnormal, periodic columns set); the
 * library maps them along the field lines with the iLossCone = 1 | 2 formulas using bf / bsq of the last
 * computeBandJacob, builds sigma and tau, optionally reduces the anisotropy of mirror-unstable lines
 * (iReduceAnisotropy = 1, :1127-1160) and takes the Steffen derivatives dPPerd{Theta,Rho,Zeta,Psi,Alpha},
 * dBsqd{...} on the device (what rsg_scb_set_pressure would otherwise upload: 15 3-D arrays). */
int rsg_scb_pressure_aniso(rsg_scb* h, const double* pperEq, const double* pparEq, int iLossCone, int iReduceAnisotropy);
/* The 2-D FRONT END of `pressure` on the device (src/ModScbRun.f90:838-1086, anisotropic branch with RAM pressures): with it
 * an SCB outer iteration has no host hop left.  set_ram_pressure, once per scb_run (the RAM pressures do not change while SCB
 * iterates): PPerT, PParT (nS,NR,NT) of ram_run / ANISCH, scb[nS] = species%SCB, LZ(NR+1), PHI(NT); the library sums the
 * species on the RAM grid, extends it radially (PressMode 0 SKD | 1 ROE | 2 EXT | 3 FLT; ModScbParams default SKD) and smooths
 * (iSm2 0 none | 1 SavGol7 | 3 Gaussian | 4 both, default 4; SavGolIters default 11).  pressure_front = one `pressure` call:
 * equatorial foot points -> bilinear interpolation in (r^2, azimuth) (GSL_Interpolation_2D), extap inside 2 RE, floor, periodic
 * columns, normalisation, then the 3-D tail of rsg_scb_pressure_aniso; pperEq / pparEq (npsi, nzeta+1; may be NULL) return the
 * equatorial pressures.  rsg_scb_run with pressure == NULL uses it. */
int rsg_scb_set_ram_pressure(rsg_scb* h, int nS, int NR, int NT, const double* PPerT, const double* PParT, const int* scb,
                             const double* LZ, const double* PHI, int PressMode, int iSm2, int SavGolIters);
int rsg_scb_get_ram_pressure(rsg_scb* h, int* nX, int* nAz, double* rad2, double* azim, double* per, double* par);
int rsg_scb_pressure_front(rsg_scb* h, int iLossCone, int iReduceAnisotropy, double* pperEq, double* pparEq);
/* FLC_Radius (src/ModRamLoss.f90:176-336): curvature radius of the field lines and the zeta parameters of the
 * field-line-curvature scattering model from x, y, z and Bx, By, Bz of the last computeBandJacob (resident), interpolated
 * to the RAM equatorial points radRaw(i) (cos, sin)(azimRaw(j) 2 pi / 24 - pi) by GSL_Interpolation_2D (the 9-nearest-
 * neighbour rule).  r_curvEq, zeta1Eq, zeta2Eq (nR,nT) are what rsg_para_flc takes.  REarth = 6.4e6 m. */
int rsg_scb_flc_radius(rsg_scb* h, int nR, int nT, const double* radRaw, const double* azimRaw, double REarth, double* r_curvEq,
                       double* zeta1Eq, double* zeta2Eq);
/* Glue of the outer iteration (src/ModScbRun.f90:232-262, 418-440), so that alfa, psi, x, y, z need
 * not visit the host between the solves: device snapshots of a named field (alfaSav1, alphaPrev,
 * xPrev... of the reference; slot 0..3), the blend  field = snap(slot_new)*blend +
 * snap(slot_sav)*(1-blend)  (:236, :422) and MINVAL(jacobian(2:nthe-1,2:npsi-1,2:nzeta)) (:248, :434;
 * -1e300 if the Jacobian holds a NaN). */
int rsg_scb_snapshot(rsg_scb* h, const char* name, int slot);
int rsg_scb_restore(rsg_scb* h, const char* name, int slot);
int rsg_scb_blend(rsg_scb* h, const char* name, int slot_new, int slot_sav, double blend);
int rsg_scb_min_jacobian(rsg_scb* h, double* minjac);
/* Multi-GPU, iterateAlpha sharded along ZETA (the periodic azimuthal axis): a rank relaxes
 * the zeta planes (0-based rows of the (theta, zeta) problem) [k0, k0+nk) within 1..nzeta-1 of EVERY psi
 * surface, in the 4-colour order of RSG_SOR_COLOR4 cut at the row parity.  Protocol per sweep:
 *   half(0); exchange the EVEN edge planes with the zeta neighbours; half(1); exchange the ODD edge
 *   planes; all-reduce(MAX) the state vector (2*nsub doubles: residual maxima, failure flags); commit
 * (a plane k is the contiguous block alfa[k*nthe*npsi ...] of rsg_scb_field_device("alfa")).  commit
 * applies the loop control of src/ModScbEuler.f90:204-262 per surface (ni, EXIT on failure or
 * convergence, nimax); pending returns how many surfaces still iterate (synchronises; poll it every few
 * sweeps -- finished surfaces are skipped on the device).  Afterwards all-gather the planes and call
 * rsg_scb_iterate_finish(alpha = 1): alfa, ni, diffmx, sumb, sumdb are bit-identical to the one-GPU
 * RSG_SOR_COLOR4 solve.  (iteratePsi's sub-problems ARE zeta planes: rsg_scb_iterate_part.) */
int rsg_scb_zsolve_begin(rsg_scb* h, double InConAlpha, int nimax, int theChange, int psiChange, int k0, int nk);
int rsg_scb_zsolve_half(rsg_scb* h, int parity);
int rsg_scb_zsolve_state_device(rsg_scb* h, void** ptr, long long* n);
int rsg_scb_zsolve_commit(rsg_scb* h);
int rsg_scb_zsolve_pending(rsg_scb* h, int* pending);
/* scb_run (src/ModScbRun.f90:149-440, method = 2, iAMR = 0): the whole outer iteration of the Euler-
 * potential solve in ONE call, every 3-D array resident on the device -- the fused level of the drop-in,
 * like rsg_ram_run for ram_run.  Per outer iteration: computeBandJacob, pressure, metrica, newk,
 * iterateAlpha, blend + mapAlpha + mapTheta + computeBandJacob + MINVAL(jacobian) test (damped retries),
 * computeBandJacob, pressure, Compute_convergence, metric, newj, iteratePsi, blend + mapPsi + mapTheta + ...,
 * then the reference's exit test (errorAlpha/Psi < decreaseConv*, numit, MinSCBIterations).  On SORFail
 * x, y, z, alfa, psi are restored to their values at entry (:397-413).
 * `pressure` is the 2-D front end of the reference's routine (src/ModScbRun.f90:753-1086, RAM pressures ->
 * equatorial points; host): it gets xEq, yEq (npsi, nzeta+1; the foot points x/y(nThetaEquator,j,k)) and
 * fills the normalised pperEq, pparEq (npsi, nzeta+1, periodic columns set); return 0, non-zero aborts.
 * pressure == NULL: the front end runs on the device too (rsg_scb_set_ram_pressure before the call).
 * Anisotropic pressure (isotropy = 0, the reference's RAM-coupled mode) only.  Uses snapshot slots 0..2 of x, y,
 * z, alfa, psi.  Needs set_grid, set_geometry, set_map_targets. */
typedef int (*rsg_scb_pressure_fn)(void* user, int npsi, int nzetap, const double* xEq, const double* yEq, double* pperEq,
                                   double* pparEq);
typedef struct rsg_scb_run_params {
  double InConAlpha, InConPsi;                  /* 1e-6, 1e-6  (src/ModScbParams.f90:37-38) */
  double blendInitial, blendMin, blendMax;      /* 0.5 (src/ModScbRun.f90:178), 0.01, 1.0 (ModScbParams.f90:35-36) */
  double damp;                                  /* 0.9 (src/ModScbMain.f90:58) */
  double decreaseConvAlpha, decreaseConvPsi;    /* 0.5, 0.5 (ModScbParams.f90:29-32, ModScbRun.f90:84-87) */
  int nimax, theChange, psiChange;              /* 5001, 4, 0 */
  int numit, MinSCBIterations;                  /* 200, 11 (src/ModScbMain.f90:45, ModScbParams.f90:27) */
  int ordering;                                 /* RSG_SOR_LEX | RSG_SOR_COLOR4 */
  int iLossCone, iReduceAnisotropy;             /* 1, 0 */
} rsg_scb_run_params;
typedef struct rsg_scb_run_result {
  int iterations, iConvGlobal, SORFail, nisaveAlpha, nisavePsi, blendRetries;
  double blendAlpha, blendPsi, errorAlpha, errorPsi, sumbAlpha, sumdbAlpha, sumbPsi, sumdbPsi;
  double normDiffStart, normJxBStart, normGradPStart, normDiff, normJxB, normGradP;
} rsg_scb_run_result;
int rsg_scb_run(rsg_scb* h, const rsg_scb_run_params* p, rsg_scb_pressure_fn pressure, void* user, rsg_scb_run_result* out);
/* device time (CUDA events on the launching stream) of the kernels of the last call */
/* Multi-GPU: the independent sub-problems of a solve (psi surfaces for alpha, zeta planes for psi)
 * split among ranks.  part solves sub-problems [sub0, sub0+nsub) (0-based: q is jz = q+2 / k = q+2);
 * the caller all-gathers the solved planes of the field (rsg_scb_field_device gives the device
 * array, theta fastest: a[i + nthe*(j + npsi*k)]) and calls finish (sums, extrapolation, periodic
 * wrap) on every rank; nisave / diffmx / ni of finish cover this rank's sub-problems only.
 * rsg_scb_set_stream: run on the caller's stream (NULL: back to the handle's own). */
int rsg_scb_iterate_part(rsg_scb* h, int alpha, double tol, int nimax, int theChange, int psiChange, int ordering, int sub0,
                         int nsub);
int rsg_scb_iterate_finish(rsg_scb* h, int alpha, int theChange, int psiChange, int* nisave, double* sumb, double* sumdb,
                           double* diffmx, int* sorfail, int* ni);
int rsg_scb_field_device(rsg_scb* h, const char* name, void** ptr, long long* n);
int rsg_scb_set_stream(rsg_scb* h, void* stream);
double rsg_scb_last_ms(rsg_scb* h);
/* RSG_SOR_COLOR4 runs on thread-block clusters (2..8 CTAs per sub-problem) with the unknown and all
 * ten coefficient arrays resident in distributed shared memory for the whole solve when they fit
 * (default on); 0 selects one CTA per sub-problem with coefficients streamed from L2.  Results are
 * bit-identical either way.  rsg_scb_last_cluster: cluster size the last solve used (0 = none). */
int rsg_scb_use_cluster(rsg_scb* h, int on);
int rsg_scb_last_cluster(rsg_scb* h);
long long rsg_scb_launch_count(rsg_scb* h);

/* ---- the integral block of computehI (src/ModRamScb.f90:372-410) ---------------------
 * Replaces, for all RAM field lines at once, the loop that computes length / r0, fixes the equatorial
 * field, builds bfMirror and calls GSL_Integration_hI + GSL_BounceAverage (src/ModRamGSL.f90:125-200,
 * src/RamGSL.c:479-602), then assigns I_cart, H_cart, HDens_cart (nR,nT,nPa) and bZEq_Cart (nR,nT).
 * Inputs in the reference's shapes: chiVal(nthe), mu(nPa), xRAM / yRAM / zRAM / bRAM / density (nthe,nR,nT),
 * outsideMGNP(nR,nT); nThetaEquator is 1-based.  HDens_cart is in/out (skipped lines keep their value).
 * The h / I / bounce-average integrals are summed in closed form per grid segment (the integrands are built
 * on linear tables), i.e. they are the limit the reference's cquad(1e-3) calls approximate.
 * ms (may be NULL): device time of the kernel. */
int rsg_hI_integrals(int device, int nthe, int nR, int nT, int nPa, int nThetaEquator, double bnormal, const double* chiVal,
                     const double* mu, const double* xRAM, const double* yRAM, const double* zRAM, const double* bRAM,
                     const double* density, const int* outsideMGNP, double* I_cart, double* H_cart, double* HDens_cart,
                     double* bZEq_cart, double* ms);
/* The rest of computehI (src/ModRamScb.f90:413-637), on the arrays rsg_hI_integrals returns: scaling at the outer SCB
 * boundary (ScaleAt(nT): 1-based radial index of the first RAM point outside the SCB domain, 0 = none), MLT continuity,
 * near-90-degree corrections, negative / too-large repairs, GSL_Interpolation_1D (Steffen) of h and I from PA onto PAbn,
 * gaussian_kernel(1.0) / convolve smoothing when integral_smooth != 0 (srcExternal/gaussian_filter.f90), then the RAM
 * variables FNHS, FNIS, BOUNHS, BOUNIS, HDNS (nR+1,nT,nPa), BNES (nR+1,nT) (in: previous values, out: new) with dIdt,
 * dHdt, dIbndt, dBdt (DthI = TimeRamElapsed - TOld), the I = 1 row and the NaN repair.  Lz(nR+1), PA(nPa), PAbn(nPa).
 * EIR / EIP(1,J) = 0 (:611-612) stay with the caller.  *gslerr: number of lines the interpolation failed on. */
int rsg_hI_tail(int device, int nR, int nT, int nPa, double* I_cart, double* H_cart, double* HDens_cart, double* bZEq_cart,
                const int* ScaleAt, const int* outsideMGNP, const double* Lz, const double* PA, const double* PAbn,
                int integral_smooth, double DthI, double* FNHS, double* FNIS, double* BOUNHS, double* BOUNIS, double* HDNS,
                double* BNES, double* dIdt, double* dHdt, double* dIbndt, double* dBdt, int* gslerr, double* ms);
/* The first block of computehI, "Convert SCB field lines to RAM field lines" (src/ModRamScb.f90:252-300): winding-number
 * test of every RAM equatorial point (Lz(i+1), MLT(j)) against the outer SCB ring, psiRAM and then x, y, z, bf of every
 * node of the line by GSL_Interpolation_2D = Interpolation_2D_NN_point + NN_Interpolation_2D (src/ModRamGSL.f90:368-422,
 * :872-917: 9 nearest scattered points, inverse-distance-squared weights).  SCB arrays (nthe,npsi,nzeta+1), outputs
 * xRAM .. bRAM (nthe,nR,nT) (zero on lines outside the SCB domain) and outsideSCB(nR,nT).  The Geopack tracing of the
 * outside lines (:302-368) stays on the host. */
int rsg_hI_convert_lines(int device, int nthe, int npsi, int nzeta, int nR, int nT, int nThetaEquator, const double* x,
                         const double* y, const double* z, const double* bf, const double* psi, const double* alfa, const double* Lz,
                         const double* MLT, double* xRAM, double* yRAM, double* zRAM, double* bRAM, int* outsideSCB, double* ms);


/* ---- computehI resident on the device (src/ModRamScb.f90:249-637 in one object) ---------------------------------
 * The three calls above move every intermediate array through the host; rsg_hi keeps them on the device: the SCB arrays
 * come from the host once per call or straight from an rsg_scb handle, xRAM .. bRAM, I_cart .. bZEq_Cart stay resident
 * between the blocks, HDens_cart and the previous FNHS .. BNES persist between calls like the reference's module
 * variables (ModRamVariables), and the new field arrays go device-to-device into an rsg_ram handle
 * (rsg_hi_device_fields + rsg_ram_set_fields_device).  Same kernels, same results as the three stateless calls. */
typedef struct rsg_hi rsg_hi;
int rsg_hi_create(rsg_hi** out, int device, int nthe, int npsi, int nzeta, int nR, int nT, int nPa, int nThetaEquator, double bnormal,
                  const double* chiVal, const double* mu, const double* Lz, const double* MLT, const double* PA, const double* PAbn);
void rsg_hi_destroy(rsg_hi* h);
/* previous RAM variables (the module arrays at entry); HDens_cart may be NULL */
int rsg_hi_set_ram_fields(rsg_hi* h, const double* FNHS, const double* FNIS, const double* BOUNHS, const double* BOUNIS, const double* HDNS,
                          const double* BNES, const double* HDens_cart);
/* block 1 (:252-300): host SCB arrays x, y, z, bf, psi, alfa (nthe,npsi,nzeta+1), or all six NULL and an rsg_scb handle on
 * the same device.  outsideSCB(nR,nT) and *nOutside may be NULL (then nothing is read back). */
int rsg_hi_convert(rsg_hi* h, const double* x, const double* y, const double* z, const double* bf, const double* psi, const double* alfa,
                   rsg_scb* scb, int* outsideSCB, int* nOutside);
/* a line the host traced (Geopack, :330-362) for RAM point (i,j), 1-based: nthe nodes each of x, y, z, b / bnormal */
int rsg_hi_set_line(rsg_hi* h, int i, int j, const double* xl, const double* yl, const double* zl, const double* bl);
/* blocks 2 and 3.  ScaleAt(nT) + outsideMGNP(nR,nT) from the host's magnetopause logic (:306-329), or both NULL: the 'SWMF'
 * branch on the device (first outside point per MLT, every outside line flagged).  density(nthe,nR,nT) or NULL = RAIRDEN
 * (:362-371) on the device. */
int rsg_hi_finish(rsg_hi* h, const int* ScaleAt, const int* outsideMGNP, const double* density, int integral_smooth, double DthI,
                  int* gslerr);
/* convert + finish with the device-side defaults: nothing but the SCB arrays in (or nothing at all with an rsg_scb handle) */
int rsg_computehI(rsg_hi* h, const double* x, const double* y, const double* z, const double* bf, const double* psi, const double* alfa,
                  rsg_scb* scb, int integral_smooth, double DthI, int* gslerr);
/* results by name: xRAM yRAM zRAM bRAM density (nthe,nR,nT), psiRAM bZEq_cart (nR,nT), I_cart H_cart HDens_cart (nR,nT,nPa),
 * FNHS FNIS BOUNHS BOUNIS HDNS dIdt dHdt dIbndt (nR+1,nT,nPa), BNES dBdt (nR+1,nT) */
int rsg_hi_get(rsg_hi* h, const char* name, double* host);
/* which = 0 outsideSCB(nR,nT), 1 outsideMGNP(nR,nT), 2 ScaleAt(nT) */
int rsg_hi_get_int(rsg_hi* h, int which, int* host);
double rsg_hi_last_ms(rsg_hi* h);            /* device time convert .. tail of the last rsg_hi_finish */
long long rsg_hi_launch_count(rsg_hi* h);
/* device pointers of the new field arrays, in the order rsg_ram_set_fields_device takes them */
int rsg_hi_device_fields(rsg_hi* h, const double** ptrs9, const int** outsideMGNP);
/* rsg_ram_set_fields from DEVICE arrays: BNES, dBdt, FNHS, FNIS, BOUNHS, BOUNIS, HDNS, dIdt, dIbndt; outsideMGNP(NR,NT) */
int rsg_ram_set_fields_device(rsg_ram* h, const double* const* ptrs9, const int* d_outsideMGNP);

#ifdef __cplusplus
}
#endif
#endif /* RAMSCB_GPU_H */
