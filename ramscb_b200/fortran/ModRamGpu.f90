!============================================================================
! ModRamGpu -- ISO_C_BINDING interface to libramscb_gpu.so (include/ramscb_gpu.h)
!
! Drop-in shim for the RAM hot path of lanl/RAM-SCB.  The reference's operators
! are module procedures that take only the species index and work on the
! ModRamVariables globals (src/ModRamRun.f90:64-185).  With this file and the
! replacement bodies below, ModRamDrift / ModRamLoss / ModRamWPI / ModRamRun keep
! their names and signatures; the bodies forward c_loc() of the module arrays to
! the C ABI.  Binding style follows the reference's only existing C boundary,
! src/ModRamGSL.f90:10-71 <-> src/RamGSL.c, but passes the contiguous
! allocatables directly instead of copying them.
!
! NOTE: there is no Fortran compiler in the build container of this repository,
! so this file is shipped uncompiled; the C ABI it binds is exercised from
! Python/ctypes with the same pointers-and-sizes calling convention
! (ramscb_b200/host.py) and checked symbol by symbol in tests/test_cpu.py.
!============================================================================
module ModRamGpu

  use, intrinsic :: iso_c_binding
  implicit none

  type(c_ptr), save :: hRam = c_null_ptr     ! rsg_ram*

  integer(c_int), parameter :: RSG_MODE_EXACT = 0, RSG_MODE_FAST = 1
  integer(c_int), parameter :: RSG_F_WPI = 1, RSG_F_COULOMB = 2, RSG_F_EMIC = 4

  interface
     function rsg_ram_create(h, nS, nR, nT, nE, nPa, device) bind(C, name='rsg_ram_create') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), intent(out) :: h
       integer(c_int), value :: nS, nR, nT, nE, nPa, device
       integer(c_int) :: ierr
     end function
     function rsg_ram_destroy(h) bind(C, name='rsg_ram_destroy') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_mode(h, mode) bind(C, name='rsg_ram_set_mode') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: mode
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_grids(h, RLZ, LZ, EKEV, WE, DE, EBND, MU, WMU, DMU, UPA, GREL, GRBND, V, VBND, EPP, ERNH, &
                                RMAS, FFACTOR, QS, kind, khi, MDR, DPHI, CONF1, CONF2, BetaLim, FracCFL) &
          bind(C, name='rsg_ram_set_grids') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: RLZ(*), LZ(*), EKEV(*), WE(*), DE(*), EBND(*), MU(*), WMU(*), DMU(*), UPA(*), &
                                     GREL(*), GRBND(*), V(*), VBND(*), EPP(*), ERNH(*), RMAS(*), FFACTOR(*)
       integer(c_int), intent(in) :: QS(*), kind(*), khi(*)
       real(c_double), value :: MDR, DPHI, CONF1, CONF2, BetaLim, FracCFL
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_fields(h, BNES, dBdt, FNHS, FNIS, BOUNHS, BOUNIS, HDNS, dIdt, dIbndt, outsideMGNP) &
          bind(C, name='rsg_ram_set_fields') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: BNES(*), dBdt(*), FNHS(*), FNIS(*), BOUNHS(*), BOUNIS(*), HDNS(*), dIdt(*), dIbndt(*)
       integer(c_int), intent(in) :: outsideMGNP(*)
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_efield(h, VT, EIR, EIP) bind(C, name='rsg_ram_set_efield') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: VT(*), EIR(*), EIP(*)
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_boundary(h, FGEOS) bind(C, name='rsg_ram_set_boundary') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: FGEOS(*)
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_wavelo(h, WALOS1, WALOS2, WALOS3, Kp, Kpmax12) bind(C, name='rsg_ram_set_wavelo') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: WALOS1(*), WALOS2(*), WALOS3(*)
       real(c_double), value :: Kp, Kpmax12
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_diffcoef(h, which, D) bind(C, name='rsg_ram_set_diffcoef') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: which
       real(c_double), intent(in) :: D(*)
       integer(c_int) :: ierr
     end function
     function rsg_ram_f2_h2d(h, F2, S) bind(C, name='rsg_ram_f2_h2d') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: F2(*)
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_ram_f2_d2h(h, F2, S) bind(C, name='rsg_ram_f2_d2h') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: F2(*)
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_driftpara(h, S, DTs) bind(C, name='rsg_driftpara') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), value :: DTs
       integer(c_int) :: ierr
     end function
     function rsg_driftr(h, S) bind(C, name='rsg_driftr') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_driftp(h, S) bind(C, name='rsg_driftp') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_drifte(h, S) bind(C, name='rsg_drifte') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_driftmu(h, S) bind(C, name='rsg_driftmu') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_get_dtdrift(h, S, out4) bind(C, name='rsg_get_dtdrift') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), intent(out) :: out4(4)
       integer(c_int) :: ierr
     end function
     function rsg_cepara(h, S, DTs) bind(C, name='rsg_cepara') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), value :: DTs
       integer(c_int) :: ierr
     end function
     function rsg_charexchange(h, S) bind(C, name='rsg_charexchange') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_atmol(h, S) bind(C, name='rsg_atmol') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_wavelo(h, S, DTs) bind(C, name='rsg_wavelo') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), value :: DTs
       integer(c_int) :: ierr
     end function
     function rsg_wpadif(h, S, DTs, nviolation) bind(C, name='rsg_wpadif') result(ierr)
       import :: c_ptr, c_int, c_double, c_long_long
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), value :: DTs
       integer(c_long_long), intent(out) :: nviolation
       integer(c_int) :: ierr
     end function
     function rsg_sumrc(h, S, setrc, elorc) bind(C, name='rsg_sumrc') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), intent(out) :: setrc, elorc
       integer(c_int) :: ierr
     end function
     function rsg_anisch(h, S, PPERT_S, PPART_S) bind(C, name='rsg_anisch') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), intent(out) :: PPERT_S(*), PPART_S(*)
       integer(c_int) :: ierr
     end function
     ! ModRamCoul / FLCscatter (optional operators)
     function rsg_ram_set_plasmasphere(h, NECR) bind(C, name='rsg_ram_set_plasmasphere') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: NECR(*)
       integer(c_int) :: ierr
     end function
     function rsg_coulpara(h, S, DTs) bind(C, name='rsg_coulpara') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), value :: DTs
       integer(c_int) :: ierr
     end function
     function rsg_coulen(h, S) bind(C, name='rsg_coulen') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: S
       integer(c_int) :: ierr
     end function
     function rsg_coulmu(h, S, T) bind(C, name='rsg_coulmu') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), value :: T
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_flc_coef(h, S, FLC_coef) bind(C, name='rsg_ram_set_flc_coef') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), intent(in) :: FLC_coef(*)     ! contiguous copy of FLC_coef(S,:,:,:,:)
       integer(c_int) :: ierr
     end function
     function rsg_para_flc(h, S, r_curvEq, zeta1Eq, zeta2Eq) bind(C, name='rsg_para_flc') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), intent(in) :: r_curvEq(*), zeta1Eq(*), zeta2Eq(*)
       integer(c_int) :: ierr
     end function
     function rsg_flcscatter(h, S, DTs, T, Dt_bc, nviolation) bind(C, name='rsg_flcscatter') result(ierr)
       import :: c_ptr, c_int, c_double, c_long_long
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), value :: DTs, T, Dt_bc
       integer(c_long_long), intent(out) :: nviolation
       integer(c_int) :: ierr
     end function
     ! execution options of rsg_ram_run (both default on; results do not depend on them)
     function rsg_ram_use_fused(h, on) bind(C, name='rsg_ram_use_fused') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: on
       integer(c_int) :: ierr
     end function
     function rsg_ram_use_graph(h, on) bind(C, name='rsg_ram_use_graph') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: on
       integer(c_int) :: ierr
     end function
     function rsg_ram_run(h, DTs, DtsMin, T, flags, dts_next, DtDrift, losses, SETRC, PPERT, PPART) &
          bind(C, name='rsg_ram_run') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), value :: DTs, DtsMin, T
       integer(c_int), value :: flags
       real(c_double), intent(out) :: dts_next, DtDrift(*), losses(*), SETRC(*), PPERT(*), PPART(*)
       integer(c_int) :: ierr
     end function
     function rsg_ram_run_host(h, F2, DTs, DtsMin, T, flags, dts_next, DtDrift, losses, SETRC, PPERT, PPART) &
          bind(C, name='rsg_ram_run_host') result(ierr)
       ! rsg_ram_f2_h2d + rsg_ram_run + rsg_ram_f2_d2h in one pipelined call
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: F2(*)
       real(c_double), value :: DTs, DtsMin, T
       integer(c_int), value :: flags
       real(c_double), intent(out) :: dts_next, DtDrift(*), losses(*), SETRC(*), PPERT(*), PPART(*)
       integer(c_int) :: ierr
     end function
     ! ---- ANISCH, second half: the diffusion-coefficient rebuild on the device (src/ModRamRun.f90:422-605) ----
     function rsg_ram_set_wave_tables(h, ENG, NCF, ENOR, fpofc, NDAAJ, DAAR, use_bas, ENG_emic, NCF_emic, EKEV_emic, &
          fp2c_emic, Daa_emic_h, Daa_emic_he, Ihs_emic, Ihes_emic, PAbn) bind(C, name='rsg_ram_set_wave_tables') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: ENG, NCF, use_bas, ENG_emic, NCF_emic
       real(c_double), intent(in) :: ENOR(*), fpofc(*), NDAAJ(*), DAAR(*), EKEV_emic(*), fp2c_emic(*), Daa_emic_h(*), &
                                     Daa_emic_he(*), Ihs_emic(*), Ihes_emic(*), PAbn(*)
       integer(c_int) :: ierr
     end function
     function rsg_anisch_diffcoef(h, S, flags, XNE, AE, gslerr) bind(C, name='rsg_anisch_diffcoef') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S, flags, AE
       real(c_double), intent(in) :: XNE(*)
       integer(c_int), intent(out) :: gslerr
       integer(c_int) :: ierr
     end function
     function rsg_ram_get_diffcoef(h, which, D) bind(C, name='rsg_ram_get_diffcoef') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: which
       real(c_double), intent(out) :: D(*)
       integer(c_int) :: ierr
     end function
     ! ---- the multi-GPU step inside the library (one process / MPI rank per GPU; include/ramscb_gpu.h) ----
     function rsg_ram_peer_export(h, blob) bind(C, name='rsg_ram_peer_export') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       type(c_ptr), value :: blob            ! c_loc of RSG_PEER_BLOB_BYTES = 192 bytes
       integer(c_int) :: ierr
     end function
     function rsg_ram_peer_attach(h, rank, world, policy, blobs) bind(C, name='rsg_ram_peer_attach') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: rank, world, policy
       type(c_ptr), value :: blobs           ! c_loc of world x 192 bytes, in rank order
       integer(c_int) :: ierr
     end function
     function rsg_ram_peer_detach(h) bind(C, name='rsg_ram_peer_detach') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int) :: ierr
     end function
     function rsg_ram_run_sharded(h, DTs, DtsMin, T, flags, dts_next, DtDrift, losses, SETRC, PPERT, PPART) &
          bind(C, name='rsg_ram_run_sharded') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), value :: DTs, DtsMin, T
       integer(c_int), value :: flags
       real(c_double), intent(out) :: dts_next, DtDrift(*), losses(*), SETRC(*), PPERT(*), PPART(*)
       integer(c_int) :: ierr
     end function
     function rsg_ram_f2_h2d_shard(h, F2) bind(C, name='rsg_ram_f2_h2d_shard') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: F2(*)
       integer(c_int) :: ierr
     end function
     function rsg_ram_f2_d2h_shard(h, F2) bind(C, name='rsg_ram_f2_d2h_shard') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: F2(*)
       integer(c_int) :: ierr
     end function
     function rsg_geosb(h, S, FluxLanl, s_comp) bind(C, name='rsg_geosb') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: S
       real(c_double), intent(in) :: FluxLanl(*)
       real(c_double), value :: s_comp
       integer(c_int) :: ierr
     end function
     function rsg_get_electric_field(h, vols, VTOL, VTN, TimeRamElapsed, TOLV, DtEfi, Kp, PHI, PHIOFS, VT_out) &
          bind(C, name='rsg_get_electric_field') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: vols
       real(c_double), intent(in) :: VTOL(*), VTN(*), PHI(*)
       real(c_double), value :: TimeRamElapsed, TOLV, DtEfi, Kp, PHIOFS
       real(c_double), intent(out) :: VT_out(*)
       integer(c_int) :: ierr
     end function
     function rsg_host_register(p, bytes) bind(C, name='rsg_host_register') result(ierr)
       import :: c_ptr, c_int, c_long_long
       type(c_ptr), value :: p
       integer(c_long_long), value :: bytes
       integer(c_int) :: ierr
     end function
  end interface

contains

  subroutine rsg_check(ierr, where)
    ! No return codes cross the reference's interfaces: a failure of the device path is fatal,
    ! exactly like the reference's CON_stop (src/Main.f90:135-168).
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: where
    if (ierr /= 0) call CON_stop('ramscb_gpu failure in '//where)
  end subroutine rsg_check

end module ModRamGpu
