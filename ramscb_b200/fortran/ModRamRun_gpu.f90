!============================================================================
! Fused alternative to the species loop of ram_run (src/ModRamRun.f90:64-222):
! one C call advances all species with F2 resident on the device.  A maintainer
! replaces lines 64-222 of ram_run by `call ram_run_gpu` (everything before --
! the Volland-Stern potential, T = UTs -- stays).
!============================================================================
subroutine ram_run_gpu
  use ModRamGpu
  use ModRamParams,    ONLY: DoUseWPI, DoUseCoulomb, DoUseEMIC, DoUseFLC, DoUseBASdiff, DoUseKpDiff
  use ModRamGrids,     ONLY: nS
  use ModRamTiming,    ONLY: DTs, DtsMin, DtsNext, TimeRamElapsed, T, Dt_bc
  use ModRamVariables, ONLY: F2, VT, EIR, EIP, DtDriftR, DtDriftP, DtDriftE, DtDriftMu, SETRC, &
                             LSDR, LSCHA, LSATM, LSWAE, LSCOE, LSCSC, PPerT, PParT, XNE, AE, species
  use, intrinsic :: iso_c_binding
  implicit none
  integer(c_int) :: flags, gslerr
  real(c_double) :: dtn, dt(4, nS), ls(6, nS)
  integer :: iS

  ! What the fused step covers of :64-222: CEPARA .. the second DRIFTR, the epilogue, the pressure moments of ANISCH and
  ! (below) ANISCH's diffusion-coefficient rebuild.  NOT covered, and therefore refused instead of silently skipped:
  !  * the FLC stage (PARA_FLC + FLCscatter for species%FLC, :116-119 and :132-135): use the routine-level drop-ins of
  !    ModRamLoss_gpu.f90 with the reference's own species loop;
  !  * WAPARA_Kp (:178-180, DoUseWPI .and. DoUseBASdiff .and. DoUseKpDiff): a host routine that re-reads CDAAR; call it on
  !    the host and upload with rsg_ram_set_wave_tables before the next rebuild.
  if (DoUseFLC) call CON_stop('ram_run_gpu: DoUseFLC is not part of the fused step (use PARA_FLC / FLCscatter of ModRamLoss_gpu)')
  if (DoUseWPI .and. DoUseBASdiff .and. DoUseKpDiff) &
       call CON_stop('ram_run_gpu: DoUseKpDiff needs WAPARA_Kp on the host + rsg_ram_set_wave_tables (see ModRamRun_gpu.f90)')
  flags = 0
  if (DoUseWPI) flags = ior(flags, RSG_F_WPI)
  if (DoUseCoulomb) flags = ior(flags, RSG_F_COULOMB)
  if (DoUseEMIC) flags = ior(flags, RSG_F_EMIC)
  call rsg_check(rsg_ram_set_efield(hRam, VT, EIR, EIP), 'ram_run')            ! VT changes every call (:45-54)
  ! F2 up, the step, F2 back (outputs / restart / Compute3DFlux read it) in ONE call, pipelined over chunks of pitch angles.
  ! A host that leaves F2 untouched between steps calls rsg_ram_run instead and fetches F2 (rsg_ram_f2_d2h) when it writes output.
  call rsg_check(rsg_ram_run_host(hRam, F2, real(DTs, c_double), real(DtsMin, c_double), real(TimeRamElapsed, c_double), &
                                  flags, dtn, dt, ls, SETRC, PPerT, PParT), 'ram_run')
  DtsNext = dtn
  do iS = 1, nS
     DtDriftR(iS) = dt(1, iS); DtDriftP(iS) = dt(2, iS); DtDriftE(iS) = dt(3, iS); DtDriftMu(iS) = dt(4, iS)
     LSDR(iS) = LSDR(iS) + ls(1, iS); LSCHA(iS) = LSCHA(iS) + ls(2, iS); LSATM(iS) = LSATM(iS) + ls(3, iS)
     LSWAE(iS) = LSWAE(iS) + ls(4, iS); LSCOE(iS) = LSCOE(iS) + ls(5, iS); LSCSC(iS) = LSCSC(iS) + ls(6, iS)
  end do
  ! second half of ANISCH (:422-605), on the device: the coefficients the NEXT steps' WPADIF reads are rebuilt every
  ! Dt_bc from the tables of rsg_ram_set_wave_tables (uploaded once, after WAVEPARA / the EMIC table reader)
  if (MOD(INT(T), INT(Dt_bc)) == 0 .and. (DoUseWPI .or. DoUseEMIC)) then
     do iS = 1, nS
        if ((DoUseWPI .and. species(iS)%s_name == 'Electron') .or. (DoUseEMIC .and. species(iS)%EMIC)) then
           call rsg_check(rsg_anisch_diffcoef(hRam, int(iS, c_int), flags, XNE, int(AE, c_int), gslerr), 'ANISCH')
        end if
     end do
  end if
end subroutine ram_run_gpu

!============================================================================
! Routine-level drop-ins for the two remaining hot routines of MODULE ModRamRun:
! replace the bodies of SUMRC(S) (src/ModRamRun.f90:231-259) and of the moment part
! of ANISCH(S) (:343-415) by these (same names, same signatures).  The rest of
! ANISCH -- the diffusion-coefficient rebuild, :422-605 -- is ANISCH_diffcoef below.
!============================================================================
SUBROUTINE SUMRC(S)
  use ModRamGpu
  use ModRamVariables, ONLY: SETRC, ELORC
  use, intrinsic :: iso_c_binding
  implicit none
  integer, intent(in) :: S
  real(c_double) :: setrc_c, elorc_c
  call rsg_check(rsg_sumrc(hRam, int(S, c_int), setrc_c, elorc_c), 'SUMRC')
  SETRC(S) = setrc_c          ! new total, ELORC = old - new (:243-257)
  ELORC(S) = elorc_c
END SUBROUTINE SUMRC

SUBROUTINE ANISCH_moments(S)
  ! PPerT(S,:,:), PParT(S,:,:) (S is the fastest index: contiguous temporaries) and the
  ! side effect F2(S,I,J,K,1) = F2(S,I,J,K,2) (:366), which the device applies to its F2
  use ModRamGpu
  use ModRamGrids,     ONLY: NR, NT
  use ModRamVariables, ONLY: PPerT, PParT
  use, intrinsic :: iso_c_binding
  implicit none
  integer, intent(in) :: S
  real(c_double) :: pper(NR, NT), ppar(NR, NT)
  call rsg_check(rsg_anisch(hRam, int(S, c_int), pper, ppar), 'ANISCH')
  PPerT(S, :, :) = pper
  PParT(S, :, :) = ppar
END SUBROUTINE ANISCH_moments


SUBROUTINE ANISCH_diffcoef(S)
  ! second half of ANISCH(S) (src/ModRamRun.f90:422-605): ATAW / ATAC (electrons, DoUseWPI) or ATAW_emic_h / _he (EMIC
  ! species) rebuilt on the device, where WPADIF reads them; the host arrays are refreshed for diagnostics only
  use ModRamGpu
  use ModRamParams,    ONLY: DoUseWPI, DoUseEMIC
  use ModRamTiming,    ONLY: T, Dt_bc
  use ModRamVariables, ONLY: XNE, AE, ATAW, ATAC, ATAW_emic_h, ATAW_emic_he, species
  use, intrinsic :: iso_c_binding
  implicit none
  integer, intent(in) :: S
  integer(c_int) :: flags, gslerr
  if (MOD(INT(T), INT(Dt_bc)) /= 0) return
  flags = 0
  if (DoUseWPI) flags = ior(flags, RSG_F_WPI)
  if (DoUseEMIC) flags = ior(flags, RSG_F_EMIC)
  call rsg_check(rsg_anisch_diffcoef(hRam, int(S, c_int), flags, XNE, int(AE, c_int), gslerr), 'ANISCH')
  if (DoUseWPI .and. species(S)%s_name == 'Electron') then
     call rsg_check(rsg_ram_get_diffcoef(hRam, 0_c_int, ATAW), 'ANISCH')
     call rsg_check(rsg_ram_get_diffcoef(hRam, 1_c_int, ATAC), 'ANISCH')
  end if
  if (DoUseEMIC .and. species(S)%EMIC) then
     call rsg_check(rsg_ram_get_diffcoef(hRam, 2_c_int, ATAW_emic_h), 'ANISCH')
     call rsg_check(rsg_ram_get_diffcoef(hRam, 3_c_int, ATAW_emic_he), 'ANISCH')
  end if
END SUBROUTINE ANISCH_diffcoef

!============================================================================
! Multi-GPU: one MPI rank per GPU of a node.  ram_shard_init once after the device mirrors exist (the reference already
! initialises MPI, src/Main.f90:54-57); then ram_run_sharded_gpu replaces ram_run_gpu -- same outputs on every rank, F2
! sharded (this rank's pitch-angle slab of the host array goes up and comes back).
!============================================================================
subroutine ram_shard_init(policy)
  use ModRamGpu
  use ModRamMpi, ONLY: iProc, nProc, iComm      ! rank, size, communicator of the reference's MPI set-up
  use mpi
  use, intrinsic :: iso_c_binding
  implicit none
  integer, intent(in) :: policy                  ! 0: species first, 1: pitch-angle slabs of all species
  character(kind=c_char), target :: mine(192)
  character(kind=c_char), allocatable, target :: everyone(:)
  integer :: ierr
  allocate(everyone(192 * nProc))
  call rsg_check(rsg_ram_peer_export(hRam, c_loc(mine)), 'ram_shard_init')
  call MPI_Allgather(mine, 192, MPI_CHARACTER, everyone, 192, MPI_CHARACTER, iComm, ierr)
  call rsg_check(rsg_ram_peer_attach(hRam, int(iProc, c_int), int(nProc, c_int), int(policy, c_int), c_loc(everyone)), 'ram_shard_init')
  deallocate(everyone)
end subroutine ram_shard_init

subroutine ram_run_sharded_gpu
  use ModRamGpu
  use ModRamParams,    ONLY: DoUseWPI, DoUseCoulomb, DoUseEMIC
  use ModRamGrids,     ONLY: nS
  use ModRamTiming,    ONLY: DTs, DtsMin, DtsNext, TimeRamElapsed
  use ModRamVariables, ONLY: F2, VT, EIR, EIP, DtDriftR, DtDriftP, DtDriftE, DtDriftMu, SETRC, &
                             LSDR, LSCHA, LSATM, LSWAE, LSCOE, LSCSC, PPerT, PParT
  use, intrinsic :: iso_c_binding
  implicit none
  integer(c_int) :: flags
  real(c_double) :: dtn, dt(4, nS), ls(6, nS)
  integer :: iS
  flags = 0
  if (DoUseWPI) flags = ior(flags, RSG_F_WPI)
  if (DoUseCoulomb) flags = ior(flags, RSG_F_COULOMB)
  if (DoUseEMIC) flags = ior(flags, RSG_F_EMIC)
  call rsg_check(rsg_ram_set_efield(hRam, VT, EIR, EIP), 'ram_run_sharded')
  call rsg_check(rsg_ram_f2_h2d_shard(hRam, F2), 'ram_run_sharded')           ! this rank's share only
  call rsg_check(rsg_ram_run_sharded(hRam, real(DTs, c_double), real(DtsMin, c_double), real(TimeRamElapsed, c_double), &
                                     flags, dtn, dt, ls, SETRC, PPerT, PParT), 'ram_run_sharded')
  call rsg_check(rsg_ram_f2_d2h_shard(hRam, F2), 'ram_run_sharded')
  DtsNext = dtn
  do iS = 1, nS
     DtDriftR(iS) = dt(1, iS); DtDriftP(iS) = dt(2, iS); DtDriftE(iS) = dt(3, iS); DtDriftMu(iS) = dt(4, iS)
     LSDR(iS) = LSDR(iS) + ls(1, iS); LSCHA(iS) = LSCHA(iS) + ls(2, iS); LSATM(iS) = LSATM(iS) + ls(3, iS)
     LSWAE(iS) = LSWAE(iS) + ls(4, iS); LSCOE(iS) = LSCOE(iS) + ls(5, iS); LSCSC(iS) = LSCSC(iS) + ls(6, iS)
  end do
end subroutine ram_run_sharded_gpu
