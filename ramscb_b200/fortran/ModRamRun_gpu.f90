!============================================================================
! Fused alternative to the species loop of ram_run (src/ModRamRun.f90:64-222):
! one C call advances all species with F2 resident on the device.  A maintainer
! replaces lines 64-222 of ram_run by `call ram_run_gpu` (everything before --
! the Volland-Stern potential, T = UTs -- stays).
!============================================================================
subroutine ram_run_gpu
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
  call rsg_check(rsg_ram_set_efield(hRam, VT, EIR, EIP), 'ram_run')            ! VT changes every call (:45-54)
  call rsg_check(rsg_ram_f2_h2d(hRam, F2, 0_c_int), 'ram_run')                 ! skip when F2 was not touched on the host
  call rsg_check(rsg_ram_run(hRam, real(DTs, c_double), real(DtsMin, c_double), real(TimeRamElapsed, c_double), &
                             flags, dtn, dt, ls, SETRC, PPerT, PParT), 'ram_run')
  call rsg_check(rsg_ram_f2_d2h(hRam, F2, 0_c_int), 'ram_run')                 ! needed by outputs / restart / Compute3DFlux
  DtsNext = dtn
  do iS = 1, nS
     DtDriftR(iS) = dt(1, iS); DtDriftP(iS) = dt(2, iS); DtDriftE(iS) = dt(3, iS); DtDriftMu(iS) = dt(4, iS)
     LSDR(iS) = LSDR(iS) + ls(1, iS); LSCHA(iS) = LSCHA(iS) + ls(2, iS); LSATM(iS) = LSATM(iS) + ls(3, iS)
     LSWAE(iS) = LSWAE(iS) + ls(4, iS); LSCOE(iS) = LSCOE(iS) + ls(5, iS); LSCSC(iS) = LSCSC(iS) + ls(6, iS)
  end do
end subroutine ram_run_gpu

!============================================================================
! Routine-level drop-ins for the two remaining hot routines of MODULE ModRamRun:
! replace the bodies of SUMRC(S) (src/ModRamRun.f90:231-259) and of the moment part
! of ANISCH(S) (:343-415) by these (same names, same signatures).  The rest of
! ANISCH -- the diffusion-coefficient rebuild, :422-605 -- stays on the host.
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
