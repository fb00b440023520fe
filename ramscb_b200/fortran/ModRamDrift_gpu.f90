!============================================================================
! Replacement for src/ModRamDrift.f90: identical module name, public names and
! signatures (DRIFTPARA(S), DRIFTR(S), DRIFTP(S), DRIFTE(S), DRIFTMU(S), DRIFTEND);
! the bodies call the C ABI.  ram_run (src/ModRamRun.f90:64-185) compiles against
! it unchanged.  The device F2 is authoritative between the calls of one ram_run;
! ModRamRun_gpu.f90 uploads it before the species loop and downloads it after.
!============================================================================
MODULE ModRamDrift

  use ModRamMain,      ONLY: Real8_
  use ModRamVariables, ONLY: DtDriftR, DtDriftP, DtDriftE, DtDriftMu
  use ModRamGpu
  use, intrinsic :: iso_c_binding

  implicit none

contains

  SUBROUTINE DRIFTEND
    ! the device scratch is persistent: nothing to free (src/ModRamDrift.f90:23-30)
  END SUBROUTINE DRIFTEND

  SUBROUTINE DRIFTPARA(S)
    use ModRamTiming, ONLY: DTs
    integer, intent(in) :: S
    call rsg_check(rsg_driftpara(hRam, int(S, c_int), real(DTs, c_double)), 'DRIFTPARA')
  END SUBROUTINE DRIFTPARA

  SUBROUTINE DRIFTR(S)
    integer, intent(in) :: S
    real(c_double) :: dt4(4)
    call rsg_check(rsg_driftr(hRam, int(S, c_int)), 'DRIFTR')
    call rsg_check(rsg_get_dtdrift(hRam, int(S, c_int), dt4), 'DRIFTR')
    DtDriftR(S) = dt4(1)
  END SUBROUTINE DRIFTR

  SUBROUTINE DRIFTP(S)
    integer, intent(in) :: S
    real(c_double) :: dt4(4)
    call rsg_check(rsg_driftp(hRam, int(S, c_int)), 'DRIFTP')
    call rsg_check(rsg_get_dtdrift(hRam, int(S, c_int), dt4), 'DRIFTP')
    DtDriftP(S) = dt4(2)
  END SUBROUTINE DRIFTP

  SUBROUTINE DRIFTE(S)
    integer, intent(in) :: S
    real(c_double) :: dt4(4)
    call rsg_check(rsg_drifte(hRam, int(S, c_int)), 'DRIFTE')
    call rsg_check(rsg_get_dtdrift(hRam, int(S, c_int), dt4), 'DRIFTE')
    DtDriftE(S) = dt4(3)
  END SUBROUTINE DRIFTE

  SUBROUTINE DRIFTMU(S)
    integer, intent(in) :: S
    real(c_double) :: dt4(4)
    call rsg_check(rsg_driftmu(hRam, int(S, c_int)), 'DRIFTMU')
    call rsg_check(rsg_get_dtdrift(hRam, int(S, c_int), dt4), 'DRIFTMU')
    DtDriftMu(S) = dt4(4)
  END SUBROUTINE DRIFTMU

END MODULE ModRamDrift
