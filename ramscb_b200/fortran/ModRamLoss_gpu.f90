!============================================================================
! Replacement bodies for the hot routines of src/ModRamLoss.f90: identical module
! name, public names and signatures (CEPARA(S), CHAREXCHANGE(S), ATMOL(S),
! FLCscatter(S)); the bodies call the C ABI (include/ramscb_gpu.h).  ram_run
! (src/ModRamRun.f90:67-183) compiles against it unchanged.
!
! FLC_Radius (src/ModRamLoss.f90:176-340) is NOT on the device: it derives the (NR,NT)
! curvature radius and zeta parameters from the SCB field-line geometry with 2-D
! interpolation once per Dt_bc.  Keep its reference text in this module unchanged
! (marked below).  PARA_FLC(S) (:342-455) builds FLC_coef on the device from them.
!
! Shipped uncompiled (no Fortran compiler in the build container): see the note in
! ModRamGpu.f90.  The C entry points bound here are exercised with the same
! arguments from ramscb_b200/host.py (tests/test_ram_parity_gpu.py).
!============================================================================
MODULE ModRamLoss

  use ModRamGpu
  use, intrinsic :: iso_c_binding

  implicit none

contains

  SUBROUTINE CEPARA(S)
    ! src/ModRamLoss.f90:19-170: CHARGE and ATLOS for the time step DTs.  The device
    ! recomputes CHARGE = exp(-sigma*V*HDNS*DTs) inside the loss kernels instead of
    ! storing the (nS,NR,NT,NE,NPA) array, so the host arrays ACHAR / ATLOS are not filled.
    use ModRamTiming, ONLY: DTs
    integer, intent(in) :: S
    call rsg_check(rsg_cepara(hRam, int(S, c_int), real(DTs, c_double)), 'CEPARA')
  END SUBROUTINE CEPARA

  ! --- keep src/ModRamLoss.f90:176-340 (subroutine FLC_Radius) here, unchanged ---

  subroutine PARA_FLC(S)
    ! src/ModRamLoss.f90:342-455 on the device: FLC_coef of species S is built from the (NR,NT) outputs of
    ! FLC_Radius where FLCscatter reads it; the host array FLC_coef is no longer filled (the optional
    ! DoWriteFLCDiffCoeff dump would need rsg_ram_get_flc_coef).  Same "every Dt_bc" gate as :371.
    use ModRamTiming,    ONLY: TimeRamElapsed, Dt_bc
    use ModRamVariables, ONLY: r_curvEq, zeta1Eq, zeta2Eq
    integer, intent(in) :: S
    if (mod(int(TimeRamElapsed), int(Dt_bc)) .gt. 1e-6) return
    call rsg_check(rsg_para_flc(hRam, int(S, c_int), r_curvEq, zeta1Eq, zeta2Eq), 'PARA_FLC')
  end subroutine PARA_FLC

  subroutine FLC_Radius
    ! src/ModRamLoss.f90:176-336 on the device: curvature radius and zeta parameters from the SCB geometry / field that
    ! computeBandJacob left resident (hScb), interpolated to the RAM equatorial points; same "every Dt_bc" gate as :207
    use ModRamGrids,     ONLY: NR, NT
    use ModRamTiming,    ONLY: TimeRamElapsed, Dt_bc
    use ModRamVariables, ONLY: r_curvEq, zeta1Eq, zeta2Eq
    use ModScbMain,      ONLY: REarth
    use ModScbVariables, ONLY: radRaw, azimRaw
    use ModScbGpu,       ONLY: hScb, rsg_scb_flc_radius
    if (mod(int(TimeRamElapsed), int(Dt_bc)) .gt. 1e-6) return
    call rsg_check(rsg_scb_flc_radius(hScb, int(NR, c_int), int(NT, c_int), radRaw(1:NR), azimRaw, real(REarth, c_double), &
                                      r_curvEq, zeta1Eq, zeta2Eq), 'FLC_Radius')
  end subroutine FLC_Radius

  SUBROUTINE CHAREXCHANGE(S)
    ! src/ModRamLoss.f90:457-478
    integer, intent(in) :: S
    call rsg_check(rsg_charexchange(hRam, int(S, c_int)), 'CHAREXCHANGE')
  END SUBROUTINE CHAREXCHANGE

  SUBROUTINE ATMOL(S)
    ! src/ModRamLoss.f90:485-507
    integer, intent(in) :: S
    call rsg_check(rsg_atmol(hRam, int(S, c_int)), 'ATMOL')
  END SUBROUTINE ATMOL

  subroutine FLCscatter(S)
    ! src/ModRamLoss.f90:513-575; the library skips the operator while
    ! TimeRamElapsed < Dt_bc exactly like :523
    use ModRamTiming, ONLY: DTs, TimeRamElapsed, Dt_bc
    integer, intent(in) :: S
    integer(c_long_long) :: nviol
    call rsg_check(rsg_flcscatter(hRam, int(S, c_int), real(DTs, c_double), real(TimeRamElapsed, c_double), &
                                  real(Dt_bc, c_double), nviol), 'FLCscatter')
  end subroutine FLCscatter

END MODULE ModRamLoss
