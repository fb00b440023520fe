!============================================================================
! Replacement bodies for the hot routines of src/ModRamLoss.f90: identical module
! name, public names and signatures (CEPARA(S), CHAREXCHANGE(S), ATMOL(S),
! FLCscatter(S)); the bodies call the C ABI (include/ramscb_gpu.h).  ram_run
! (src/ModRamRun.f90:67-183) compiles against it unchanged.
!
! FLC_Radius (src/ModRamLoss.f90:176-340) and PARA_FLC(S) (:342-455) are NOT on the
! device: they build FLC_coef from the SCB field-line geometry once per SCB call.
! Keep their reference text in this module unchanged (marked below); PARA_FLC only
! gains one line at its end that hands the species' coefficients to the library.
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

  ! --- keep src/ModRamLoss.f90:342-455 (subroutine PARA_FLC(S)) here and add, before its
  !     END SUBROUTINE:
  !        call rsg_flc_upload(S)

  subroutine rsg_flc_upload(S)
    ! FLC_coef(S,:,:,:,:) is strided (S is the fastest index): hand a contiguous copy over
    use ModRamGrids,     ONLY: NR, NT, NE, NPA
    use ModRamVariables, ONLY: FLC_coef
    integer, intent(in) :: S
    real(c_double), allocatable :: slab(:,:,:,:)
    allocate(slab(NR, NT, NE, NPA))
    slab = FLC_coef(S, :, :, :, :)
    call rsg_check(rsg_ram_set_flc_coef(hRam, int(S, c_int), slab), 'PARA_FLC')
    deallocate(slab)
  end subroutine rsg_flc_upload

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
