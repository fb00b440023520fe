!============================================================================
! Replacement bodies for the hot routines of src/ModRamWPI.f90: WAVELO(S) (:580-636)
! and WPADIF(S) (:643-714) keep their names and signatures and call the C ABI.
! The table builders WAVEPARA1/2, WAPARA_EMIC/HISS/CHORUS/Kp/BAS and I_emic
! (:18-578, :720-) stay as they are in the reference (host, once per run or per
! Dt_bc); after them the maintainer uploads what they produced:
!     rsg_ram_set_wavelo(hRam, WALOS1, WALOS2, WALOS3, Kp, Kpmax12)        after WAVEPARA1/2
!     rsg_ram_set_diffcoef(hRam, 0|1|2|3, ATAW|ATAC|ATAW_emic_h|ATAW_emic_he)  after the ANISCH rebuild
! (INTEGRATION.md section 2).  Shipped uncompiled, see ModRamGpu.f90.
!============================================================================
MODULE ModRamWPI

  use ModRamGpu
  use, intrinsic :: iso_c_binding

  implicit none

contains

  ! --- keep src/ModRamWPI.f90:18-578 (WAVEPARA1 ... WAPARA_BAS) here, unchanged ---

  SUBROUTINE WAVELO(S)
    ! electron lifetimes against wave scattering, F2 = F2*exp(-DTs/tau), :580-636
    use ModRamTiming, ONLY: DTs
    integer, intent(in) :: S
    call rsg_check(rsg_wavelo(hRam, int(S, c_int), real(DTs, c_double)), 'WAVELO')
  END SUBROUTINE WAVELO

  SUBROUTINE WPADIF(S)
    ! implicit pitch-angle diffusion, one Thomas line per (I,J,K), :643-714.  The
    ! reference prints and clamps F2 < 0 results (:703-707); the library clamps and
    ! returns how many cells it clamped.
    use ModRamTiming, ONLY: DTs
    integer, intent(in) :: S
    integer(c_long_long) :: nviol
    call rsg_check(rsg_wpadif(hRam, int(S, c_int), real(DTs, c_double), nviol), 'WPADIF')
  END SUBROUTINE WPADIF

  ! --- keep src/ModRamWPI.f90:720- (subroutine I_emic) here, unchanged ---

END MODULE ModRamWPI
