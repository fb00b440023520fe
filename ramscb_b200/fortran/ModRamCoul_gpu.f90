!============================================================================
! Replacement for src/ModRamCoul.f90: identical module name, public names and
! signatures (COULPARA(S), COULEN(S), COULMU(S)); the bodies call the C ABI.
! NECR must have been handed over (rsg_ram_set_plasmasphere) after every
! plasmasphere update.  Shipped uncompiled, see ModRamGpu.f90.
!============================================================================
MODULE ModRamCoul

  use ModRamGpu
  use, intrinsic :: iso_c_binding

  implicit none

contains

  SUBROUTINE COULPARA(S)
    ! rate tables COULE, COULI, ATA, GTA for the step DTs (src/ModRamCoul.f90:17-125);
    ! built by the library from the grids of rsg_ram_set_grids and cached per DTs
    use ModRamTiming, ONLY: DTs
    integer, intent(in) :: S
    call rsg_check(rsg_coulpara(hRam, int(S, c_int), real(DTs, c_double)), 'COULPARA')
  END SUBROUTINE COULPARA

  SUBROUTINE COULEN(S)
    ! Coulomb energy drag, flux-limited sweep along K (:133-221)
    integer, intent(in) :: S
    call rsg_check(rsg_coulen(hRam, int(S, c_int)), 'COULEN')
  END SUBROUTINE COULEN

  SUBROUTINE COULMU(S)
    ! Coulomb pitch-angle scattering, implicit along L (:229-296); T arms the F2 < 0 clamp (:289)
    use ModRamTiming, ONLY: TimeRamElapsed
    integer, intent(in) :: S
    call rsg_check(rsg_coulmu(hRam, int(S, c_int), real(TimeRamElapsed, c_double)), 'COULMU')
  END SUBROUTINE COULMU

END MODULE ModRamCoul
