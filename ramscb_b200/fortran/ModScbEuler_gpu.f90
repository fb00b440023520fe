!============================================================================
! Replacement bodies for the hot routines of src/ModScbEuler.f90: iterateAlpha
! (:160-299), iteratePsi (:469-612) and the re-gridding steps mapAlpha (:97-147),
! mapPsi (:403-457), mapTheta (:15-75) keep their argument-less interfaces and call
! the C ABI.  alfges, psiges, directAlpha, directPsi, psiFunctions, InterpolatePsiR
! stay as they are in the reference (1-D / initialisation work).
! alfa, psi, x, y, z stay on the device between the calls of one outer iteration of
! scb_run (src/ModScbRun.f90:149-429); scb_gpu_download_state brings them back when
! the host reads them (pressure's 2-D front end, computehI, output files).
! Shipped uncompiled, see ModScbGpu.f90.
!============================================================================
MODULE ModScbEuler

  use ModScbGpu
  use, intrinsic :: iso_c_binding

  implicit none

contains

  ! --- keep alfges, directAlpha, psiFunctions, InterpolatePsiR, psiges, directPsi of
  !     src/ModScbEuler.f90 here, unchanged ---

  SUBROUTINE iterateAlpha
    use ModScbMain,      ONLY: nimax
    use ModScbParams,    ONLY: InConAlpha, psiChange, theChange
    use ModScbVariables, ONLY: nisave, sumb, sumdb, diffmx, SORFail
    integer(c_int) :: ni_c, fail
    real(c_double) :: sb, sdb, dmx
    call rsg_scb_check(rsg_scb_iterate_alpha(hScb, real(InConAlpha, c_double), int(nimax, c_int), int(theChange, c_int), &
                                             int(psiChange, c_int), scbSorOrdering, ni_c, sb, sdb, dmx, fail, c_null_ptr), &
                       'iterateAlpha')
    nisave = ni_c; sumb = sb; sumdb = sdb; diffmx = dmx
    if (fail /= 0) SORFail = .true.                      ! NaN or >= 1e10 iterate (:226-240)
  END SUBROUTINE iterateAlpha

  SUBROUTINE iteratePsi
    use ModScbMain,      ONLY: nimax
    use ModScbParams,    ONLY: InConPsi, psiChange, theChange
    use ModScbVariables, ONLY: nisave, sumb, sumdb, diffmx, SORFail
    integer(c_int) :: ni_c, fail
    real(c_double) :: sb, sdb, dmx
    call rsg_scb_check(rsg_scb_iterate_psi(hScb, real(InConPsi, c_double), int(nimax, c_int), int(theChange, c_int), &
                                           int(psiChange, c_int), scbSorOrdering, ni_c, sb, sdb, dmx, fail, c_null_ptr), &
                       'iteratePsi')
    nisave = ni_c; sumb = sb; sumdb = sdb; diffmx = dmx
    if (fail /= 0) SORFail = .true.                      ! :539-553
  END SUBROUTINE iteratePsi

  SUBROUTINE mapAlpha
    ! x, y, z moved along the zeta lines so that alfa = alphaVal(k) again; alfa reset (alfges),
    ! periodic planes refreshed -- all on the device
    use ModScbVariables, ONLY: SORFail
    integer(c_int) :: fail
    call rsg_scb_check(rsg_scb_map_alpha(hScb, fail), 'mapAlpha')
    if (fail /= 0) SORFail = .true.                      ! GSLerr > 0 (:123-128)
  END SUBROUTINE mapAlpha

  SUBROUTINE mapPsi
    use ModScbVariables, ONLY: SORFail
    integer(c_int) :: fail
    call rsg_scb_check(rsg_scb_map_psi(hScb, fail), 'mapPsi')
    if (fail /= 0) SORFail = .true.                      ! :432-437
  END SUBROUTINE mapPsi

  SUBROUTINE mapTheta
    use ModScbVariables, ONLY: SORFail
    integer(c_int) :: fail
    call rsg_scb_check(rsg_scb_map_theta(hScb, fail), 'mapTheta')
    if (fail /= 0) SORFail = .true.                      ! :52-57
  END SUBROUTINE mapTheta

  subroutine scb_gpu_download_state
    ! x, y, z, alfa, psi as the device holds them (before pressure's front end, computehI, outputs)
    use ModScbVariables, ONLY: x, y, z, alfa, psi
    call rsg_scb_check(rsg_scb_get_field(hScb, 'x'//c_null_char, x), 'download_state')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'y'//c_null_char, y), 'download_state')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'z'//c_null_char, z), 'download_state')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'alfa'//c_null_char, alfa), 'download_state')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'psi'//c_null_char, psi), 'download_state')
  end subroutine scb_gpu_download_state

END MODULE ModScbEuler
