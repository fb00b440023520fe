!============================================================================
! Replacement bodies for the hot routines of src/ModScbCompute.f90: computeBandJacob
! (:412-496) and Compute_convergence (:499-754) keep their (argument-less) interfaces
! and call the C ABI.  `metrics` (:13-410, output diagnostics, once per SCB call)
! stays as it is in the reference.  Shipped uncompiled, see ModScbGpu.f90.
!============================================================================
MODULE ModScbCompute

  use ModScbGpu
  use, intrinsic :: iso_c_binding

  implicit none

contains

  ! --- keep src/ModScbCompute.f90:13-410 (subroutine metrics) here, unchanged ---

  SUBROUTINE computeBandJacob
    ! x, y, z are on the device already (scb_gpu_upload_domain, or moved there by mapAlpha /
    ! mapPsi / mapTheta).  All 30 outputs stay on the device for metrica / metric / newk / newj /
    ! Compute_convergence; the host gets back what scb_run, pressure and the output files read.
    use ModScbVariables, ONLY: jacobian, bf, bsq, Bx, By, Bz, SORFail
    integer(c_int) :: fail
    call rsg_scb_check(rsg_scb_bandjacob(hScb, fail), 'computeBandJacob')
    if (fail /= 0) SORFail = .true.                      ! GSLerr > 0 in the reference (:442-444)
    call rsg_scb_check(rsg_scb_get_field(hScb, 'jacobian'//c_null_char, jacobian), 'computeBandJacob')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'bf'//c_null_char, bf), 'computeBandJacob')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'bsq'//c_null_char, bsq), 'computeBandJacob')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'Bx'//c_null_char, Bx), 'computeBandJacob')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'By'//c_null_char, By), 'computeBandJacob')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'Bz'//c_null_char, Bz), 'computeBandJacob')
  END SUBROUTINE computeBandJacob

  SUBROUTINE Compute_convergence
    ! the three volume-weighted norms; jGradRho/Zeta/Theta, Jx..Jz, GradPx..GradPz, jCrossB, GradP
    ! are filled on the device (rsg_scb_get_field when an output needs them)
    use ModScbVariables, ONLY: normDiff, normJxB, normGradP, SORFail
    integer(c_int) :: fail
    real(c_double) :: nd, nj, ng
    call rsg_scb_check(rsg_scb_convergence(hScb, nd, nj, ng, fail), 'Compute_convergence')
    normDiff = nd; normJxB = nj; normGradP = ng
    if (fail /= 0) SORFail = .true.
  END SUBROUTINE Compute_convergence

END MODULE ModScbCompute
