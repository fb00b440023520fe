!============================================================================
! Fused alternative to the outer iteration of scb_run (src/ModScbRun.f90:134-440,
! method = 2, iAMR = 0): one C call runs every outer iteration with x, y, z, alfa,
! psi and all coefficient arrays resident on the device.  A maintainer replaces
! lines 134-429 of scb_run by `call scb_run_gpu`; the tail (entropy, metrics,
! bounextp, outputs, :431-447) stays and reads the arrays scb_run_gpu brings back.
!
! `pressure` is split where its 3-D part starts (:1087): the reference's 2-D front
! end (RAM pressures -> equatorial foot points, smoothing; :753-1086) becomes the
! callback below -- it must fill pperEq / pparEq from the xEq / yEq it is given
! instead of from x(nThetaEquator,:,:), y(nThetaEquator,:,:).
! Shipped uncompiled, see ModScbGpu.f90.
!============================================================================
module ModScbRunGpu

  use ModScbGpu
  use, intrinsic :: iso_c_binding
  implicit none

contains

  function scb_pressure_front_end(user, npsi_c, nzetap_c, xEq, yEq, pperEq, pparEq) bind(C) result(ierr)
    ! rsg_scb_pressure_fn.  pressure_equatorial(xEq, yEq, pperEq, pparEq) is lines 753-1086 of the
    ! reference's `pressure` with radGrid/angleGrid taken from the arguments (normalised by pnormal,
    ! periodic columns k = 1 and nzeta+1 set, :1076-1083).
    type(c_ptr), value :: user
    integer(c_int), value :: npsi_c, nzetap_c
    real(c_double), intent(in) :: xEq(npsi_c, nzetap_c), yEq(npsi_c, nzetap_c)
    real(c_double), intent(out) :: pperEq(npsi_c, nzetap_c), pparEq(npsi_c, nzetap_c)
    integer(c_int) :: ierr
    external :: pressure_equatorial
    call pressure_equatorial(xEq, yEq, pperEq, pparEq)
    ierr = 0
  end function scb_pressure_front_end

  subroutine scb_run_gpu
    use ModScbMain,      ONLY: damp, numit, nimax
    use ModScbParams,    ONLY: InConAlpha, InConPsi, blendMin, blendMax, MinSCBIterations, theChange, psiChange, &
                               iLossCone, iReduceAnisotropy
    use ModScbVariables, ONLY: x, y, z, alfa, psi, SORFail, hICalc, iteration, iConvGlobal, nisave, sumb, sumdb, &
                               blendAlpha, blendPsi, errorAlpha, errorPsi, decreaseConvAlpha, decreaseConvPsi, &
                               normDiff, normJxB, normGradP
    type(rsg_scb_run_params) :: p
    type(rsg_scb_run_result) :: r

    p%InConAlpha = InConAlpha; p%InConPsi = InConPsi
    p%blendInitial = 0.5_c_double                                   ! src/ModScbRun.f90:178
    p%blendMin = blendMin; p%blendMax = blendMax; p%damp = damp
    p%decreaseConvAlpha = decreaseConvAlpha; p%decreaseConvPsi = decreaseConvPsi   ! set at :84-87
    p%nimax = nimax; p%theChange = theChange; p%psiChange = psiChange
    p%numit = numit; p%MinSCBIterations = MinSCBIterations
    p%ordering = scbSorOrdering; p%iLossCone = iLossCone; p%iReduceAnisotropy = iReduceAnisotropy

    call scb_gpu_upload_domain                                       ! x, y, z, alfa, psi, grids, map targets
    call rsg_scb_check(rsg_scb_run(hScb, p, c_funloc(scb_pressure_front_end), c_null_ptr, r), 'scb_run')

    iteration = r%iterations; iConvGlobal = r%iConvGlobal
    nisave = r%nisavePsi; sumb = r%sumbPsi; sumdb = r%sumdbPsi
    blendAlpha = r%blendAlpha; blendPsi = r%blendPsi; errorAlpha = r%errorAlpha; errorPsi = r%errorPsi
    normDiff = r%normDiff; normJxB = r%normJxB; normGradP = r%normGradP
    SORFail = r%SORFail /= 0
    if (SORFail) then
       hICalc = .false.                                              ! :408; the device has restored the start state
       return
    end if
    ! what the tail of scb_run, computehI and the output files read
    call rsg_scb_check(rsg_scb_get_field(hScb, 'x'//c_null_char, x), 'scb_run')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'y'//c_null_char, y), 'scb_run')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'z'//c_null_char, z), 'scb_run')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'alfa'//c_null_char, alfa), 'scb_run')
    call rsg_scb_check(rsg_scb_get_field(hScb, 'psi'//c_null_char, psi), 'scb_run')
  end subroutine scb_run_gpu

end module ModScbRunGpu
