!============================================================================
! ModScbGpu -- ISO_C_BINDING interface to the SCB half of libramscb_gpu.so
! (include/ramscb_gpu.h, rsg_scb_*).  Same binding style as ModRamGpu.f90 and as
! the reference's own C boundary (src/ModRamGSL.f90:10-71).  Field names passed
! to rsg_scb_set_field / rsg_scb_get_field are the reference's ModScbVariables
! names ('alfa', 'psi', 'x', 'jacobian', 'bsq', 'GradRhoSq', 'vecd', ...),
! NUL-terminated.  Shipped uncompiled (no Fortran compiler in the build container).
!============================================================================
module ModScbGpu

  use, intrinsic :: iso_c_binding
  implicit none

  type(c_ptr), save :: hScb = c_null_ptr     ! rsg_scb*

  integer(c_int), parameter :: RSG_SOR_LEX = 0, RSG_SOR_COLOR4 = 1
  ! production ordering; RSG_SOR_LEX reproduces the reference's sweep order bit for bit
  integer(c_int), save :: scbSorOrdering = RSG_SOR_COLOR4

  ! rsg_scb_run_params / rsg_scb_run_result of include/ramscb_gpu.h (same member order)
  type, bind(C) :: rsg_scb_run_params
     real(c_double) :: InConAlpha, InConPsi, blendInitial, blendMin, blendMax, damp, decreaseConvAlpha, decreaseConvPsi
     integer(c_int) :: nimax, theChange, psiChange, numit, MinSCBIterations, ordering, iLossCone, iReduceAnisotropy
  end type
  type, bind(C) :: rsg_scb_run_result
     integer(c_int) :: iterations, iConvGlobal, SORFail, nisaveAlpha, nisavePsi, blendRetries
     real(c_double) :: blendAlpha, blendPsi, errorAlpha, errorPsi, sumbAlpha, sumdbAlpha, sumbPsi, sumdbPsi
     real(c_double) :: normDiffStart, normJxBStart, normGradPStart, normDiff, normJxB, normGradP
  end type

  interface
     function rsg_scb_create(h, nthe, npsi, nzeta, device) bind(C, name='rsg_scb_create') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), intent(out) :: h
       integer(c_int), value :: nthe, npsi, nzeta, device
       integer(c_int) :: ierr
     end function
     function rsg_scb_destroy(h) bind(C, name='rsg_scb_destroy') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int) :: ierr
     end function
     function rsg_scb_set_grid(h, thetaVal, rhoVal, zetaVal, f, fzet) bind(C, name='rsg_scb_set_grid') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: thetaVal(*), rhoVal(*), zetaVal(*), f(*), fzet(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_set_geometry(h, x, y, z) bind(C, name='rsg_scb_set_geometry') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: x(*), y(*), z(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_set_pressure(h, isotropy, pper, ppar, sigma, dPPerdTheta, dPPerdRho, dPPerdZeta, dBsqdTheta, &
                                   dBsqdRho, dBsqdZeta, dPPerdPsi, dPPerdAlpha, dBsqdPsi, dBsqdAlpha, dPdAlpha, dPdPsi) &
          bind(C, name='rsg_scb_set_pressure') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: isotropy
       real(c_double), intent(in) :: pper(*), ppar(*), sigma(*), dPPerdTheta(*), dPPerdRho(*), dPPerdZeta(*), &
                                     dBsqdTheta(*), dBsqdRho(*), dBsqdZeta(*), dPPerdPsi(*), dPPerdAlpha(*), &
                                     dBsqdPsi(*), dBsqdAlpha(*), dPdAlpha(*), dPdPsi(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_flc_radius(h, nR, nT, radRaw, azimRaw, REarth, r_curvEq, zeta1Eq, zeta2Eq) &
          bind(C, name='rsg_scb_flc_radius') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: nR, nT
       real(c_double), value :: REarth
       real(c_double), intent(in) :: radRaw(*), azimRaw(*)
       real(c_double), intent(out) :: r_curvEq(*), zeta1Eq(*), zeta2Eq(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_set_ram_pressure(h, nS, NR, NT, PPerT, PParT, scb, LZ, PHI, PressMode, iSm2, SavGolIters) &
          bind(C, name='rsg_scb_set_ram_pressure') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: nS, NR, NT, PressMode, iSm2, SavGolIters
       real(c_double), intent(in) :: PPerT(*), PParT(*), LZ(*), PHI(*)
       integer(c_int), intent(in) :: scb(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_pressure_front(h, iLossCone, iReduceAnisotropy, pperEq, pparEq) &
          bind(C, name='rsg_scb_pressure_front') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: iLossCone, iReduceAnisotropy
       real(c_double), intent(out) :: pperEq(*), pparEq(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_pressure_aniso(h, pperEq, pparEq, iLossCone, iReduceAnisotropy) &
          bind(C, name='rsg_scb_pressure_aniso') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: pperEq(*), pparEq(*)
       integer(c_int), value :: iLossCone, iReduceAnisotropy
       integer(c_int) :: ierr
     end function
     function rsg_scb_set_field(h, name, src) bind(C, name='rsg_scb_set_field') result(ierr)
       import :: c_ptr, c_int, c_double, c_char
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: name(*)
       real(c_double), intent(in) :: src(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_get_field(h, name, dst) bind(C, name='rsg_scb_get_field') result(ierr)
       import :: c_ptr, c_int, c_double, c_char
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: name(*)
       real(c_double), intent(inout) :: dst(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_bandjacob(h, sorfail) bind(C, name='rsg_scb_bandjacob') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), intent(out) :: sorfail
       integer(c_int) :: ierr
     end function
     function rsg_scb_metrica(h) bind(C, name='rsg_scb_metrica') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int) :: ierr
     end function
     function rsg_scb_metric(h) bind(C, name='rsg_scb_metric') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int) :: ierr
     end function
     function rsg_scb_newk(h) bind(C, name='rsg_scb_newk') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int) :: ierr
     end function
     function rsg_scb_newj(h) bind(C, name='rsg_scb_newj') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int) :: ierr
     end function
     function rsg_scb_iterate_alpha(h, InConAlpha, nimax, theChange, psiChange, ordering, nisave, sumb, sumdb, diffmx, &
                                    sorfail, ni) bind(C, name='rsg_scb_iterate_alpha') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), value :: InConAlpha
       integer(c_int), value :: nimax, theChange, psiChange, ordering
       integer(c_int), intent(out) :: nisave, sorfail
       real(c_double), intent(out) :: sumb, sumdb, diffmx
       type(c_ptr), value :: ni                         ! int ni(npsi), or c_null_ptr
       integer(c_int) :: ierr
     end function
     function rsg_scb_iterate_psi(h, InConPsi, nimax, theChange, psiChange, ordering, nisave, sumb, sumdb, diffmx, &
                                  sorfail, ni) bind(C, name='rsg_scb_iterate_psi') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), value :: InConPsi
       integer(c_int), value :: nimax, theChange, psiChange, ordering
       integer(c_int), intent(out) :: nisave, sorfail
       real(c_double), intent(out) :: sumb, sumdb, diffmx
       type(c_ptr), value :: ni                         ! int ni(nzeta), or c_null_ptr
       integer(c_int) :: ierr
     end function
     function rsg_scb_convergence(h, normDiff, normJxB, normGradP, sorfail) bind(C, name='rsg_scb_convergence') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(out) :: normDiff, normJxB, normGradP
       integer(c_int), intent(out) :: sorfail
       integer(c_int) :: ierr
     end function
     function rsg_scb_set_map_targets(h, alphaVal, psiVal, chiVal) bind(C, name='rsg_scb_set_map_targets') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: alphaVal(*), psiVal(*), chiVal(*)
       integer(c_int) :: ierr
     end function
     function rsg_scb_map_alpha(h, sorfail) bind(C, name='rsg_scb_map_alpha') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), intent(out) :: sorfail
       integer(c_int) :: ierr
     end function
     function rsg_scb_map_psi(h, sorfail) bind(C, name='rsg_scb_map_psi') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), intent(out) :: sorfail
       integer(c_int) :: ierr
     end function
     function rsg_scb_map_theta(h, sorfail) bind(C, name='rsg_scb_map_theta') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), intent(out) :: sorfail
       integer(c_int) :: ierr
     end function
     function rsg_scb_snapshot(h, name, slot) bind(C, name='rsg_scb_snapshot') result(ierr)
       import :: c_ptr, c_int, c_char
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: name(*)
       integer(c_int), value :: slot
       integer(c_int) :: ierr
     end function
     function rsg_scb_restore(h, name, slot) bind(C, name='rsg_scb_restore') result(ierr)
       import :: c_ptr, c_int, c_char
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: name(*)
       integer(c_int), value :: slot
       integer(c_int) :: ierr
     end function
     function rsg_scb_blend(h, name, slot_new, slot_sav, blend) bind(C, name='rsg_scb_blend') result(ierr)
       import :: c_ptr, c_int, c_double, c_char
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: name(*)
       integer(c_int), value :: slot_new, slot_sav
       real(c_double), value :: blend
       integer(c_int) :: ierr
     end function
     function rsg_scb_run(h, p, pressure, user, res) bind(C, name='rsg_scb_run') result(ierr)
       import :: c_ptr, c_int, c_funptr, rsg_scb_run_params, rsg_scb_run_result
       type(c_ptr), value :: h
       type(rsg_scb_run_params), intent(in) :: p
       type(c_funptr), value :: pressure                 ! rsg_scb_pressure_fn
       type(c_ptr), value :: user
       type(rsg_scb_run_result), intent(out) :: res
       integer(c_int) :: ierr
     end function
     function rsg_scb_min_jacobian(h, minjac) bind(C, name='rsg_scb_min_jacobian') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(out) :: minjac
       integer(c_int) :: ierr
     end function
     function rsg_hI_integrals(device, nthe, nR, nT, nPa, nThetaEquator, bnormal, chiVal, mu, xRAM, yRAM, zRAM, bRAM, &
          density, outsideMGNP, I_cart, H_cart, HDens_cart, bZEq_cart, ms) bind(C, name='rsg_hI_integrals') result(ierr)
       ! the integral block of computehI, src/ModRamScb.f90:372-410
       import :: c_int, c_double
       integer(c_int), value :: device, nthe, nR, nT, nPa, nThetaEquator
       real(c_double), value :: bnormal
       real(c_double), intent(in) :: chiVal(*), mu(*), xRAM(*), yRAM(*), zRAM(*), bRAM(*), density(*)
       integer(c_int), intent(in) :: outsideMGNP(*)
       real(c_double), intent(inout) :: I_cart(*), H_cart(*), HDens_cart(*), bZEq_cart(*)
       real(c_double), intent(out) :: ms
       integer(c_int) :: ierr
     end function
     function rsg_hI_tail(device, nR, nT, nPa, I_cart, H_cart, HDens_cart, bZEq_cart, ScaleAt, outsideMGNP, Lz, PA, PAbn, &
          integral_smooth, DthI, FNHS, FNIS, BOUNHS, BOUNIS, HDNS, BNES, dIdt, dHdt, dIbndt, dBdt, gslerr, ms) &
          bind(C, name='rsg_hI_tail') result(ierr)
       ! computehI after the integral block, src/ModRamScb.f90:413-637
       import :: c_int, c_double
       integer(c_int), value :: device, nR, nT, nPa, integral_smooth
       real(c_double), value :: DthI
       real(c_double), intent(inout) :: I_cart(*), H_cart(*), HDens_cart(*), bZEq_cart(*)
       integer(c_int), intent(in) :: ScaleAt(*), outsideMGNP(*)
       real(c_double), intent(in) :: Lz(*), PA(*), PAbn(*)
       real(c_double), intent(inout) :: FNHS(*), FNIS(*), BOUNHS(*), BOUNIS(*), HDNS(*), BNES(*)
       real(c_double), intent(inout) :: dIdt(*), dHdt(*), dIbndt(*), dBdt(*)
       integer(c_int), intent(out) :: gslerr
       real(c_double), intent(out) :: ms
       integer(c_int) :: ierr
     end function
     function rsg_hI_convert_lines(device, nthe, npsi, nzeta, nR, nT, nThetaEquator, x, y, z, bf, psi, alfa, Lz, MLT, &
          xRAM, yRAM, zRAM, bRAM, outsideSCB, ms) bind(C, name='rsg_hI_convert_lines') result(ierr)
       ! computehI, "Convert SCB field lines to RAM field lines", src/ModRamScb.f90:252-300
       import :: c_int, c_double
       integer(c_int), value :: device, nthe, npsi, nzeta, nR, nT, nThetaEquator
       real(c_double), intent(in) :: x(*), y(*), z(*), bf(*), psi(*), alfa(*), Lz(*), MLT(*)
       real(c_double), intent(inout) :: xRAM(*), yRAM(*), zRAM(*), bRAM(*)
       integer(c_int), intent(inout) :: outsideSCB(*)
       real(c_double), intent(out) :: ms
       integer(c_int) :: ierr
     end function
     function rsg_hi_create(hOut, device, nthe, npsi, nzeta, nR, nT, nPa, nThetaEquator, bnormal, chiVal, mu, Lz, MLT, PA, PAbn) &
          bind(C, name='rsg_hi_create') result(ierr)
       ! the resident computehI (src/ModRamScb.f90:249-637 as one device object)
       import :: c_ptr, c_int, c_double
       type(c_ptr), intent(out) :: hOut
       integer(c_int), value :: device, nthe, npsi, nzeta, nR, nT, nPa, nThetaEquator
       real(c_double), value :: bnormal
       real(c_double), intent(in) :: chiVal(*), mu(*), Lz(*), MLT(*), PA(*), PAbn(*)
       integer(c_int) :: ierr
     end function
     function rsg_hi_set_ram_fields(h, FNHS, FNIS, BOUNHS, BOUNIS, HDNS, BNES, HDens_cart) &
          bind(C, name='rsg_hi_set_ram_fields') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: FNHS(*), FNIS(*), BOUNHS(*), BOUNIS(*), HDNS(*), BNES(*)
       type(c_ptr), value :: HDens_cart                        ! c_loc(HDens_cart) or c_null_ptr
       integer(c_int) :: ierr
     end function
     function rsg_hi_convert(h, x, y, z, bf, psi, alfa, scb, outsideSCB, nOutside) bind(C, name='rsg_hi_convert') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h, x, y, z, bf, psi, alfa, scb    ! six c_loc(array) and c_null_ptr, or six c_null_ptr and hScb
       integer(c_int), intent(inout) :: outsideSCB(*)
       integer(c_int), intent(out) :: nOutside
       integer(c_int) :: ierr
     end function
     function rsg_hi_set_line(h, i, j, xl, yl, zl, bl) bind(C, name='rsg_hi_set_line') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: i, j
       real(c_double), intent(in) :: xl(*), yl(*), zl(*), bl(*)
       integer(c_int) :: ierr
     end function
     function rsg_hi_finish(h, ScaleAt, outsideMGNP, density, integral_smooth, DthI, gslerr) bind(C, name='rsg_hi_finish') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h, ScaleAt, outsideMGNP, density  ! c_loc(...) or c_null_ptr: device-side defaults
       integer(c_int), value :: integral_smooth
       real(c_double), value :: DthI
       integer(c_int), intent(out) :: gslerr
       integer(c_int) :: ierr
     end function
     function rsg_hi_get(h, name, dst) bind(C, name='rsg_hi_get') result(ierr)
       import :: c_ptr, c_int, c_double, c_char
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: name(*)
       real(c_double), intent(inout) :: dst(*)
       integer(c_int) :: ierr
     end function
     function rsg_hi_get_int(h, which, dst) bind(C, name='rsg_hi_get_int') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       integer(c_int), value :: which
       integer(c_int), intent(inout) :: dst(*)
       integer(c_int) :: ierr
     end function
     function rsg_hi_device_fields(h, ptrs9, outsideMGNP) bind(C, name='rsg_hi_device_fields') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h
       type(c_ptr), value :: ptrs9, outsideMGNP                ! c_loc of a type(c_ptr) array(9) / of a type(c_ptr) scalar
       integer(c_int) :: ierr
     end function
     function rsg_ram_set_fields_device(h, ptrs9, d_outsideMGNP) bind(C, name='rsg_ram_set_fields_device') result(ierr)
       import :: c_ptr, c_int
       type(c_ptr), value :: h, ptrs9, d_outsideMGNP
       integer(c_int) :: ierr
     end function
  end interface

contains

  subroutine rsg_scb_check(ierr, where)
    ! a failure of the device path itself is fatal (CON_stop, src/Main.f90:135-168);
    ! numerical failures come back through the sorfail out-parameters and set SORFail,
    ! which scb_run handles by rolling back (src/ModScbRun.f90:226-229, 398-414)
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: where
    if (ierr /= 0) call CON_stop('ramscb_gpu (SCB) failure in '//where)
  end subroutine rsg_scb_check

  subroutine scb_gpu_upload_domain
    ! after computational_domain / Update_Domain (src/ModScbIO.f90:179-205, 211-519) and
    ! whenever the host has moved x, y, z, alfa or psi
    use ModScbVariables, ONLY: x, y, z, alfa, psi, thetaVal, rhoVal, zetaVal, f, fzet, alphaVal, psiVal, chiVal
    call rsg_scb_check(rsg_scb_set_grid(hScb, thetaVal, rhoVal, zetaVal, f, fzet), 'upload_domain')
    call rsg_scb_check(rsg_scb_set_geometry(hScb, x, y, z), 'upload_domain')
    call rsg_scb_check(rsg_scb_set_field(hScb, 'alfa'//c_null_char, alfa), 'upload_domain')
    call rsg_scb_check(rsg_scb_set_field(hScb, 'psi'//c_null_char, psi), 'upload_domain')
    call rsg_scb_check(rsg_scb_set_map_targets(hScb, alphaVal, psiVal, chiVal), 'upload_domain')
  end subroutine scb_gpu_upload_domain

  subroutine computehI_integrals_gpu(xRAM, yRAM, zRAM, bRAM, density, I_cart, H_cart, HDens_cart, bZEq_Cart)
    ! replaces the "BEGIN INTEGRAl CALCULATION" loop nest of computehI (src/ModRamScb.f90:372-410): call it with the
    ! routine's local arrays in place of the !$OMP PARALLEL DO over (i, j); the scaling / smoothing tail stays as it is
    use ModRamGrids,     ONLY: nR, nT, nPa
    use ModRamVariables, ONLY: MU, outsideMGNP
    use ModScbGrids,     ONLY: nthe
    use ModScbVariables, ONLY: chiVal, nThetaEquator, bnormal
    real(c_double), intent(in)    :: xRAM(nthe,nR,nT), yRAM(nthe,nR,nT), zRAM(nthe,nR,nT), bRAM(nthe,nR,nT), density(nthe,nR,nT)
    real(c_double), intent(inout) :: I_cart(nR,nT,nPa), H_cart(nR,nT,nPa), HDens_cart(nR,nT,nPa), bZEq_Cart(nR,nT)
    real(c_double) :: ms
    call rsg_scb_check(rsg_hI_integrals(0_c_int, int(nthe,c_int), int(nR,c_int), int(nT,c_int), int(nPa,c_int), &
         int(nThetaEquator,c_int), bnormal, chiVal, MU, xRAM, yRAM, zRAM, bRAM, density, outsideMGNP, &
         I_cart, H_cart, HDens_cart, bZEq_Cart, ms), 'computehI_integrals')
  end subroutine computehI_integrals_gpu

  subroutine computehI_tail_gpu(I_cart, H_cart, HDens_cart, bZEq_Cart, ScaleAt, DthI)
    ! replaces computehI from "Scale based on outer SCB boundary" to the NaN check (src/ModRamScb.f90:413-637);
    ! the caller keeps EIR(1,:) = EIP(1,:) = 0 (:611-612) and TOld = TimeRamElapsed (:623)
    use ModRamGrids,     ONLY: nR, nT, nPa
    use ModRamParams,    ONLY: integral_smooth
    use ModRamVariables, ONLY: FNHS, FNIS, BOUNHS, BOUNIS, HDNS, BNES, dIdt, dHdt, dIbndt, dBdt, LZ, PA, PAbn, outsideMGNP
    real(c_double), intent(inout) :: I_cart(nR,nT,nPa), H_cart(nR,nT,nPa), HDens_cart(nR,nT,nPa), bZEq_Cart(nR,nT)
    integer(c_int), intent(in)    :: ScaleAt(nT)
    real(c_double), intent(in)    :: DthI
    integer(c_int) :: gslerr, ismooth
    real(c_double) :: ms
    ismooth = 0
    if (integral_smooth) ismooth = 1
    call rsg_scb_check(rsg_hI_tail(0_c_int, int(nR,c_int), int(nT,c_int), int(nPa,c_int), I_cart, H_cart, HDens_cart, &
         bZEq_Cart, ScaleAt, outsideMGNP, LZ, PA, PAbn, ismooth, DthI, FNHS, FNIS, BOUNHS, BOUNIS, HDNS, BNES, &
         dIdt, dHdt, dIbndt, dBdt, gslerr, ms), 'computehI_tail')
    if (gslerr /= 0) call CON_stop('computehI_tail_gpu: GSL_Interpolation_1D failed on a pitch-angle line')
  end subroutine computehI_tail_gpu

  subroutine computehI_convert_lines_gpu(xRAM, yRAM, zRAM, bRAM, outsideSCB)
    ! replaces the first !$OMP PARALLEL DO of computehI's default branch (src/ModRamScb.f90:252-300)
    use ModRamGrids,     ONLY: nR, nT
    use ModRamVariables, ONLY: LZ, MLT
    use ModScbGrids,     ONLY: nthe, npsi, nzeta
    use ModScbVariables, ONLY: x, y, z, bf, psi, alfa, nThetaEquator
    real(c_double), intent(inout) :: xRAM(nthe,nR,nT), yRAM(nthe,nR,nT), zRAM(nthe,nR,nT), bRAM(nthe,nR,nT)
    integer(c_int), intent(inout) :: outsideSCB(nR,nT)
    real(c_double) :: ms
    call rsg_scb_check(rsg_hI_convert_lines(0_c_int, int(nthe,c_int), int(npsi,c_int), int(nzeta,c_int), int(nR,c_int), &
         int(nT,c_int), int(nThetaEquator,c_int), x, y, z, bf, psi, alfa, LZ, MLT, xRAM, yRAM, zRAM, bRAM, outsideSCB, ms), &
         'computehI_convert_lines')
  end subroutine computehI_convert_lines_gpu

  subroutine computehI_gpu(hRam, DthI)
    ! replaces computehI's default branch from "Convert SCB field lines to RAM field lines" to the NaN check
    ! (src/ModRamScb.f90:249-637) with everything resident on the device: SCB arrays from the host, the RAM field arrays
    ! kept in the rsg_hi object between calls and handed device-to-device to the RAM state hRam.  Points outside the SCB
    ! domain that the reference traces with Geopack (:330-362) are traced here by the host exactly as before and uploaded
    ! line by line; magnetopause logic (ScaleAt / outsideMGNP, :306-329) stays on the host.
    use ModRamGrids,     ONLY: nR, nT, nPa
    use ModRamParams,    ONLY: integral_smooth
    use ModRamVariables, ONLY: FNHS, FNIS, BOUNHS, BOUNIS, HDNS, BNES, dIdt, dHdt, dIbndt, dBdt, LZ, MLT, MU, PA, PAbn, outsideMGNP
    use ModScbGrids,     ONLY: nthe, npsi, nzeta
    use ModScbVariables, ONLY: x, y, z, bf, psi, alfa, chiVal, nThetaEquator, bnormal
    type(c_ptr), intent(in)    :: hRam
    real(c_double), intent(in) :: DthI
    type(c_ptr), save :: hHi = c_null_ptr
    type(c_ptr), target :: ptrs(9), pOut
    integer(c_int), target :: ScaleAt(nT), outsideSCB(nR,nT)
    integer(c_int) :: nOutside, gslerr, ismooth, i, j
    if (.not. c_associated(hHi)) then
       call rsg_scb_check(rsg_hi_create(hHi, 0_c_int, int(nthe,c_int), int(npsi,c_int), int(nzeta,c_int), int(nR,c_int), &
            int(nT,c_int), int(nPa,c_int), int(nThetaEquator,c_int), bnormal, chiVal, MU, LZ, MLT, PA, PAbn), 'computehI')
       call rsg_scb_check(rsg_hi_set_ram_fields(hHi, FNHS, FNIS, BOUNHS, BOUNIS, HDNS, BNES, c_null_ptr), 'computehI')
    end if
    call rsg_scb_check(rsg_hi_convert(hHi, c_loc(x), c_loc(y), c_loc(z), c_loc(bf), c_loc(psi), c_loc(alfa), c_null_ptr, &
         outsideSCB, nOutside), 'computehI')
    ScaleAt = 0
    outsideMGNP = 0
    do j = 1, nT
       do i = 1, nR
          if (outsideSCB(i,j) == 1) then
             if (ScaleAt(j) == 0) ScaleAt(j) = i
             ! the reference's magnetopause test and Geopack trace (:311-362) go here unchanged; a traced line is handed
             ! over with rsg_hi_set_line(hHi, i, j, xRAM(:,i,j), yRAM(:,i,j), zRAM(:,i,j), bRAM(:,i,j)); without a tracer:
             outsideMGNP(i,j) = 1
          end if
       end do
    end do
    ismooth = 0
    if (integral_smooth) ismooth = 1
    call rsg_scb_check(rsg_hi_finish(hHi, c_loc(ScaleAt), c_loc(outsideMGNP), c_null_ptr, ismooth, DthI, gslerr), 'computehI')
    if (gslerr /= 0) call CON_stop('computehI_gpu: GSL_Interpolation_1D failed on a pitch-angle line')
    ! the new field arrays: device-to-device into the RAM state, and to the host copies the rest of the code reads
    call rsg_scb_check(rsg_hi_device_fields(hHi, c_loc(ptrs), c_loc(pOut)), 'computehI')
    call rsg_scb_check(rsg_ram_set_fields_device(hRam, c_loc(ptrs), pOut), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'FNHS'//c_null_char, FNHS), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'FNIS'//c_null_char, FNIS), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'BOUNHS'//c_null_char, BOUNHS), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'BOUNIS'//c_null_char, BOUNIS), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'HDNS'//c_null_char, HDNS), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'BNES'//c_null_char, BNES), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'dIdt'//c_null_char, dIdt), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'dHdt'//c_null_char, dHdt), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'dIbndt'//c_null_char, dIbndt), 'computehI')
    call rsg_scb_check(rsg_hi_get(hHi, 'dBdt'//c_null_char, dBdt), 'computehI')
  end subroutine computehI_gpu

end module ModScbGpu
