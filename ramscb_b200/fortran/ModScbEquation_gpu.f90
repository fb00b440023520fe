!============================================================================
! Replacement for src/ModScbEquation.f90: identical module name and (argument-less)
! public routines metrica (:18-280), metric (:283-540), newk (:546-604), newj (:607-665).
! Inputs (x, y, z; the outputs of computeBandJacob and of `pressure`) and outputs
! (vecd, vec1-4, vec6-9, vecx, vecr) live on the device; nothing crosses the bus.
! Shipped uncompiled, see ModScbGpu.f90.
!============================================================================
MODULE ModScbEquation

  use ModScbGpu
  use, intrinsic :: iso_c_binding

  implicit none

contains

  SUBROUTINE metrica
    call rsg_scb_check(rsg_scb_metrica(hScb), 'metrica')
  END SUBROUTINE metrica

  SUBROUTINE metric
    call rsg_scb_check(rsg_scb_metric(hScb), 'metric')
  END SUBROUTINE metric

  SUBROUTINE newk
    call rsg_scb_check(rsg_scb_newk(hScb), 'newk')
  END SUBROUTINE newk

  SUBROUTINE newj
    call rsg_scb_check(rsg_scb_newj(hScb), 'newj')
  END SUBROUTINE newj

END MODULE ModScbEquation
