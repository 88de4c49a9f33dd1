! Stub modules for the stand-alone build of the reference's traadv_fct.F90 (oracle/_ref_recipe/README.md): the diagnostics the
! routine can hand its fluxes to.  All switches are off, so none of the stubbed routines is ever reached (traadv_fct.F90:96-112).
MODULE trdtra
   USE par_kind
   IMPLICIT NONE
   PUBLIC
CONTAINS
   SUBROUTINE trd_tra( kt, ctype, ktra, ktrd, ptrd, pun, ptra )
      INTEGER, INTENT(in) :: kt, ktra, ktrd
      CHARACTER(len=3), INTENT(in) :: ctype
      REAL(wp), DIMENSION(:,:,:), INTENT(in) :: ptrd
      REAL(wp), DIMENSION(:,:,:), INTENT(in), OPTIONAL :: pun, ptra
      STOP 'trd_tra stub reached'
   END SUBROUTINE trd_tra
END MODULE trdtra

MODULE diaptr
   USE par_kind
   IMPLICIT NONE
   PUBLIC
   LOGICAL :: ln_diaptr = .FALSE.
CONTAINS
   SUBROUTINE dia_ptr_hst( ktra, cptr, pva )
      INTEGER, INTENT(in) :: ktra
      CHARACTER(len=3), INTENT(in) :: cptr
      REAL(wp), DIMENSION(:,:,:), INTENT(in) :: pva
      STOP 'dia_ptr_hst stub reached'
   END SUBROUTINE dia_ptr_hst
END MODULE diaptr

MODULE diaar5
   USE par_kind
   IMPLICIT NONE
   PUBLIC
CONTAINS
   SUBROUTINE dia_ar5_hst( ktra, cptr, pua, pva )
      INTEGER, INTENT(in) :: ktra
      CHARACTER(len=3), INTENT(in) :: cptr
      REAL(wp), DIMENSION(:,:,:), INTENT(in) :: pua, pva
      STOP 'dia_ar5_hst stub reached'
   END SUBROUTINE dia_ar5_hst
END MODULE diaar5

MODULE iom
   IMPLICIT NONE
   PUBLIC
CONTAINS
   LOGICAL FUNCTION iom_use( cdname )
      CHARACTER(len=*), INTENT(in) :: cdname
      iom_use = .FALSE.
   END FUNCTION iom_use
END MODULE iom
