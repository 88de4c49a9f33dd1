MODULE traadv_mus
   !!======================================================================
   !!                       ***  MODULE  traadv_mus  ***   (B200 device path)
   !! Ocean  tracers:  horizontal & vertical advective trend, MUSCL scheme
   !!======================================================================
   !! Drop-in replacement of src/OCE/TRA/traadv_mus.F90: same module name, same PUBLIC routine and argument list
   !!     tra_adv_mus( kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, pta, kjpt, ld_msc_ups )
   !! so traadv.F90:152 and trcadv.F90:129 compile unchanged (ln_traadv_mus / ln_trcadv_mus -> nadv = np_MUS; it is
   !! the scheme tests/BENCH and ORCA2_ICE_PISCES select for the passive tracers, namelist_top_cfg:79).
   !! The arithmetic runs in libnemo_fct.so through the C ABI of include/nemo_fct.h; the device context is the one
   !! created by tra_adv_fct_gpu_init (MODULE traadv_fct, shim/traadv_fct_gpu.F90).  NO CPU fallback.
   !! Compiled on the NEMO side only (no Fortran compiler in the build container, see INTEGRATION.md).
   !!----------------------------------------------------------------------
   USE, INTRINSIC :: ISO_C_BINDING
   USE oce            ! ocean dynamics and active tracers
   USE dom_oce        ! r1_e1e2u, r1_e1e2v, e3u_n, e3v_n, e3w_n, ln_linssh
   USE trc_oce        ! share passive tracers/Ocean variables
   USE trd_oce        ! trends: ocean variables (l_trdtra, l_trdtrc)
   USE sbcrnf  , ONLY : rnfmsk, rnfmsk_z   ! river mouth masks of the upstream indicator (traadv_mus.F90:108-111)
   USE diaptr  , ONLY : ln_diaptr
   USE in_out_manager ! I/O manager (lwp, numout)
   USE iom     , ONLY : iom_use
   USE lib_mpp        ! ctl_stop
   USE traadv_fct, ONLY : nhandle, gpu_stop

   IMPLICIT NONE
   PRIVATE

   PUBLIC   tra_adv_mus   ! routine called by traadv.F90 and trcadv.F90

   INTERFACE
      INTEGER(C_INT) FUNCTION nemo_fct_set_mus_metrics( h, r1_e1e2u, r1_e1e2v ) BIND(C, NAME='nemo_fct_set_mus_metrics')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE ::   h
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in) ::   r1_e1e2u, r1_e1e2v
      END FUNCTION nemo_fct_set_mus_metrics
      INTEGER(C_INT) FUNCTION nemo_fct_set_e3uvw( h, e3u_n, e3v_n, e3w_n, is_device ) BIND(C, NAME='nemo_fct_set_e3uvw')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE ::   h
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in) ::   e3u_n, e3v_n, e3w_n
         INTEGER(C_INT), VALUE ::   is_device
      END FUNCTION nemo_fct_set_e3uvw
      INTEGER(C_INT) FUNCTION nemo_fct_set_mus_upstream( h, ld_msc_ups, rnfmsk, rnfmsk_z ) BIND(C, NAME='nemo_fct_set_mus_upstream')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE ::   h
         INTEGER(C_INT), VALUE ::   ld_msc_ups
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in) ::   rnfmsk, rnfmsk_z
      END FUNCTION nemo_fct_set_mus_upstream
      INTEGER(C_INT) FUNCTION nemo_fct_set_e3t( h, e3t_b, e3t_n, e3t_a, is_device ) BIND(C, NAME='nemo_fct_set_e3t')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE ::   h
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in) ::   e3t_b, e3t_n, e3t_a
         INTEGER(C_INT), VALUE ::   is_device
      END FUNCTION nemo_fct_set_e3t
      INTEGER(C_INT) FUNCTION nemo_tra_adv_mus( h, kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, pta, kjpt )   &
         &                                     BIND(C, NAME='nemo_tra_adv_mus')
         IMPORT :: C_INT, C_PTR, C_DOUBLE, C_CHAR
         TYPE(C_PTR), VALUE ::   h
         INTEGER(C_INT), VALUE ::   kt, kit000, kjpt
         CHARACTER(KIND=C_CHAR), DIMENSION(*), INTENT(in) ::   cdtype
         REAL(C_DOUBLE), VALUE ::   p2dt
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in   ) ::   pun, pvn, pwn, ptb
         REAL(C_DOUBLE), DIMENSION(*), INTENT(inout) ::   pta
      END FUNCTION nemo_tra_adv_mus
   END INTERFACE

   LOGICAL, SAVE ::   l_first = .TRUE.   ! upstream indicator and metrics sent to the device at the first call (:98-113)

CONTAINS

   SUBROUTINE tra_adv_mus( kt, kit000, cdtype, p2dt, pun, pvn, pwn,             &
      &                                              ptb, pta, kjpt, ld_msc_ups )
      !!----------------------------------------------------------------------
      !!                    ***  ROUTINE tra_adv_mus  ***
      !! Same interface as the reference routine (traadv_mus.F90:55-79).
      !!----------------------------------------------------------------------
      INTEGER                              , INTENT(in   ) ::   kt              ! ocean time-step index
      INTEGER                              , INTENT(in   ) ::   kit000          ! first time step index
      CHARACTER(len=3)                     , INTENT(in   ) ::   cdtype          ! =TRA or TRC (tracer indicator)
      INTEGER                              , INTENT(in   ) ::   kjpt            ! number of tracers
      LOGICAL                              , INTENT(in   ) ::   ld_msc_ups      ! use upstream scheme within muscl
      REAL(wp)                             , INTENT(in   ) ::   p2dt            ! tracer time-step
      REAL(wp), DIMENSION(jpi,jpj,jpk     ), INTENT(in   ) ::   pun, pvn, pwn   ! 3 ocean velocity components
      REAL(wp), DIMENSION(jpi,jpj,jpk,kjpt), INTENT(in   ) ::   ptb             ! before tracer field
      REAL(wp), DIMENSION(jpi,jpj,jpk,kjpt), INTENT(inout) ::   pta             ! tracer trend
      !
      CHARACTER(KIND=C_CHAR), DIMENSION(4) ::   cltype
      INTEGER ::   ji
      !!----------------------------------------------------------------------
      IF( kt == kit000 )  THEN
         IF(lwp) WRITE(numout,*)
         IF(lwp) WRITE(numout,*) 'tra_adv : MUSCL advection scheme on ', cdtype, ' (B200 device path, libnemo_fct)'
         IF(lwp) WRITE(numout,*) '        : mixed up-stream           ', ld_msc_ups
         IF(lwp) WRITE(numout,*) '~~~~~~~'
      ENDIF
      IF( l_first ) THEN            ! what the reference allocates / defines at kt == kit000 (:98-113)
         IF( nemo_fct_set_mus_metrics( nhandle, r1_e1e2u, r1_e1e2v ) /= 0 )   CALL gpu_stop( 'tra_adv_mus' )
         IF( nemo_fct_set_mus_upstream( nhandle, MERGE( 1, 0, ld_msc_ups ), rnfmsk, rnfmsk_z ) /= 0 )   CALL gpu_stop( 'tra_adv_mus' )
         l_first = .FALSE.
      ENDIF
      ! diagnostics of the reference (:117-124, 207-214, 268) are not on the device path
      IF( ( cdtype == 'TRA' .AND. l_trdtra ) .OR. ( cdtype == 'TRC' .AND. l_trdtrc ) )   &
         &   CALL ctl_stop( 'STOP', 'tra_adv_mus (device path): trend diagnostics (l_trdtra/l_trdtrc) not supported' )
      IF( cdtype == 'TRA' .AND. ln_diaptr )   &
         &   CALL ctl_stop( 'STOP', 'tra_adv_mus (device path): poleward transport diagnostics (ln_diaptr) not supported' )
      IF( cdtype == 'TRA' .AND. ( iom_use("uadv_heattr") .OR. iom_use("vadv_heattr") .OR.   &
         &                        iom_use("uadv_salttr") .OR. iom_use("vadv_salttr") ) )   &
         &   CALL ctl_stop( 'STOP', 'tra_adv_mus (device path): heat/salt transport diagnostics not supported' )
      !
      DO ji = 1, 3   ;   cltype(ji) = cdtype(ji:ji)   ;   END DO
      cltype(4) = C_NULL_CHAR
      IF( .NOT.ln_linssh .OR. kt == kit000 ) THEN      ! thicknesses change every step with the non-linear free surface
         IF( nemo_fct_set_e3t  ( nhandle, e3t_b, e3t_n, e3t_a, 0_C_INT ) /= 0 )   CALL gpu_stop( 'tra_adv_mus' )
         IF( nemo_fct_set_e3uvw( nhandle, e3u_n, e3v_n, e3w_n, 0_C_INT ) /= 0 )   CALL gpu_stop( 'tra_adv_mus' )
      ENDIF
      IF( nemo_tra_adv_mus( nhandle, kt, kit000, cltype, p2dt, pun, pvn, pwn, ptb, pta, kjpt ) /= 0 )   CALL gpu_stop( 'tra_adv_mus' )
      !
   END SUBROUTINE tra_adv_mus

   !!======================================================================
END MODULE traadv_mus
