MODULE traadv_fct
   !!==============================================================================
   !!                       ***  MODULE  traadv_fct  ***   (B200 device path)
   !! Ocean tracers: horizontal & vertical advective trend, 2nd/4th order FCT scheme
   !!==============================================================================
   !! Drop-in replacement of src/OCE/TRA/traadv_fct.F90: same module name, same PUBLIC
   !! routines and the exact argument lists
   !!     tra_adv_fct( kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, ptn, pta, kjpt, kn_fct_h, kn_fct_v )
   !!     interp_4th_cpt( pt_in, pt_out )
   !! so traadv.F90:150, trcadv.F90:127 and traadv_cen.F90:158 compile unchanged and the scheme is
   !! still selected by namelist (ln_traadv_fct / ln_trcadv_fct -> nadv = np_FCT).
   !! All arithmetic runs in libnemo_fct.so (hand-written fp64 CUDA for sm_100a) through the C ABI of
   !! include/nemo_fct.h; this module only binds it.  There is NO CPU fallback: an error from the
   !! library is turned into CALL ctl_stop('STOP', ...), as the reference does for fatal errors
   !! (lib_mpp.F90:1868-1907).
   !!
   !! NOTE: the build container has no Fortran compiler, so this file is compiled and linked only on
   !! the NEMO side (see INTEGRATION.md); the same entry points are exercised by the test-suite
   !! through ctypes with Fortran-layout buffers.
   !!----------------------------------------------------------------------
   USE, INTRINSIC :: ISO_C_BINDING
   USE oce            ! ocean dynamics and active tracers
   USE dom_oce        ! ocean space and time domain (masks, e3t_b/n/a, e1e2t, mikt, mbkt, decomposition scalars)
   USE trc_oce        ! share passive tracers/Ocean variables
   USE trd_oce        ! trends: ocean variables (l_trdtra, l_trdtrc)
   USE diaptr  , ONLY : ln_diaptr
   USE in_out_manager ! I/O manager (lwp, numout)
   USE iom     , ONLY : iom_use
   USE lib_mpp        ! ctl_stop, mpi_comm_oce
   USE lbcnfd  , ONLY : isendto, nsndto   ! no-gather fold partners (lbcnfd.F90:53-55), part of the domain descriptor

   IMPLICIT NONE
#if defined key_mpp_mpi
   ! lib_mpp is PRIVATE and includes mpif.h inside the module (lib_mpp.F90:63,116): only mpi_comm_oce is exported, so the MPI
   ! constants and MPI_BCAST used by tra_adv_fct_gpu_init have to come from here
   INCLUDE 'mpif.h'
#endif
   PRIVATE

   PUBLIC   tra_adv_fct        ! called by traadv.F90 and trcadv.F90
   PUBLIC   interp_4th_cpt     ! called by traadv_cen.F90
   PUBLIC   tra_adv_fct_gpu_init   ! called once from nemo_init, after dom_init (nemogcm.F90:417)
   PUBLIC   nhandle, gpu_stop      ! shared with the other device-path modules (traadv_mus_gpu.F90)

   !                                     ! struct nemo_fct_domain (include/nemo_fct.h), same field order
   TYPE, BIND(C) ::   nemo_fct_domain
      INTEGER(C_INT) ::   jpiglo, jpjglo, jpk, jperio, jpni, jpnj, narea, jpi, jpj, jpimax, jpjmax
      INTEGER(C_INT) ::   nimpp, njmpp, nlci, nlcj, nldi, nlei, nldj, nlej, nbondi, nbondj
      INTEGER(C_INT) ::   noea, nowe, noso, nono, npolj, l_Iperio, l_Jperio, nsndto
      INTEGER(C_INT) ::   isendto(3)
      INTEGER(C_INT) ::   key_mpp_mpi
   END TYPE nemo_fct_domain

   TYPE(C_PTR), SAVE ::   nhandle = C_NULL_PTR   ! nemo_fct_handle of this MPI rank

   INTERFACE
      INTEGER(C_INT) FUNCTION nemo_fct_create( dom, device, handle ) BIND(C, NAME='nemo_fct_create')
         IMPORT :: C_INT, C_PTR, nemo_fct_domain
         TYPE(nemo_fct_domain), INTENT(in   ) ::   dom
         INTEGER(C_INT), VALUE                ::   device
         TYPE(C_PTR)          , INTENT(  out) ::   handle
      END FUNCTION nemo_fct_create
      INTEGER(C_INT) FUNCTION nemo_fct_set_domain_arrays( h, tmask, umask, vmask, wmask, e1e2t, r1_e1e2t, mikt, mbkt,   &
         &                                               ln_linssh, ln_isfcav ) BIND(C, NAME='nemo_fct_set_domain_arrays')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE ::   h
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in) ::   tmask, umask, vmask, wmask, e1e2t, r1_e1e2t
         INTEGER(C_INT), DIMENSION(*), INTENT(in) ::   mikt, mbkt
         INTEGER(C_INT), VALUE ::   ln_linssh, ln_isfcav
      END FUNCTION nemo_fct_set_domain_arrays
      INTEGER(C_INT) FUNCTION nemo_fct_set_e3t( h, e3t_b, e3t_n, e3t_a, is_device ) BIND(C, NAME='nemo_fct_set_e3t')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE ::   h
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in) ::   e3t_b, e3t_n, e3t_a
         INTEGER(C_INT), VALUE ::   is_device
      END FUNCTION nemo_fct_set_e3t
      INTEGER(C_INT) FUNCTION nemo_fct_comm_unique_id( id128 ) BIND(C, NAME='nemo_fct_comm_unique_id')
         IMPORT :: C_INT, C_CHAR
         CHARACTER(KIND=C_CHAR), DIMENSION(128), INTENT(out) ::   id128
      END FUNCTION nemo_fct_comm_unique_id
      INTEGER(C_INT) FUNCTION nemo_fct_comm_init( h, id128, nranks, rank ) BIND(C, NAME='nemo_fct_comm_init')
         IMPORT :: C_INT, C_PTR, C_CHAR
         TYPE(C_PTR), VALUE ::   h
         CHARACTER(KIND=C_CHAR), DIMENSION(128), INTENT(in) ::   id128
         INTEGER(C_INT), VALUE ::   nranks, rank
      END FUNCTION nemo_fct_comm_init
      INTEGER(C_INT) FUNCTION nemo_tra_adv_fct( h, kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, ptn, pta,   &
         &                                     kjpt, kn_fct_h, kn_fct_v ) BIND(C, NAME='nemo_tra_adv_fct')
         IMPORT :: C_INT, C_PTR, C_DOUBLE, C_CHAR
         TYPE(C_PTR), VALUE ::   h
         INTEGER(C_INT), VALUE ::   kt, kit000, kjpt, kn_fct_h, kn_fct_v
         CHARACTER(KIND=C_CHAR), DIMENSION(*), INTENT(in) ::   cdtype
         REAL(C_DOUBLE), VALUE ::   p2dt
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in   ) ::   pun, pvn, pwn, ptb, ptn
         REAL(C_DOUBLE), DIMENSION(*), INTENT(inout) ::   pta
      END FUNCTION nemo_tra_adv_fct
      INTEGER(C_INT) FUNCTION nemo_interp_4th_cpt( h, pt_in, pt_out ) BIND(C, NAME='nemo_interp_4th_cpt')
         IMPORT :: C_INT, C_PTR, C_DOUBLE
         TYPE(C_PTR), VALUE ::   h
         REAL(C_DOUBLE), DIMENSION(*), INTENT(in   ) ::   pt_in
         REAL(C_DOUBLE), DIMENSION(*), INTENT(inout) ::   pt_out
      END FUNCTION nemo_interp_4th_cpt
      FUNCTION nemo_fct_last_error() BIND(C, NAME='nemo_fct_last_error')
         IMPORT :: C_PTR
         TYPE(C_PTR) ::   nemo_fct_last_error
      END FUNCTION nemo_fct_last_error
   END INTERFACE

   !!----------------------------------------------------------------------
   !! NEMO/OCE 4.0 interface layer -- see INTEGRATION.md
   !!----------------------------------------------------------------------
CONTAINS

   SUBROUTINE gpu_stop( cdroutine )
      !!----------------------------------------------------------------------
      !! non-zero return of the C ABI  ==>  ctl_stop('STOP', ...) (lib_mpp.F90:1868)
      !!----------------------------------------------------------------------
      CHARACTER(len=*), INTENT(in) ::   cdroutine
      CHARACTER(KIND=C_CHAR), DIMENSION(:), POINTER ::   zmsg
      CHARACTER(len=512) ::   clmsg
      INTEGER ::   ji
      CALL C_F_POINTER( nemo_fct_last_error(), zmsg, (/ 512 /) )
      clmsg = ' '
      DO ji = 1, 512
         IF( zmsg(ji) == C_NULL_CHAR )   EXIT
         clmsg(ji:ji) = zmsg(ji)
      END DO
      CALL ctl_stop( 'STOP', TRIM(cdroutine)//' (libnemo_fct): '//TRIM(clmsg) )
   END SUBROUTINE gpu_stop


   SUBROUTINE tra_adv_fct_gpu_init
      !!----------------------------------------------------------------------
      !!                  ***  ROUTINE tra_adv_fct_gpu_init  ***
      !! ** Purpose :   create the device context ONCE (as nemo_alloc does for the module arrays,
      !!                nemogcm.F90:640-673), upload the time-invariant dom_oce arrays, open the NCCL
      !!                communicator over the same ranks as mpi_comm_oce.
      !!----------------------------------------------------------------------
      TYPE(nemo_fct_domain) ::   ydom
      CHARACTER(KIND=C_CHAR), DIMENSION(128) ::   clid
      INTEGER ::   ierr
      !!----------------------------------------------------------------------
      ydom%jpiglo = jpiglo   ;   ydom%jpjglo = jpjglo   ;   ydom%jpk = jpk   ;   ydom%jperio = jperio
      ydom%jpni   = jpni     ;   ydom%jpnj   = jpnj     ;   ydom%narea = narea
      ydom%jpi    = jpi      ;   ydom%jpj    = jpj      ;   ydom%jpimax = jpimax   ;   ydom%jpjmax = jpjmax
      ydom%nimpp  = nimpp    ;   ydom%njmpp  = njmpp
      ydom%nlci   = nlci     ;   ydom%nlcj   = nlcj
      ydom%nldi   = nldi     ;   ydom%nlei   = nlei     ;   ydom%nldj = nldj   ;   ydom%nlej = nlej
      ydom%nbondi = nbondi   ;   ydom%nbondj = nbondj
      ydom%noea   = noea     ;   ydom%nowe   = nowe     ;   ydom%noso = noso   ;   ydom%nono = nono
      ydom%npolj  = npolj
      ydom%l_Iperio = MERGE( 1, 0, l_Iperio )   ;   ydom%l_Jperio = MERGE( 1, 0, l_Jperio )
      ydom%nsndto = nsndto   ;   ydom%isendto(1:3) = isendto(1:3)
#if defined key_mpp_mpi
      ydom%key_mpp_mpi = 1
#else
      ydom%key_mpp_mpi = 0
#endif
      ! device = -1 : the library takes the node-local rank from the launcher's environment (LOCAL_RANK, SLURM_LOCALID,
      ! OMPI_COMM_WORLD_LOCAL_RANK, MV2_COMM_WORLD_LOCAL_RANK, MPI_LOCALRANKID: one MPI rank per GPU) and refuses to guess
      IF( nemo_fct_create( ydom, -1_C_INT, nhandle ) /= 0 )   CALL gpu_stop( 'tra_adv_fct_gpu_init' )
      IF( nemo_fct_set_domain_arrays( nhandle, tmask, umask, vmask, wmask, e1e2t, r1_e1e2t, mikt, mbkt,   &
         &                            MERGE( 1, 0, ln_linssh ), MERGE( 1, 0, ln_isfcav ) ) /= 0 )   CALL gpu_stop( 'tra_adv_fct_gpu_init' )
#if defined key_mpp_mpi
      IF( jpnij > 1 ) THEN          ! NCCL bootstrap: rank 0 creates the id, MPI broadcasts the 128 bytes
         IF( narea == 1 ) THEN
            IF( nemo_fct_comm_unique_id( clid ) /= 0 )   CALL gpu_stop( 'tra_adv_fct_gpu_init' )
         ENDIF
         CALL MPI_BCAST( clid, 128, MPI_CHARACTER, 0, mpi_comm_oce, ierr )
         IF( nemo_fct_comm_init( nhandle, clid, jpnij, narea-1 ) /= 0 )   CALL gpu_stop( 'tra_adv_fct_gpu_init' )
      ENDIF
#endif
   END SUBROUTINE tra_adv_fct_gpu_init


   SUBROUTINE tra_adv_fct( kt, kit000, cdtype, p2dt, pun, pvn, pwn,       &
      &                                              ptb, ptn, pta, kjpt, kn_fct_h, kn_fct_v )
      !!----------------------------------------------------------------------
      !!                  ***  ROUTINE tra_adv_fct  ***
      !! Same interface as the reference routine (traadv_fct.F90:54-80).
      !!----------------------------------------------------------------------
      INTEGER                              , INTENT(in   ) ::   kt              ! ocean time-step index
      INTEGER                              , INTENT(in   ) ::   kit000          ! first time step index
      CHARACTER(len=3)                     , INTENT(in   ) ::   cdtype          ! =TRA or TRC (tracer indicator)
      INTEGER                              , INTENT(in   ) ::   kjpt            ! number of tracers
      INTEGER                              , INTENT(in   ) ::   kn_fct_h        ! order of the FCT scheme (=2 or 4)
      INTEGER                              , INTENT(in   ) ::   kn_fct_v        ! order of the FCT scheme (=2 or 4)
      REAL(wp)                             , INTENT(in   ) ::   p2dt            ! tracer time-step
      REAL(wp), DIMENSION(jpi,jpj,jpk     ), INTENT(in   ) ::   pun, pvn, pwn   ! 3 ocean transport components
      REAL(wp), DIMENSION(jpi,jpj,jpk,kjpt), INTENT(in   ) ::   ptb, ptn        ! before and now tracer fields
      REAL(wp), DIMENSION(jpi,jpj,jpk,kjpt), INTENT(inout) ::   pta             ! tracer trend
      !
      CHARACTER(KIND=C_CHAR), DIMENSION(4) ::   cltype
      INTEGER ::   ji
      !!----------------------------------------------------------------------
      IF( kt == kit000 )  THEN
         IF(lwp) WRITE(numout,*)
         IF(lwp) WRITE(numout,*) 'tra_adv_fct : FCT advection scheme on ', cdtype, ' (B200 device path, libnemo_fct)'
         IF(lwp) WRITE(numout,*) '~~~~~~~~~~~'
      ENDIF
      ! optional diagnostics of the reference (traadv_fct.F90:96-112, 172-176, 299-316) are not on the device path
      IF( ( cdtype == 'TRA' .AND. l_trdtra ) .OR. ( cdtype == 'TRC' .AND. l_trdtrc ) )   &
         &   CALL ctl_stop( 'STOP', 'tra_adv_fct (device path): trend diagnostics (l_trdtra/l_trdtrc) not supported' )
      IF( cdtype == 'TRA' .AND. ln_diaptr )   &
         &   CALL ctl_stop( 'STOP', 'tra_adv_fct (device path): poleward transport diagnostics (ln_diaptr) not supported' )
      IF( cdtype == 'TRA' .AND. ( iom_use("uadv_heattr") .OR. iom_use("vadv_heattr") .OR.   &
         &                        iom_use("uadv_salttr") .OR. iom_use("vadv_salttr") ) )   &
         &   CALL ctl_stop( 'STOP', 'tra_adv_fct (device path): heat/salt transport diagnostics not supported' )
      !
      DO ji = 1, 3   ;   cltype(ji) = cdtype(ji:ji)   ;   END DO
      cltype(4) = C_NULL_CHAR
      ! e3t_b/n/a change every step with the non-linear free surface (dom_vvl_sf_swp, domvvl.F90:620)
      IF( .NOT.ln_linssh .OR. kt == kit000 ) THEN
         IF( nemo_fct_set_e3t( nhandle, e3t_b, e3t_n, e3t_a, 0_C_INT ) /= 0 )   CALL gpu_stop( 'tra_adv_fct' )
      ENDIF
      IF( nemo_tra_adv_fct( nhandle, kt, kit000, cltype, p2dt, pun, pvn, pwn, ptb, ptn, pta,   &
         &                  kjpt, kn_fct_h, kn_fct_v ) /= 0 )   CALL gpu_stop( 'tra_adv_fct' )
      !
   END SUBROUTINE tra_adv_fct


   SUBROUTINE interp_4th_cpt( pt_in, pt_out )
      !!----------------------------------------------------------------------
      !!                  ***  ROUTINE interp_4th_cpt  ***
      !! Same interface as the reference routine (traadv_fct.F90:517-527).
      !!----------------------------------------------------------------------
      REAL(wp),DIMENSION(jpi,jpj,jpk), INTENT(in   ) ::   pt_in    ! field at t-point
      REAL(wp),DIMENSION(jpi,jpj,jpk), INTENT(  out) ::   pt_out   ! field interpolated at w-point
      !!----------------------------------------------------------------------
      IF( nemo_interp_4th_cpt( nhandle, pt_in, pt_out ) /= 0 )   CALL gpu_stop( 'interp_4th_cpt' )
   END SUBROUTINE interp_4th_cpt

   !!======================================================================
END MODULE traadv_fct
