! Stand-alone driver of the reference's tra_adv_fct (oracle/_ref_recipe/README.md).  Reads one binary record file written by
! pin_oracle.py (stream access, native endianness), runs ONE call of tra_adv_fct on a mono-domain (jpni = jpnj = 1, no MPI: the
! lateral boundary conditions go through the non-MPI lbc_lnk of LBC/lbclnk.F90) and writes pta back.
PROGRAM fct_ref_driver
   USE par_oce
   USE dom_oce
   USE oce
   USE trd_oce , ONLY : l_trdtra, l_trdtrc
   USE in_out_manager
   USE traadv_fct
   IMPLICIT NONE
   INTEGER :: ihdr(10), kjpt, kn_fct_h, kn_fct_v, ierr, iu
   REAL(wp) :: zp2dt
   REAL(wp), ALLOCATABLE, DIMENSION(:,:,:)   :: zun, zvn, zwn
   REAL(wp), ALLOCATABLE, DIMENSION(:,:,:,:) :: ztb, ztn, zta
   CHARACTER(len=256) :: cin, cout
   !
   CALL get_command_argument( 1, cin ) ;   CALL get_command_argument( 2, cout )
   OPEN( newunit=iu, file=TRIM(cin), access='stream', form='unformatted', status='old' )
   READ(iu) ihdr               ! jpi, jpj, jpk, jperio, kjpt, kn_fct_h, kn_fct_v, ln_linssh, ln_isfcav, spare
   READ(iu) zp2dt
   jpi = ihdr(1) ; jpj = ihdr(2) ; jpk = ihdr(3) ; jperio = ihdr(4) ; kjpt = ihdr(5) ; kn_fct_h = ihdr(6) ; kn_fct_v = ihdr(7)
   jpiglo = jpi ; jpjglo = jpj ; jpkglo = jpk ; jpim1 = jpi-1 ; jpjm1 = jpj-1 ; jpkm1 = jpk-1 ; jpij = jpi*jpj
   jpni = 1 ; jpnj = 1 ; jpnij = 1 ; narea = 1 ; jpimax = jpi ; jpjmax = jpj
   nlci = jpi ; nlcj = jpj ; nldi = 1 ; nlei = jpi ; nldj = 1 ; nlej = jpj ; nimpp = 1 ; njmpp = 1
   npolj = 0 ; IF( jperio >= 3 .AND. jperio <= 6 ) npolj = jperio          ! mppini.F90:86 (no key_mpp_mpi)
   l_Iperio = ( jperio == 1 .OR. jperio == 4 .OR. jperio == 6 .OR. jperio == 7 )      ! mppini.F90:345-346 with jpni = jpnj = 1
   l_Jperio = ( jperio == 2 .OR. jperio == 7 )
   ln_linssh = ihdr(8) /= 0 ; ln_isfcav = ihdr(9) /= 0
   lwp = .FALSE. ; numout = 6 ; l_trdtra = .FALSE. ; l_trdtrc = .FALSE.
   ierr = dom_oce_alloc() ;   IF( ierr /= 0 ) STOP 'dom_oce_alloc'
   ALLOCATE( zun(jpi,jpj,jpk), zvn(jpi,jpj,jpk), zwn(jpi,jpj,jpk), ztb(jpi,jpj,jpk,kjpt), ztn(jpi,jpj,jpk,kjpt), zta(jpi,jpj,jpk,kjpt) )
   READ(iu) tmask, umask, vmask, wmask, e3t_b, e3t_n, e3t_a, e1e2t, r1_e1e2t
   READ(iu) mikt, mbkt
   READ(iu) zun, zvn, zwn, ztb, ztn, zta
   CLOSE(iu)
   CALL tra_adv_fct( 1, 1, 'TRA', zp2dt, zun, zvn, zwn, ztb, ztn, zta, kjpt, kn_fct_h, kn_fct_v )
   OPEN( newunit=iu, file=TRIM(cout), access='stream', form='unformatted', status='replace' )
   WRITE(iu) zta
   CLOSE(iu)
END PROGRAM fct_ref_driver
