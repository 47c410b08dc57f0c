! grtrans_b200 -- Fortran 90 binding of the C ABI (include/grtrans_b200.h) for jadexter/grtrans.
!
! SOURCE ONLY: there is no Fortran compiler in the build image or on the GPU boxes, so this file is not compiled or
! tested by this repository (DESIGN.md section 1).  It is what a maintainer of the reference adds to its Makefile:
!   gfortran -c grtrans_b200_c.f90 pgrtrans_b200.f90 ; link with -L<repo>/grtrans_b200 -lgrtrans_b200 -lcudart
! pgrtrans_b200.f90 replaces pgrtrans.f90 (module pgrtrans: ivals, ab, freqs, grtrans_main, del_pgrtrans_data), so
! grtrans.f90, grtrans_program.f90, read_inputs.f90 and grtrans_batch.py stay untouched.
      module grtrans_b200_c
        use iso_c_binding
        implicit none
        integer, parameter :: GRT_STRLEN = 100, GRT_PATHLEN = 300
      
        type, bind(C) :: grt_args
           integer(c_int) :: standard
           real(c_double) :: mumin, mumax
           integer(c_int) :: nmu
           real(c_double) :: phi0, spin, uout, uin, rcut
           integer(c_int) :: nrotype
           real(c_double) :: gridvals(4)
           integer(c_int) :: nn(3)
           integer(c_int) :: i1, i2
           character(kind=c_char) :: fname(GRT_STRLEN)
           real(c_double) :: dt
           integer(c_int) :: nt, nload, nmdot
           real(c_double) :: mdotmin, mdotmax
           character(kind=c_char) :: ename(GRT_STRLEN)
           real(c_double) :: mbh
           integer(c_int) :: nfreq
           real(c_double) :: fmin, fmax, muval, gmin, gmax, p1, p2, jetalpha
           character(kind=c_char) :: stype(GRT_STRLEN)
           integer(c_int) :: use_geokerr
           integer(c_int) :: nvals
           character(kind=c_char) :: iname(GRT_STRLEN)
           integer(c_int) :: cflag, extra, debug
           character(kind=c_char) :: outfile(GRT_PATHLEN)
           character(kind=c_char) :: fdfile(GRT_PATHLEN), fhfile(GRT_PATHLEN), fgfile(GRT_PATHLEN), fsim(GRT_STRLEN)
           integer(c_int) :: fnt, findf, fnfiles, fjonfix, fnw, fnfreq_tab, fnr, foffset, fdindf, fmagcrit
           real(c_double) :: frspot, fr0spot, fn0spot, ftscl, frscl, fwmin, fwmax, ffmin, ffmax, frmax, fsigt, ffcol, &
                             fmdot, fnscl, fnnthscl, fnnthp, fbeta
           integer(c_int) :: fbl06
           real(c_double) :: fnp, ftp, frin, frout, fthin, fthout, fphiin, fphiout, fscalefac, sigcut
           integer(c_int) :: epcoefindx(7)
           integer(c_int) :: nepotherargs
           real(c_double) :: epotherargs(8)
        end type grt_args
      
        interface
           subroutine grt_args_defaults(a) bind(C, name='grt_args_defaults')
             import :: grt_args
             type(grt_args), intent(out) :: a
           end subroutine
           ! pgrtrans.f90:12  grtrans_main
           integer(c_int) function grt_main(a) bind(C, name='grt_main')
             import :: grt_args, c_int
             type(grt_args), intent(in) :: a
           end function
           ! pgrtrans.f90:6-8  module arrays ivals(nvals,npix,ncams), ab(2,npix), freqs(nfreq)
           integer(c_int) function grt_get_ivals(p, dims) bind(C, name='grt_get_ivals')
             import :: c_ptr, c_int
             type(c_ptr), intent(out) :: p
             integer(c_int), intent(out) :: dims(3)
           end function
           integer(c_int) function grt_get_ab(p, dims) bind(C, name='grt_get_ab')
             import :: c_ptr, c_int
             type(c_ptr), intent(out) :: p
             integer(c_int), intent(out) :: dims(2)
           end function
           integer(c_int) function grt_get_freqs(p, n) bind(C, name='grt_get_freqs')
             import :: c_ptr, c_int
             type(c_ptr), intent(out) :: p
             integer(c_int), intent(out) :: n
           end function
           ! pgrtrans.f90:255  del_pgrtrans_data
           subroutine grt_del_data() bind(C, name='grt_del_data')
           end subroutine
           ! grtrans.f90:2  grtrans_run(ifile,outfile)
           integer(c_int) function grt_run(ifile, ofile) bind(C, name='grt_run')
             import :: c_char, c_int
             character(kind=c_char), intent(in) :: ifile(*), ofile(*)
           end function
           ! grtrans_driver.f90:86-110, 341-427  the debug = 1 dump `geodebug.out` of one pixel
           integer(c_int) function grt_geodebug(a, pixel, path) bind(C, name='grt_geodebug')
             import :: grt_args, c_int, c_char
             type(grt_args), intent(in) :: a
             integer(c_int), value :: pixel
             character(kind=c_char), intent(in) :: path(*)
           end function
           ! devices grt_main uses (n = 0: every visible device, or GRT_DEVICES / GRT_NGPU)
           integer(c_int) function grt_set_devices(devices, n) bind(C, name='grt_set_devices')
             import :: c_int
             integer(c_int), intent(in) :: devices(*)
             integer(c_int), value :: n
           end function
           ! radtrans_integrate.f90:139  integrate (batched over rays)
           integer(c_int) function grt_rt_integrate(iflag, neq, nrays, npts, s, j, K, tau, intensity, nptsout, &
                max_tau, hmax, oatol, ortol, thin, maxsteps) bind(C, name='grt_rt_integrate')
             import :: c_int, c_double
             integer(c_int), value :: iflag, neq, nrays, npts, maxsteps
             real(c_double), intent(in) :: s(*), j(*), K(*), tau(*)
             real(c_double), intent(out) :: intensity(*)
             integer(c_int), intent(out) :: nptsout(*)
             real(c_double), value :: max_tau, hmax, oatol, ortol, thin
           end function
        end interface
      
      contains
        subroutine f2c_string(f, c)
          character(len=*), intent(in) :: f
          character(kind=c_char), intent(out) :: c(:)
          integer :: i, n
          n = min(len_trim(f), size(c) - 1)
          do i = 1, n
             c(i) = f(i:i)
          end do
          c(n + 1) = c_null_char
        end subroutine
      end module grtrans_b200_c
