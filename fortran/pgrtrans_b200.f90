! grtrans_b200 -- Fortran 90 binding of the C ABI (include/grtrans_b200.h) for jadexter/grtrans.
!
! SOURCE ONLY: there is no Fortran compiler in the build image or on the GPU boxes, so this file is not compiled or
! tested by this repository (DESIGN.md section 1).  It is what a maintainer of the reference adds to its Makefile:
!   gfortran -c grtrans_b200_c.f90 pgrtrans_b200.f90 ; link with -L<repo>/grtrans_b200 -lgrtrans_b200 -lcudart
! pgrtrans_b200.f90 replaces pgrtrans.f90 (module pgrtrans: ivals, ab, freqs, grtrans_main, del_pgrtrans_data), so
! grtrans.f90, grtrans_program.f90, read_inputs.f90 and grtrans_batch.py stay untouched.
      module pgrtrans
      use iso_c_binding
      use grtrans_b200_c
      implicit none
      real(kind=8), dimension(:,:), allocatable :: ab
      real(kind=8), dimension(:,:,:), allocatable :: ivals
      real(kind=8), dimension(:), allocatable :: freqs

      contains

      ! same dummy arguments, same order and meaning as the reference's grtrans_main (pgrtrans.f90:12-58)
      subroutine grtrans_main( &
       standard,mumin,mumax,nmu,phi0,spin,uout,uin,rcut,nrotype,gridvals,nn,i1,i2,fname,dt,nt,nload,nmdot, &
       mdotmin,mdotmax,ename,mbh,nfreq,fmin,fmax,muval,gmin,gmax,p1,p2,jetalpha,stype,use_geokerr,nvals, &
       iname,cflag,extra,debug,outfile,fdfile,fhfile,fgfile,fsim,fnt,findf,fnfiles,fjonfix,fnw,fnfreq_tab, &
       fnr,foffset,fdindf,fmagcrit,frspot,fr0spot,fn0spot,ftscl,frscl,fwmin,fwmax,ffmin,ffmax,frmax,fsigt, &
       ffcol,fmdot,fnscl,fnnthscl,fnnthp,fbeta,fbl06,fnp,ftp,frin,frout,fthin,fthout,fphiin,fphiout, &
       fscalefac,sigcut,epcoefindx,epotherargs,nepotherargs)
      integer, intent(in) :: standard,nrotype,nvals,nfreq,nmu,cflag,nt,nmdot,nload,extra,i1,i2,nepotherargs,debug
      logical, intent(in) :: use_geokerr
      real(kind=8), intent(in) :: mumax,mumin,rcut,mbh,uout,uin,fmin,fmax,dt,mdotmin,mdotmax,phi0,muval,gmin,gmax,p1,p2,jetalpha
      character(len=*), intent(in) :: ename,fname,iname,stype,outfile
      real(kind=8), dimension(4), intent(in) :: gridvals
      real(kind=8) :: spin
      integer, dimension(3), intent(in) :: nn
      character(len=*), intent(in) :: fdfile,fhfile,fgfile,fsim
      integer, intent(in) :: fnt,findf,fnfiles,fjonfix,fnw,fnfreq_tab,fnr,foffset,fdindf,fmagcrit,fbl06
      real(kind=8), intent(in) :: frspot,fr0spot,fn0spot,ftscl,frscl,fwmin,fwmax,ffmin,ffmax,frmax,fsigt,ffcol,fmdot, &
           fnnthp,fnnthscl,fnscl,fbeta,ftp,fnp,frin,frout,fthin,fthout,fphiin,fphiout,fscalefac,sigcut
      real(kind=8), dimension(nepotherargs), intent(in) :: epotherargs
      integer, dimension(7), intent(in) :: epcoefindx
      type(grt_args) :: a
      type(c_ptr) :: p
      integer(c_int) :: dims(3), d2(2), nf, rc
      real(c_double), pointer :: iv(:,:,:), abp(:,:), fr(:)

      call grt_args_defaults(a)
      a%standard = standard
      a%mumin = mumin
      a%mumax = mumax
      a%nmu = nmu
      a%phi0 = phi0
      a%spin = spin
      a%uout = uout
      a%uin = uin
      a%rcut = rcut
      a%nrotype = nrotype
      a%gridvals = gridvals
      a%nn = nn
      a%i1 = i1
      a%i2 = i2
      call f2c_string(fname, a%fname)
      a%dt = dt
      a%nt = nt
      a%nload = nload
      a%nmdot = nmdot
      a%mdotmin = mdotmin
      a%mdotmax = mdotmax
      call f2c_string(ename, a%ename)
      a%mbh = mbh
      a%nfreq = nfreq
      a%fmin = fmin
      a%fmax = fmax
      a%muval = muval
      a%gmin = gmin
      a%gmax = gmax
      a%p1 = p1
      a%p2 = p2
      a%jetalpha = jetalpha
      call f2c_string(stype, a%stype)
      a%use_geokerr = merge(1_c_int, 0_c_int, use_geokerr)
      a%nvals = nvals
      call f2c_string(iname, a%iname)
      a%cflag = cflag
      a%extra = extra
      a%debug = debug
      call f2c_string(outfile, a%outfile)
      call f2c_string(fdfile, a%fdfile)
      call f2c_string(fhfile, a%fhfile)
      call f2c_string(fgfile, a%fgfile)
      call f2c_string(fsim, a%fsim)
      a%fnt = fnt
      a%findf = findf
      a%fnfiles = fnfiles
      a%fjonfix = fjonfix
      a%fnw = fnw
      a%fnfreq_tab = fnfreq_tab
      a%fnr = fnr
      a%foffset = foffset
      a%fdindf = fdindf
      a%fmagcrit = fmagcrit
      a%frspot = frspot
      a%fr0spot = fr0spot
      a%fn0spot = fn0spot
      a%ftscl = ftscl
      a%frscl = frscl
      a%fwmin = fwmin
      a%fwmax = fwmax
      a%ffmin = ffmin
      a%ffmax = ffmax
      a%frmax = frmax
      a%fsigt = fsigt
      a%ffcol = ffcol
      a%fmdot = fmdot
      a%fnscl = fnscl
      a%fnnthscl = fnnthscl
      a%fnnthp = fnnthp
      a%fbeta = fbeta
      a%fbl06 = fbl06
      a%fnp = fnp
      a%ftp = ftp
      a%frin = frin
      a%frout = frout
      a%fthin = fthin
      a%fthout = fthout
      a%fphiin = fphiin
      a%fphiout = fphiout
      a%fscalefac = fscalefac
      a%sigcut = sigcut
      a%epcoefindx = epcoefindx
      a%epotherargs = 0d0
      a%epotherargs(1:min(nepotherargs,8)) = epotherargs(1:min(nepotherargs,8))
      a%nepotherargs = nepotherargs
      rc = grt_main(a)          ! GPU: all pixels i1..i2 of all mu0 / mdot / frequency cameras; writes outfile if non-blank
      if (rc /= 0) then
         write(6,*) 'grtrans_b200: grt_main failed with status ', rc
         return
      endif
      rc = grt_get_ivals(p, dims)
      call c_f_pointer(p, iv, dims)
      allocate(ivals(dims(1),dims(2),dims(3)))
      ivals = iv
      rc = grt_get_ab(p, d2)
      call c_f_pointer(p, abp, d2)
      allocate(ab(d2(1),d2(2)))
      ab = abp
      rc = grt_get_freqs(p, nf)
      call c_f_pointer(p, fr, (/nf/))
      allocate(freqs(nf))
      freqs = fr
      call grt_del_data()
      end subroutine grtrans_main

      subroutine del_pgrtrans_data()
      if (allocated(ivals)) deallocate(ivals)
      if (allocated(ab)) deallocate(ab)
      if (allocated(freqs)) deallocate(freqs)
      end subroutine del_pgrtrans_data

      end module pgrtrans
