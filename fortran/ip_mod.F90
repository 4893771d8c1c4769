!> @file
!! Drop-in replacement for NCEPLIBS-ip's `ip_mod` (src/ip_mod.F90) over libipb200.so.
!!
!! Host code stays Fortran: this file re-creates the public surface of the reference's
!! `ipolates_mod` (src/ipolates.F90:21-30), `ipolatev_mod` (src/ipolatev.F90:19-28) and the descriptor
!! forms of `gdswzd_mod` (src/gdswzd_mod.F90:35-44) — same module, generic and specific names, same
!! dummy-argument lists per build kind — as thin ISO_C_BINDING wrappers.  There is no logic here:
!! every routine forwards its arguments, by reference, to the C entry point of the same name with
!! the kind suffix (include/ipb200.h).  Build once per kind exactly like the reference does
!! (src/CMakeLists.txt:36-76):
!!
!!   gfortran -cpp -ffree-line-length-none -DLSIZE=4 -c ip_mod.F90                        -> libip_4  (INTEGER(4), REAL(4))
!!   gfortran -cpp -ffree-line-length-none -DLSIZE=D -fdefault-real-8 -c ip_mod.F90           -> libip_d  (INTEGER(4), REAL(8))
!!   gfortran -cpp -DLSIZE=8 -fdefault-real-8 -fdefault-integer-8 ...   -> libip_8  (INTEGER(8), REAL(8))
!!
!! and link with -lipb200 -lcudart.  The specific routines keep the reference's bind(c) names
!! (`ipolates_grib2`, ... , `gdswzd`, `gdswzd_grib1`), so C callers of libip_<kind> keep working too.
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (see DESIGN.md §2).
!! `gdswzd_grid` (the class(ip_grid) form) is not provided: grid objects do not cross the boundary.

#if (LSIZE==8)
#define IPB_INT integer(c_long)
#define IPB_REAL real(c_double)
#define IPB_N_S2  'ipolates_grib2_8'
#define IPB_N_S2F 'ipolates_grib2_single_field_8'
#define IPB_N_S1  'ipolates_grib1_8'
#define IPB_N_S1F 'ipolates_grib1_single_field_8'
#define IPB_N_V2  'ipolatev_grib2_8'
#define IPB_N_V2F 'ipolatev_grib2_single_field_8'
#define IPB_N_V1  'ipolatev_grib1_8'
#define IPB_N_V1F 'ipolatev_grib1_single_field_8'
#define IPB_N_G2  'gdswzd_8'
#define IPB_N_G1  'gdswzd_grib1_8'
#elif (LSIZE==4)
#define IPB_INT integer(c_int)
#define IPB_REAL real(c_float)
#define IPB_N_S2  'ipolates_grib2_4'
#define IPB_N_S2F 'ipolates_grib2_single_field_4'
#define IPB_N_S1  'ipolates_grib1_4'
#define IPB_N_S1F 'ipolates_grib1_single_field_4'
#define IPB_N_V2  'ipolatev_grib2_4'
#define IPB_N_V2F 'ipolatev_grib2_single_field_4'
#define IPB_N_V1  'ipolatev_grib1_4'
#define IPB_N_V1F 'ipolatev_grib1_single_field_4'
#define IPB_N_G2  'gdswzd_4'
#define IPB_N_G1  'gdswzd_grib1_4'
#else
#define IPB_INT integer(c_int)
#define IPB_REAL real(c_double)
#define IPB_N_S2  'ipolates_grib2_d'
#define IPB_N_S2F 'ipolates_grib2_single_field_d'
#define IPB_N_S1  'ipolates_grib1_d'
#define IPB_N_S1F 'ipolates_grib1_single_field_d'
#define IPB_N_V2  'ipolatev_grib2_d'
#define IPB_N_V2F 'ipolatev_grib2_single_field_d'
#define IPB_N_V1  'ipolatev_grib1_d'
#define IPB_N_V1F 'ipolatev_grib1_single_field_d'
#define IPB_N_G2  'gdswzd_d'
#define IPB_N_G1  'gdswzd_grib1_d'
#endif

!> C entry points of libipb200.so (include/ipb200.h).
module ipb200_c
  use iso_c_binding
  implicit none
  interface
     subroutine c_ipolates_grib2(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo, &
          igdtleno, &
          mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go,iret) bind(c, name=IPB_N_S2)
       import
       IPB_INT :: ip, ipopt(20), igdtnumi, igdtmpli(*), igdtleni, igdtnumo, igdtmplo(*), &
            igdtleno, mi, mo, km, ibi(*)
       logical(c_bool) :: li(*), lo(*)
       IPB_REAL :: gi(*), rlat(*), rlon(*), go(*)
       IPB_INT :: no, ibo(*), iret
     end subroutine c_ipolates_grib2
     subroutine c_ipolates_grib2_sf(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo, &
          igdtleno, &
          mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go,iret) bind(c, name=IPB_N_S2F)
       import
       IPB_INT :: ip, ipopt(20), igdtnumi, igdtmpli(*), igdtleni, igdtnumo, igdtmplo(*), &
            igdtleno, mi, mo, km, ibi
       logical(c_bool) :: li(*), lo(*)
       IPB_REAL :: gi(*), rlat(*), rlon(*), go(*)
       IPB_INT :: no, ibo, iret
     end subroutine c_ipolates_grib2_sf
     subroutine c_ipolates_grib1(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo, &
          lo,go,iret) bind(c, name=IPB_N_S1)
       import
       IPB_INT :: ip, ipopt(20), kgdsi(200), kgdso(200), mi, mo, km, ibi(*)
       logical(c_bool) :: li(*), lo(*)
       IPB_REAL :: gi(*), rlat(*), rlon(*), go(*)
       IPB_INT :: no, ibo(*), iret
     end subroutine c_ipolates_grib1
     subroutine c_ipolates_grib1_sf(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,gi,no,rlat,rlon, &
          ibo,lo,go,iret) bind(c, name=IPB_N_S1F)
       import
       IPB_INT :: ip, ipopt(20), kgdsi(200), kgdso(200), mi, mo, km, ibi
       logical(c_bool) :: li(*), lo(*)
       IPB_REAL :: gi(*), rlat(*), rlon(*), go(*)
       IPB_INT :: no, ibo, iret
     end subroutine c_ipolates_grib1_sf
     subroutine c_ipolatev_grib2(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo, &
          igdtleno, &
          mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot,srot,ibo,lo,uo,vo,iret) bind(c, name=IPB_N_V2)
       import
       IPB_INT :: ip, ipopt(20), igdtnumi, igdtmpli(*), igdtleni, igdtnumo, igdtmplo(*), &
            igdtleno, mi, mo, km, ibi(*)
       logical(c_bool) :: li(*), lo(*)
       IPB_REAL :: ui(*), vi(*), rlat(*), rlon(*), crot(*), srot(*), uo(*), vo(*)
       IPB_INT :: no, ibo(*), iret
     end subroutine c_ipolatev_grib2
     subroutine c_ipolatev_grib2_sf(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo, &
          igdtleno, &
          mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot,srot,ibo,lo,uo,vo,iret) bind(c, &
               name=IPB_N_V2F)
       import
       IPB_INT :: ip, ipopt(20), igdtnumi, igdtmpli(*), igdtleni, igdtnumo, igdtmplo(*), &
            igdtleno, mi, mo, km, ibi
       logical(c_bool) :: li(*), lo(*)
       IPB_REAL :: ui(*), vi(*), rlat(*), rlon(*), crot(*), srot(*), uo(*), vo(*)
       IPB_INT :: no, ibo, iret
     end subroutine c_ipolatev_grib2_sf
     subroutine c_ipolatev_grib1(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,ui,vi,no,rlat,rlon, &
          crot,srot,ibo,lo,uo,vo,iret) &
          bind(c, name=IPB_N_V1)
       import
       IPB_INT :: ip, ipopt(20), kgdsi(200), kgdso(200), mi, mo, km, ibi(*)
       logical(c_bool) :: li(*), lo(*)
       IPB_REAL :: ui(*), vi(*), rlat(*), rlon(*), crot(*), srot(*), uo(*), vo(*)
       IPB_INT :: no, ibo(*), iret
     end subroutine c_ipolatev_grib1
     subroutine c_ipolatev_grib1_sf(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,ui,vi,no,rlat,rlon, &
          crot,srot,ibo,lo,uo,vo,iret) &
          bind(c, name=IPB_N_V1F)
       import
       IPB_INT :: ip, ipopt(20), kgdsi(200), kgdso(200), mi, mo, km, ibi
       logical(c_bool) :: li(*), lo(*)
       IPB_REAL :: ui(*), vi(*), rlat(*), rlon(*), crot(*), srot(*), uo(*), vo(*)
       IPB_INT :: no, ibo, iret
     end subroutine c_ipolatev_grib1_sf
     ! scalars by value, optional outputs may be absent (passed as NULL): src/gdswzd_c.F90:173-210
     subroutine c_gdswzd(igdtnum,igdtmpl,igdtlen,iopt,npts,fill,xpts,ypts,rlon,rlat,nret, &
          crot,srot,xlon,xlat,ylon,ylat,area) bind(c, name=IPB_N_G2)
       import
       IPB_INT, value :: igdtnum, igdtlen, iopt, npts
       IPB_INT :: igdtmpl(*), nret
       IPB_REAL, value :: fill
       IPB_REAL :: xpts(*), ypts(*), rlon(*), rlat(*)
       IPB_REAL, optional :: crot(*), srot(*), xlon(*), xlat(*), ylon(*), ylat(*), area(*)
     end subroutine c_gdswzd
     subroutine c_gdswzd_grib1(kgds,iopt,npts,fill,xpts,ypts,rlon,rlat,nret, &
          crot,srot,xlon,xlat,ylon,ylat,area) bind(c, name=IPB_N_G1)
       import
       IPB_INT, value :: iopt, npts
       IPB_INT :: kgds(200), nret
       IPB_REAL, value :: fill
       IPB_REAL :: xpts(*), ypts(*), rlon(*), rlat(*)
       IPB_REAL, optional :: crot(*), srot(*), xlon(*), xlat(*), ylon(*), ylat(*), area(*)
     end subroutine c_gdswzd_grib1
  end interface
end module ipb200_c

!> Mirror of src/ip_interpolators_mod.F90:17-27.
module ip_interpolators_mod
  implicit none
  integer, parameter, public :: BILINEAR_INTERP_ID = 0, BICUBIC_INTERP_ID = 1, &
       NEIGHBOR_INTERP_ID = 2, &
       BUDGET_INTERP_ID = 3, SPECTRAL_INTERP_ID = 4, NEIGHBOR_BUDGET_INTERP_ID = 6
end module ip_interpolators_mod

!> Mirror of src/ipolates.F90 (module ipolates_mod).
module ipolates_mod
  use iso_c_binding
  use ipb200_c
  implicit none
  private
  public :: ipolates, ipolates_grib2, ipolates_grib1_single_field, ipolates_grib1, &
       ipolates_grib2_single_field
  interface ipolates
     module procedure ipolates_grib1
     module procedure ipolates_grib1_single_field
     module procedure ipolates_grib2
     module procedure ipolates_grib2_single_field
  end interface ipolates
contains
  subroutine ipolates_grib2(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo,igdtleno, &
       mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go,iret) bind(c)       ! src/ipolates.F90:587-636
    IPB_INT, intent(in) :: ip, ipopt(20), km, mi, mo, ibi(km), igdtnumi, igdtleni, &
         igdtmpli(igdtleni)
    IPB_INT, intent(in) :: igdtnumo, igdtleno, igdtmplo(igdtleno)
    IPB_INT, intent(out) :: no, iret, ibo(km)
    logical(c_bool), intent(in) :: li(mi,km)
    logical(c_bool), intent(out) :: lo(mo,km)
    IPB_REAL, intent(in) :: gi(mi,km)
    IPB_REAL, intent(inout) :: rlat(mo), rlon(mo)
    IPB_REAL, intent(out) :: go(mo,km)
    call c_ipolates_grib2(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo,igdtleno, &
         mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go,iret)
  end subroutine ipolates_grib2
  subroutine ipolates_grib2_single_field(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo, &
       igdtmplo,igdtleno, &
       mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go,iret) bind(c)       ! src/ipolates.F90:808-864
    IPB_INT, intent(in) :: ip, ipopt(20), km, mi, mo, ibi, igdtnumi, igdtleni, &
         igdtmpli(igdtleni)
    IPB_INT, intent(in) :: igdtnumo, igdtleno, igdtmplo(igdtleno)
    IPB_INT, intent(out) :: no, iret, ibo
    logical(c_bool), intent(in) :: li(mi)
    logical(c_bool), intent(out) :: lo(mo)
    IPB_REAL, intent(in) :: gi(mi)
    IPB_REAL, intent(inout) :: rlat(mo), rlon(mo)
    IPB_REAL, intent(out) :: go(mo)
    call c_ipolates_grib2_sf(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo,igdtleno, &
         mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go,iret)
  end subroutine ipolates_grib2_single_field
  subroutine ipolates_grib1(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go, &
       iret) bind(c)  ! :293-334
    IPB_INT, intent(in) :: ip, ipopt(20), km, mi, mo, ibi(km), kgdsi(200), kgdso(200)
    IPB_INT, intent(out) :: no, iret, ibo(km)
    logical(c_bool), intent(in) :: li(mi,km)
    logical(c_bool), intent(out) :: lo(mo,km)
    IPB_REAL, intent(in) :: gi(mi,km)
    IPB_REAL, intent(inout) :: rlat(mo), rlon(mo)
    IPB_REAL, intent(out) :: go(mo,km)
    call c_ipolates_grib1(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go,iret)
  end subroutine ipolates_grib1
  subroutine ipolates_grib1_single_field(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,gi,no,rlat, &
       rlon,ibo,lo,go,iret) bind(c)  ! :158-206
    IPB_INT, intent(in) :: ip, ipopt(20), km, mi, mo, ibi, kgdsi(200), kgdso(200)
    IPB_INT, intent(out) :: no, iret, ibo
    logical(c_bool), intent(in) :: li(mi)
    logical(c_bool), intent(out) :: lo(mo)
    IPB_REAL, intent(in) :: gi(mi)
    IPB_REAL, intent(inout) :: rlat(mo), rlon(mo)
    IPB_REAL, intent(out) :: go(mo)
    call c_ipolates_grib1_sf(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,gi,no,rlat,rlon,ibo,lo,go, &
         iret)
  end subroutine ipolates_grib1_single_field
end module ipolates_mod

!> Mirror of src/ipolatev.F90 (module ipolatev_mod).
module ipolatev_mod
  use iso_c_binding
  use ipb200_c
  implicit none
  private
  public :: ipolatev, ipolatev_grib2, ipolatev_grib1_single_field, ipolatev_grib1, &
       ipolatev_grib2_single_field
  interface ipolatev
     module procedure ipolatev_grib1
     module procedure ipolatev_grib1_single_field
     module procedure ipolatev_grib2
     module procedure ipolatev_grib2_single_field
  end interface ipolatev
contains
  subroutine ipolatev_grib2(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo,igdtleno, &
       mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot,srot,ibo,lo,uo,vo, &
            iret) bind(c)      ! src/ipolatev.F90:382-434
    IPB_INT, intent(in) :: ip, ipopt(20), km, mi, mo, ibi(km), igdtnumi, igdtleni, &
         igdtmpli(igdtleni)
    IPB_INT, intent(in) :: igdtnumo, igdtleno, igdtmplo(igdtleno)
    IPB_INT, intent(out) :: no, iret, ibo(km)
    logical(c_bool), intent(in) :: li(mi,km)
    logical(c_bool), intent(out) :: lo(mo,km)
    IPB_REAL, intent(in) :: ui(mi,km), vi(mi,km)
    IPB_REAL, intent(inout) :: rlat(mo), rlon(mo), crot(mo), srot(mo)
    IPB_REAL, intent(out) :: uo(mo,km), vo(mo,km)
    call c_ipolatev_grib2(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo,igdtleno, &
         mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot,srot,ibo,lo,uo,vo,iret)
  end subroutine ipolatev_grib2
  subroutine ipolatev_grib2_single_field(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo, &
       igdtmplo,igdtleno, &
       mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot,srot,ibo,lo,uo,vo, &
            iret) bind(c)      ! src/ipolatev.F90:832-891
    IPB_INT, intent(in) :: ip, ipopt(20), km, mi, mo, ibi, igdtnumi, igdtleni, &
         igdtmpli(igdtleni)
    IPB_INT, intent(in) :: igdtnumo, igdtleno, igdtmplo(igdtleno)
    IPB_INT, intent(out) :: no, iret, ibo
    logical(c_bool), intent(in) :: li(mi)
    logical(c_bool), intent(out) :: lo(mo)
    IPB_REAL, intent(in) :: ui(mi), vi(mi)
    IPB_REAL, intent(inout) :: rlat(mo), rlon(mo), crot(mo), srot(mo)
    IPB_REAL, intent(out) :: uo(mo), vo(mo)
    call c_ipolatev_grib2_sf(ip,ipopt,igdtnumi,igdtmpli,igdtleni,igdtnumo,igdtmplo,igdtleno, &
         mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot,srot,ibo,lo,uo,vo,iret)
  end subroutine ipolatev_grib2_single_field
  !> The reference temporarily ORs 256 into kgds(11) for grid 203 (src/ipolatev.F90:602-626, hence
  !! intent(inout)); libipb200 applies that switch on its own copy and leaves the caller's arrays alone.
  subroutine ipolatev_grib1(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot, &
       srot,ibo,lo,uo,vo,iret) bind(c)  ! :565-628
    IPB_INT, intent(in) :: ip, ipopt(20), km, mi, mo, ibi(km)
    IPB_INT, intent(inout) :: kgdsi(200), kgdso(200)
    IPB_INT, intent(out) :: no, iret, ibo(km)
    logical(c_bool), intent(in) :: li(mi,km)
    logical(c_bool), intent(out) :: lo(mo,km)
    IPB_REAL, intent(in) :: ui(mi,km), vi(mi,km)
    IPB_REAL, intent(inout) :: rlat(mo), rlon(mo), crot(mo), srot(mo)
    IPB_REAL, intent(out) :: uo(mo,km), vo(mo,km)
    call c_ipolatev_grib1(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot,srot, &
         ibo,lo,uo,vo,iret)
  end subroutine ipolatev_grib1
  subroutine ipolatev_grib1_single_field(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,ui,vi,no,rlat, &
       rlon,crot,srot,ibo,lo,uo,vo,iret) &
       bind(c)                                                                    ! src/ipolatev.F90:680-750
    IPB_INT, intent(in) :: ip, ipopt(20), km, mi, mo, ibi
    IPB_INT, intent(inout) :: kgdsi(200), kgdso(200)
    IPB_INT, intent(out) :: no, iret, ibo
    logical(c_bool), intent(in) :: li(mi)
    logical(c_bool), intent(out) :: lo(mo)
    IPB_REAL, intent(in) :: ui(mi), vi(mi)
    IPB_REAL, intent(inout) :: rlat(mo), rlon(mo), crot(mo), srot(mo)
    IPB_REAL, intent(out) :: uo(mo), vo(mo)
    call c_ipolatev_grib1_sf(ip,ipopt,kgdsi,kgdso,mi,mo,km,ibi,li,ui,vi,no,rlat,rlon,crot, &
         srot,ibo,lo,uo,vo,iret)
  end subroutine ipolatev_grib1_single_field
end module ipolatev_mod

!> Mirror of the descriptor forms of src/gdswzd_mod.F90 (generic gdswzd) and of src/gdswzd_c.F90.
module gdswzd_mod
  use iso_c_binding
  use ipb200_c
  implicit none
  private
  public :: gdswzd_2d_array_grib1, gdswzd_grib1, gdswzd
  interface gdswzd
     module procedure gdswzd_1d_array
     module procedure gdswzd_2d_array
     module procedure gdswzd_scalar
     module procedure gdswzd_grib1
     module procedure gdswzd_2d_array_grib1
  end interface gdswzd
contains
  subroutine gdswzd_1d_array(igdtnum,igdtmpl,igdtlen,iopt,npts,fill,xpts,ypts,rlon,rlat,nret, &
       crot,srot,xlon,xlat,ylon,ylat, &
            area)                                       ! src/gdswzd_mod.F90:665-690
    integer, intent(in) :: igdtnum, igdtlen, igdtmpl(igdtlen), iopt, npts
    integer, intent(out) :: nret
    real, intent(in) :: fill
    real, intent(inout) :: rlon(npts), rlat(npts), xpts(npts), ypts(npts)
    real, optional, intent(out) :: crot(npts), srot(npts), xlon(npts), xlat(npts), &
         ylon(npts), ylat(npts), area(npts)
    call c_gdswzd(igdtnum,igdtmpl,igdtlen,iopt,npts,fill,xpts,ypts,rlon,rlat,nret,crot,srot, &
         xlon,xlat,ylon,ylat,area)
  end subroutine gdswzd_1d_array
  subroutine gdswzd_2d_array(igdtnum,igdtmpl,igdtlen,iopt,npts,fill,xpts,ypts,rlon,rlat,nret, &
       crot,srot,xlon,xlat,ylon,ylat, &
            area)                                       ! src/gdswzd_mod.F90:459-481
    integer, intent(in) :: igdtnum, igdtlen, igdtmpl(igdtlen), iopt, npts
    integer, intent(out) :: nret
    real, intent(in) :: fill
    real, intent(inout) :: rlon(:,:), rlat(:,:), xpts(:,:), ypts(:,:)    ! contiguous, &
         npts elements in all
    real, optional, intent(out) :: crot(:,:), srot(:,:), xlon(:,:), xlat(:,:), ylon(:,:), &
         ylat(:,:), area(:,:)
    call c_gdswzd(igdtnum,igdtmpl,igdtlen,iopt,npts,fill,xpts,ypts,rlon,rlat,nret,crot,srot, &
         xlon,xlat,ylon,ylat,area)
  end subroutine gdswzd_2d_array
  subroutine gdswzd_scalar(igdtnum,igdtmpl,igdtlen,iopt,npts,fill,xpts,ypts,rlon,rlat,nret, &
       crot,srot,xlon,xlat,ylon,ylat, &
            area)                                       ! src/gdswzd_mod.F90:278-381
    integer, intent(in) :: igdtnum, igdtlen, igdtmpl(igdtlen), iopt, npts
    integer, intent(out) :: nret
    real, intent(in) :: fill
    real, intent(inout) :: rlon, rlat, xpts, ypts
    real, optional, intent(out) :: crot, srot, xlon, xlat, ylon, ylat, area
    real :: x1(1), y1(1), lo1(1), la1(1), e(1,7)
    x1 = xpts; y1 = ypts; lo1 = rlon; la1 = rlat
    call c_gdswzd(igdtnum,igdtmpl,igdtlen,iopt,1,fill,x1,y1,lo1,la1,nret,e(:,1),e(:,2),e(:, &
         3),e(:,4),e(:,5),e(:,6),e(:,7))
    xpts = x1(1); ypts = y1(1); rlon = lo1(1); rlat = la1(1)
    if (present(crot)) crot = e(1,1)
    if (present(srot)) srot = e(1,2)
    if (present(xlon)) xlon = e(1,3)
    if (present(xlat)) xlat = e(1,4)
    if (present(ylon)) ylon = e(1,5)
    if (present(ylat)) ylat = e(1,6)
    if (present(area)) area = e(1,7)
  end subroutine gdswzd_scalar
  subroutine gdswzd_grib1(kgds,iopt,npts,fill,xpts,ypts,rlon,rlat,nret,crot,srot,xlon,xlat, &
       ylon,ylat,area)  ! :758-781
    integer, intent(in) :: kgds(200), iopt, npts
    integer, intent(out) :: nret
    real, intent(in) :: fill
    real, intent(inout) :: rlon(npts), rlat(npts), xpts(npts), ypts(npts)
    real, optional, intent(out) :: crot(npts), srot(npts), xlon(npts), xlat(npts), &
         ylon(npts), ylat(npts), area(npts)
    call c_gdswzd_grib1(kgds,iopt,npts,fill,xpts,ypts,rlon,rlat,nret,crot,srot,xlon,xlat, &
         ylon,ylat,area)
  end subroutine gdswzd_grib1
  subroutine gdswzd_2d_array_grib1(kgds,iopt,npts,fill,xpts,ypts,rlon,rlat,nret,crot,srot, &
       xlon,xlat,ylon,ylat,area)  ! :850-875
    integer, intent(in) :: kgds(200), iopt, npts
    integer, intent(out) :: nret
    real, intent(in) :: fill
    real, intent(inout) :: rlon(:,:), rlat(:,:), xpts(:,:), ypts(:,:)
    real, optional, intent(out) :: crot(:,:), srot(:,:), xlon(:,:), xlat(:,:), ylon(:,:), &
         ylat(:,:), area(:,:)
    call c_gdswzd_grib1(kgds,iopt,npts,fill,xpts,ypts,rlon,rlat,nret,crot,srot,xlon,xlat, &
         ylon,ylat,area)
  end subroutine gdswzd_2d_array_grib1
end module gdswzd_mod

!> Top-level module, as src/ip_mod.F90:6-20.
module ip_mod
  use ip_interpolators_mod, only: BILINEAR_INTERP_ID, BICUBIC_INTERP_ID, NEIGHBOR_INTERP_ID, &
       BUDGET_INTERP_ID, SPECTRAL_INTERP_ID, NEIGHBOR_BUDGET_INTERP_ID
  use ipolates_mod
  use ipolatev_mod
  use gdswzd_mod
end module ip_mod

!> C symbols `gdswzd` / `gdswzd_grib1` of libip_<kind> (src/gdswzd_c.F90:173-210, :259-292; prototypes
!! src/iplib_4.h:146-191): scalars by value, straight through to the kind-suffixed entry point.
subroutine gdswzd_c(igdtnum,igdtmpl,igdtlen,iopt,npts,fill,xpts,ypts,rlon,rlat,nret, &
     crot,srot,xlon,xlat,ylon,ylat,area) bind(c, name='gdswzd')
  use iso_c_binding
  use ipb200_c
  implicit none
  IPB_INT, value, intent(in) :: igdtnum, igdtlen, iopt, npts
  IPB_INT, intent(in) :: igdtmpl(igdtlen)
  IPB_INT, intent(out) :: nret
  IPB_REAL, value, intent(in) :: fill
  IPB_REAL, intent(inout) :: xpts(npts), ypts(npts), rlon(npts), rlat(npts)
  IPB_REAL, intent(out) :: crot(npts), srot(npts), xlon(npts), xlat(npts), ylon(npts), &
       ylat(npts), area(npts)
  call c_gdswzd(igdtnum,igdtmpl,igdtlen,iopt,npts,fill,xpts,ypts,rlon,rlat,nret,crot,srot, &
       xlon,xlat,ylon,ylat,area)
end subroutine gdswzd_c

subroutine gdswzd_c_grib1(kgds,iopt,npts,fill,xpts,ypts,rlon,rlat,nret, &
     crot,srot,xlon,xlat,ylon,ylat,area) bind(c, name='gdswzd_grib1')
  use iso_c_binding
  use ipb200_c
  implicit none
  IPB_INT, intent(in) :: kgds(200)
  IPB_INT, value, intent(in) :: iopt, npts
  IPB_INT, intent(out) :: nret
  IPB_REAL, value, intent(in) :: fill
  IPB_REAL, intent(inout) :: xpts(npts), ypts(npts), rlon(npts), rlat(npts)
  IPB_REAL, intent(out) :: crot(npts), srot(npts), xlon(npts), xlat(npts), ylon(npts), &
       ylat(npts), area(npts)
  call c_gdswzd_grib1(kgds,iopt,npts,fill,xpts,ypts,rlon,rlat,nret,crot,srot,xlon,xlat,ylon, &
       ylat,area)
end subroutine gdswzd_c_grib1
