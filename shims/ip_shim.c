/* ip_shim.c — libip_4.so / libip_d.so / libip_8.so: the reference's UNSUFFIXED C symbols over libipb200.so.
 *
 * NCEPLIBS-ip builds one library per kind from one source (src/CMakeLists.txt:36-76) and every one of them exports the
 * same names (ipolates_grib2, ..., gdswzd, gdswzd_grib1).  libipb200.so carries all three kinds with a suffix; this
 * file, compiled once per kind (-DLSIZE=4 | D | 8), restores the reference's link interface:
 *     gcc -shared -fPIC -DLSIZE=4 -Iinclude shims/ip_shim.c -o libip_4.so -L. -lipb200
 * so that C callers written against src/iplib_4.h etc. (tests/tst_gdswzd_grib2.c) link unchanged.  No logic lives here. */
#include "iplib.h"
#include "ipb200.h"

#if defined(LSIZE) && (LSIZE == 4)
#define K(name) name##_4
#elif defined(LSIZE) && (LSIZE == 8)
#define K(name) name##_8
#else
#define K(name) name##_d
#endif

void gdswzd(ip_int igdtnum, ip_int* igdtmpl, ip_int igdtlen, ip_int iopt, ip_int npts, ip_real fill, ip_real* xpts, ip_real* ypts,
            ip_real* rlon, ip_real* rlat, ip_int* nret, ip_real* crot, ip_real* srot, ip_real* xlon, ip_real* xlat, ip_real* ylon,
            ip_real* ylat, ip_real* area) {
  K(gdswzd)(igdtnum, igdtmpl, igdtlen, iopt, npts, fill, xpts, ypts, rlon, rlat, nret, crot, srot, xlon, xlat, ylon, ylat, area);
}
void gdswzd_grib1(ip_int* kgds, ip_int iopt, ip_int npts, ip_real fill, ip_real* xpts, ip_real* ypts, ip_real* rlon, ip_real* rlat,
                  ip_int* nret, ip_real* crot, ip_real* srot, ip_real* xlon, ip_real* xlat, ip_real* ylon, ip_real* ylat, ip_real* area) {
  K(gdswzd_grib1)(kgds, iopt, npts, fill, xpts, ypts, rlon, rlat, nret, crot, srot, xlon, xlat, ylon, ylat, area);
}

#define G2_ARGS const ip_int *ip, const ip_int *ipopt, const ip_int *igdtnumi, const ip_int *igdtmpli, const ip_int *igdtleni, \
                const ip_int *igdtnumo, const ip_int *igdtmplo, const ip_int *igdtleno, const ip_int *mi, const ip_int *mo,    \
                const ip_int *km, const ip_int *ibi, const ip_logical *li
#define G2_PASS ip, ipopt, igdtnumi, igdtmpli, igdtleni, igdtnumo, igdtmplo, igdtleno, mi, mo, km, ibi, li
#define G1_ARGS const ip_int *ip, const ip_int *ipopt, const ip_int *kgdsi, const ip_int *kgdso, const ip_int *mi, const ip_int *mo, \
                const ip_int *km, const ip_int *ibi, const ip_logical *li
#define G1_PASS ip, ipopt, kgdsi, kgdso, mi, mo, km, ibi, li
#define S_TAIL const ip_real *gi, ip_int *no, ip_real *rlat, ip_real *rlon, ip_int *ibo, ip_logical *lo, ip_real *go, ip_int *iret
#define S_PASS gi, no, rlat, rlon, ibo, lo, go, iret
#define V_TAIL const ip_real *ui, const ip_real *vi, ip_int *no, ip_real *rlat, ip_real *rlon, ip_real *crot, ip_real *srot, \
               ip_int *ibo, ip_logical *lo, ip_real *uo, ip_real *vo, ip_int *iret
#define V_PASS ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret

void ipolates_grib2(G2_ARGS, S_TAIL) { K(ipolates_grib2)(G2_PASS, S_PASS); }
void ipolates_grib2_single_field(G2_ARGS, S_TAIL) { K(ipolates_grib2_single_field)(G2_PASS, S_PASS); }
void ipolates_grib1(G1_ARGS, S_TAIL) { K(ipolates_grib1)(G1_PASS, S_PASS); }
void ipolates_grib1_single_field(G1_ARGS, S_TAIL) { K(ipolates_grib1_single_field)(G1_PASS, S_PASS); }
void ipolatev_grib2(G2_ARGS, V_TAIL) { K(ipolatev_grib2)(G2_PASS, V_PASS); }
void ipolatev_grib2_single_field(G2_ARGS, V_TAIL) { K(ipolatev_grib2_single_field)(G2_PASS, V_PASS); }
void ipolatev_grib1(G1_ARGS, V_TAIL) { K(ipolatev_grib1)(G1_PASS, V_PASS); }
void ipolatev_grib1_single_field(G1_ARGS, V_TAIL) { K(ipolatev_grib1_single_field)(G1_PASS, V_PASS); }

/* The grid expanders are plain external subroutines in the reference (src/ipxwafs.F90:81, ipxwafs2.F90:92,
 * ipxwafs3.F90:95, ipxetas.F90:90), so a Fortran caller links the compiler-mangled names: lower case plus one
 * underscore (gfortran, ifort, nvfortran).  No character arguments, hence no hidden lengths. */
#define XW_HEAD const ip_int *idir, const ip_int *nthin, const ip_int *nfull, const ip_int *km, const ip_int *num_opt, ip_int *opt, \
                const ip_int *len, ip_int *tthin, ip_real *dthin
void ipxwafs_(XW_HEAD, ip_int* tfull, ip_real* dfull, ip_int* iret) {
  K(ipxwafs)(idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, tfull, dfull, iret);
}
void ipxwafs2_(XW_HEAD, ip_int* ibthin, ip_logical* bthin, ip_int* tfull, ip_real* dfull, ip_int* ibfull, ip_logical* bfull, ip_int* iret) {
  K(ipxwafs2)(idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, ibthin, bthin, tfull, dfull, ibfull, bfull, iret);
}
void ipxwafs3_(XW_HEAD, ip_int* ibthin, ip_logical* bthin, ip_int* tfull, ip_real* dfull, ip_int* ibfull, ip_logical* bfull, ip_int* iret) {
  K(ipxwafs3)(idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, ibthin, bthin, tfull, dfull, ibfull, bfull, iret);
}
void ipxetas_(const ip_int* idir, const ip_int* numi, const ip_int* len, const ip_int* tin, const ip_int* npi, const ip_logical* bi,
              const ip_real* di, ip_int* numo, ip_int* tout, const ip_int* npo, ip_logical* bo, ip_real* dout, ip_int* iret) {
  K(ipxetas)(idir, numi, len, tin, npi, bi, di, numo, tout, npo, bo, dout, iret);
}
