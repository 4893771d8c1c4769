/* iplib.h — C prototypes of the drop-in libraries libip_4 / libip_d / libip_8 (shims/ip_shim.c over libipb200.so).
 *
 * Same role, names and argument lists as the reference's src/iplib.h + iplib_4.h / iplib_d.h / iplib_8.h (gdswzd and
 * gdswzd_grib1, src/iplib_4.h:146-191) and as its bind(c) Fortran drivers (src/ipolates.F90:158,293,587,808,
 * src/ipolatev.F90:382,565,680,832).  Select the build kind with -DLSIZE=4 (int / float), -DLSIZE=D or nothing
 * (int / double), -DLSIZE=8 (long / double) and link the matching libip_<kind>. */
#ifndef IPLIB_H
#define IPLIB_H
#include <stdint.h>

#if defined(LSIZE) && (LSIZE == 4)
typedef int32_t ip_int;
typedef float ip_real;
#elif defined(LSIZE) && (LSIZE == 8)
typedef int64_t ip_int;
typedef double ip_real;
#else
typedef int32_t ip_int;
typedef double ip_real;
#endif
typedef unsigned char ip_logical; /* LOGICAL(C_BOOL) */

#ifdef __cplusplus
extern "C" {
#endif

/* coordinate transforms: scalars by value (src/gdswzd_c.F90:173-210, :259-292); optional outputs may be NULL */
void gdswzd(ip_int igdtnum, ip_int* igdtmpl, ip_int igdtlen, ip_int iopt, ip_int npts, ip_real fill, ip_real* xpts, ip_real* ypts,
            ip_real* rlon, ip_real* rlat, ip_int* nret, ip_real* crot, ip_real* srot, ip_real* xlon, ip_real* xlat, ip_real* ylon,
            ip_real* ylat, ip_real* area);
void gdswzd_grib1(ip_int* kgds, ip_int iopt, ip_int npts, ip_real fill, ip_real* xpts, ip_real* ypts, ip_real* rlon, ip_real* rlat,
                  ip_int* nret, ip_real* crot, ip_real* srot, ip_real* xlon, ip_real* xlat, ip_real* ylon, ip_real* ylat, ip_real* area);

/* interpolation drivers: every argument by reference, Fortran array layout */
void ipolates_grib2(const ip_int* ip, const ip_int* ipopt, const ip_int* igdtnumi, const ip_int* igdtmpli, const ip_int* igdtleni,
                    const ip_int* igdtnumo, const ip_int* igdtmplo, const ip_int* igdtleno, const ip_int* mi, const ip_int* mo,
                    const ip_int* km, const ip_int* ibi, const ip_logical* li, const ip_real* gi, ip_int* no, ip_real* rlat,
                    ip_real* rlon, ip_int* ibo, ip_logical* lo, ip_real* go, ip_int* iret);
void ipolates_grib2_single_field(const ip_int* ip, const ip_int* ipopt, const ip_int* igdtnumi, const ip_int* igdtmpli,
                                 const ip_int* igdtleni, const ip_int* igdtnumo, const ip_int* igdtmplo, const ip_int* igdtleno,
                                 const ip_int* mi, const ip_int* mo, const ip_int* km, const ip_int* ibi, const ip_logical* li,
                                 const ip_real* gi, ip_int* no, ip_real* rlat, ip_real* rlon, ip_int* ibo, ip_logical* lo, ip_real* go,
                                 ip_int* iret);
void ipolates_grib1(const ip_int* ip, const ip_int* ipopt, const ip_int* kgdsi, const ip_int* kgdso, const ip_int* mi, const ip_int* mo,
                    const ip_int* km, const ip_int* ibi, const ip_logical* li, const ip_real* gi, ip_int* no, ip_real* rlat,
                    ip_real* rlon, ip_int* ibo, ip_logical* lo, ip_real* go, ip_int* iret);
void ipolates_grib1_single_field(const ip_int* ip, const ip_int* ipopt, const ip_int* kgdsi, const ip_int* kgdso, const ip_int* mi,
                                 const ip_int* mo, const ip_int* km, const ip_int* ibi, const ip_logical* li, const ip_real* gi,
                                 ip_int* no, ip_real* rlat, ip_real* rlon, ip_int* ibo, ip_logical* lo, ip_real* go, ip_int* iret);
void ipolatev_grib2(const ip_int* ip, const ip_int* ipopt, const ip_int* igdtnumi, const ip_int* igdtmpli, const ip_int* igdtleni,
                    const ip_int* igdtnumo, const ip_int* igdtmplo, const ip_int* igdtleno, const ip_int* mi, const ip_int* mo,
                    const ip_int* km, const ip_int* ibi, const ip_logical* li, const ip_real* ui, const ip_real* vi, ip_int* no,
                    ip_real* rlat, ip_real* rlon, ip_real* crot, ip_real* srot, ip_int* ibo, ip_logical* lo, ip_real* uo, ip_real* vo,
                    ip_int* iret);
void ipolatev_grib2_single_field(const ip_int* ip, const ip_int* ipopt, const ip_int* igdtnumi, const ip_int* igdtmpli,
                                 const ip_int* igdtleni, const ip_int* igdtnumo, const ip_int* igdtmplo, const ip_int* igdtleno,
                                 const ip_int* mi, const ip_int* mo, const ip_int* km, const ip_int* ibi, const ip_logical* li,
                                 const ip_real* ui, const ip_real* vi, ip_int* no, ip_real* rlat, ip_real* rlon, ip_real* crot,
                                 ip_real* srot, ip_int* ibo, ip_logical* lo, ip_real* uo, ip_real* vo, ip_int* iret);
void ipolatev_grib1(const ip_int* ip, const ip_int* ipopt, const ip_int* kgdsi, const ip_int* kgdso, const ip_int* mi, const ip_int* mo,
                    const ip_int* km, const ip_int* ibi, const ip_logical* li, const ip_real* ui, const ip_real* vi, ip_int* no,
                    ip_real* rlat, ip_real* rlon, ip_real* crot, ip_real* srot, ip_int* ibo, ip_logical* lo, ip_real* uo, ip_real* vo,
                    ip_int* iret);
void ipolatev_grib1_single_field(const ip_int* ip, const ip_int* ipopt, const ip_int* kgdsi, const ip_int* kgdso, const ip_int* mi,
                                 const ip_int* mo, const ip_int* km, const ip_int* ibi, const ip_logical* li, const ip_real* ui,
                                 const ip_real* vi, ip_int* no, ip_real* rlat, ip_real* rlon, ip_real* crot, ip_real* srot, ip_int* ibo,
                                 ip_logical* lo, ip_real* uo, ip_real* vo, ip_int* iret);

#ifdef __cplusplus
}
#endif
#endif /* IPLIB_H */
