/* ipb200.h — C ABI of the B200-native ipolates / ipolatev / gdswzd library (libipb200.so).
 *
 * Drop-in boundary for the hot path of NOAA-EMC/NCEPLIBS-ip 5.0.0.  Every interpolation entry point
 * below has exactly the argument list of the reference's own bind(c) driver (all arguments by
 * reference, arrays contiguous in Fortran column-major order, LOGICAL(C_BOOL) = one byte); the two
 * gdswzd entry points have the argument lists of the reference's C wrappers (scalars by value).
 * The reference builds three libraries from one source (src/CMakeLists.txt:36-76); here the build
 * kind is a symbol suffix instead:
 *     _4  INTEGER(4)/REAL(4)  (libip_4)     _d  INTEGER(4)/REAL(8)  (libip_d)     _8  INTEGER(8)/REAL(8)  (libip_8)
 *
 * Array shapes (Fortran):  ipopt(20), kgds*(200), igdtmpl*(igdtlen*), ibi(km), li(mi,km), gi/ui/vi(mi,km),
 * rlat/rlon/crot/srot(mo), ibo(km), lo(mo,km), go/uo/vo(mo,km).  Field k (0-based) of gi starts at gi + k*mi.
 * In/out conventions, iret codes (0 ok, 2 no overlap, 3 bad output grid / mo too small, 32 bad ipopt) and
 * the option words are the reference's (docs in src/ipolates.F90:337-586, src/ipolatev.F90:122-381).
 * Additional codes produced only by this library: iret = 1 for a grid template or method it does not
 * implement or a template shorter than its decode needs (the reference error-stops / reads past the array),
 * iret = 1000 + cudaError for a CUDA failure, iret = 1001 for a host-side failure (out of memory).
 * ip = 4 (spectral) is out of scope.  There is no CPU fallback: without a CUDA device the library aborts.
 *
 * The *_single_field entry points take scalar ibi / ibo and rank-1 li, gi, lo, go; km is ignored (1).
 */
#ifndef IPB200_H
#define IPB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char ipb_bool; /* LOGICAL(C_BOOL) / LOGICAL*1 */

/* ======================== kind _4  (libip_4: default INTEGER(4), REAL(4)) ======================== */
/* replaces ipolates_grib2, src/ipolates.F90:587-636 */
void ipolates_grib2_4(
    const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* gi,
    int32_t* no, float* rlat, float* rlon, int32_t* ibo, ipb_bool* lo, float* go, int32_t* iret);
/* replaces ipolates_grib2_single_field, src/ipolates.F90:808-864 */
void ipolates_grib2_single_field_4(
    const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* gi,
    int32_t* no, float* rlat, float* rlon, int32_t* ibo, ipb_bool* lo, float* go, int32_t* iret);
/* replaces ipolates_grib1, src/ipolates.F90:293-334 */
void ipolates_grib1_4(
    const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* gi,
    int32_t* no, float* rlat, float* rlon, int32_t* ibo, ipb_bool* lo, float* go, int32_t* iret);
/* replaces ipolates_grib1_single_field, src/ipolates.F90:158-206 */
void ipolates_grib1_single_field_4(
    const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* gi,
    int32_t* no, float* rlat, float* rlon, int32_t* ibo, ipb_bool* lo, float* go, int32_t* iret);
/* replaces ipolatev_grib2, src/ipolatev.F90:382-434 */
void ipolatev_grib2_4(
    const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* ui, const float* vi,
    int32_t* no, float* rlat, float* rlon, float* crot, float* srot, int32_t* ibo, ipb_bool* lo, float* uo, float* vo, int32_t* iret);
/* replaces ipolatev_grib2_single_field, src/ipolatev.F90:832-891 */
void ipolatev_grib2_single_field_4(
    const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* ui, const float* vi,
    int32_t* no, float* rlat, float* rlon, float* crot, float* srot, int32_t* ibo, ipb_bool* lo, float* uo, float* vo, int32_t* iret);
/* replaces ipolatev_grib1, src/ipolatev.F90:565-628 (the 203 -> wind-point switch of kgds(11) is applied internally; the caller's kgds is not modified) */
void ipolatev_grib1_4(
    const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* ui, const float* vi,
    int32_t* no, float* rlat, float* rlon, float* crot, float* srot, int32_t* ibo, ipb_bool* lo, float* uo, float* vo, int32_t* iret);
/* replaces ipolatev_grib1_single_field, src/ipolatev.F90:680-750 */
void ipolatev_grib1_single_field_4(
    const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* ui, const float* vi,
    int32_t* no, float* rlat, float* rlon, float* crot, float* srot, int32_t* ibo, ipb_bool* lo, float* uo, float* vo, int32_t* iret);
/* replaces the C symbol gdswzd, src/gdswzd_c.F90:173-210 (prototype src/iplib_4.h:146); optional outputs may be NULL */
void gdswzd_4(
    int32_t igdtnum, const int32_t* igdtmpl, int32_t igdtlen, int32_t iopt, int32_t npts, float fill,
    float* xpts, float* ypts, float* rlon, float* rlat, int32_t* nret,
    float* crot, float* srot, float* xlon, float* xlat, float* ylon, float* ylat, float* area);
/* replaces the C symbol gdswzd_grib1, src/gdswzd_c.F90:259-292 (prototype src/iplib_4.h:187) */
void gdswzd_grib1_4(
    const int32_t* kgds, int32_t iopt, int32_t npts, float fill,
    float* xpts, float* ypts, float* rlon, float* rlat, int32_t* nret,
    float* crot, float* srot, float* xlon, float* xlat, float* ylon, float* ylat, float* area);
/* Grid expanders, every argument by reference like the Fortran routines they replace.
 * ipxwafs (src/ipxwafs.F90:81): thinned WAFS grid (3447 points) <-> full 73 x 73 grid (5329), linear along each row;
 * ipxwafs2 (src/ipxwafs2.F90:92): the same with bitmaps; ipxwafs3 (src/ipxwafs3.F90:95): nearest neighbour with bitmaps.
 * idir > 0 expands, < 0 contracts, 0 only checks; iret 1 = the template is not the WAFS grid.
 * ipxetas (src/ipxetas.F90:90): staggered Eta E grid (scan 68 / 72) <-> filled E grid (scan 64); iret 1 / 2 / 3 as there. */
void ipxwafs_4(
    const int32_t* idir, const int32_t* numpts_thin, const int32_t* numpts_full, const int32_t* km, const int32_t* num_opt, int32_t* opt_pts, const int32_t* igdtlen,
    int32_t* igdtmpl_thin, float* data_thin, int32_t* igdtmpl_full, float* data_full, int32_t* iret);
void ipxwafs2_4(
    const int32_t* idir, const int32_t* numpts_thin, const int32_t* numpts_full, const int32_t* km, const int32_t* num_opt, int32_t* opt_pts, const int32_t* igdtlen,
    int32_t* igdtmpl_thin, float* data_thin, int32_t* ib_thin, ipb_bool* bitmap_thin, int32_t* igdtmpl_full, float* data_full, int32_t* ib_full, ipb_bool* bitmap_full, int32_t* iret);
void ipxwafs3_4(
    const int32_t* idir, const int32_t* numpts_thin, const int32_t* numpts_full, const int32_t* km, const int32_t* num_opt, int32_t* opt_pts, const int32_t* igdtlen,
    int32_t* igdtmpl_thin, float* data_thin, int32_t* ib_thin, ipb_bool* bitmap_thin, int32_t* igdtmpl_full, float* data_full, int32_t* ib_full, ipb_bool* bitmap_full, int32_t* iret);
void ipxetas_4(
    const int32_t* idir, const int32_t* igdtnumi, const int32_t* igdtlen, const int32_t* igdtmpli, const int32_t* npts_input, const ipb_bool* bitmap_input, const float* data_input,
    int32_t* igdtnumo, int32_t* igdtmplo, const int32_t* npts_output, ipb_bool* bitmap_output, float* data_output, int32_t* iret);
/* Device-resident variants (new): li, gi/ui/vi, lo, go/uo/vo are DEVICE pointers (on the calling
 * thread's current device) and so are rlat, rlon, crot, srot (each may be NULL to skip that output; for station-point
 * outputs they are inputs).  Scalars, ipopt, the grid definition, ibi, ibo, no and iret stay host memory.  Work is queued
 * on `stream` (a cudaStream_t) and the call returns after that stream has drained (ibo is a host output); the
 * ipgpu_plan_* entry points below queue without waiting. */
void ipgpu_ipolates_grib2_dev_4(
    void* stream, const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* gi,
    int32_t* no, float* rlat, float* rlon, int32_t* ibo, ipb_bool* lo, float* go, int32_t* iret);
void ipgpu_ipolatev_grib2_dev_4(
    void* stream, const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* ui, const float* vi,
    int32_t* no, float* rlat, float* rlon, float* crot, float* srot, int32_t* ibo, ipb_bool* lo, float* uo, float* vo, int32_t* iret);
/* the same for the GRIB1 interface (kgds arrays of 200 words) */
void ipgpu_ipolates_grib1_dev_4(
    void* stream, const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* gi,
    int32_t* no, float* rlat, float* rlon, int32_t* ibo, ipb_bool* lo, float* go, int32_t* iret);
void ipgpu_ipolatev_grib1_dev_4(
    void* stream, const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* ui, const float* vi,
    int32_t* no, float* rlat, float* rlon, float* crot, float* srot, int32_t* ibo, ipb_bool* lo, float* uo, float* vo, int32_t* iret);
/* Explicit plan handles.  ipgpu_plan_get builds the plan of (grid in, grid out, ip, ipopt) on the
 * current device or fetches it from that device's cache and returns a handle (NULL on failure: iret as ipolates gives it;
 * station-point outputs have no cacheable plan).  *vector != 0 asks for the ipolatev plan.  ipgpu_plan_apply queues the
 * km-field apply stage on `stream` (a cudaStream_t) and RETURNS WITHOUT WAITING: li, gi/ui, vi (NULL for scalars), lo,
 * go/uo, vo are device pointers; ibi is read from the host before the call returns; ibo_dev (km int32 on the device, may
 * be NULL) receives ibo when the stream gets there.  ipgpu_plan_geometry copies rlat/rlon(/crot/srot) to host or device
 * arrays (any may be NULL), queued on `stream`.  Release every handle with ipgpu_plan_release. */
void* ipgpu_plan_get_4(
    const int32_t* vector, const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, int32_t* no, int32_t* iret);
void ipgpu_plan_geometry_4(void* plan, void* stream, float* rlat, float* rlon, float* crot, float* srot, int32_t* iret);
void ipgpu_plan_apply_4(
    void* plan, void* stream, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const float* gi, const float* vi,
    int32_t* ibo_dev, ipb_bool* lo, float* go, float* vo, int32_t* iret);

/* ======================== kind _d  (libip_d: INTEGER(4), REAL(8)) ======================== */
/* replaces ipolates_grib2, src/ipolates.F90:587-636 */
void ipolates_grib2_d(
    const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* gi,
    int32_t* no, double* rlat, double* rlon, int32_t* ibo, ipb_bool* lo, double* go, int32_t* iret);
/* replaces ipolates_grib2_single_field, src/ipolates.F90:808-864 */
void ipolates_grib2_single_field_d(
    const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* gi,
    int32_t* no, double* rlat, double* rlon, int32_t* ibo, ipb_bool* lo, double* go, int32_t* iret);
/* replaces ipolates_grib1, src/ipolates.F90:293-334 */
void ipolates_grib1_d(
    const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* gi,
    int32_t* no, double* rlat, double* rlon, int32_t* ibo, ipb_bool* lo, double* go, int32_t* iret);
/* replaces ipolates_grib1_single_field, src/ipolates.F90:158-206 */
void ipolates_grib1_single_field_d(
    const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* gi,
    int32_t* no, double* rlat, double* rlon, int32_t* ibo, ipb_bool* lo, double* go, int32_t* iret);
/* replaces ipolatev_grib2, src/ipolatev.F90:382-434 */
void ipolatev_grib2_d(
    const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int32_t* no, double* rlat, double* rlon, double* crot, double* srot, int32_t* ibo, ipb_bool* lo, double* uo, double* vo, int32_t* iret);
/* replaces ipolatev_grib2_single_field, src/ipolatev.F90:832-891 */
void ipolatev_grib2_single_field_d(
    const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int32_t* no, double* rlat, double* rlon, double* crot, double* srot, int32_t* ibo, ipb_bool* lo, double* uo, double* vo, int32_t* iret);
/* replaces ipolatev_grib1, src/ipolatev.F90:565-628 (the 203 -> wind-point switch of kgds(11) is applied internally; the caller's kgds is not modified) */
void ipolatev_grib1_d(
    const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int32_t* no, double* rlat, double* rlon, double* crot, double* srot, int32_t* ibo, ipb_bool* lo, double* uo, double* vo, int32_t* iret);
/* replaces ipolatev_grib1_single_field, src/ipolatev.F90:680-750 */
void ipolatev_grib1_single_field_d(
    const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int32_t* no, double* rlat, double* rlon, double* crot, double* srot, int32_t* ibo, ipb_bool* lo, double* uo, double* vo, int32_t* iret);
/* replaces the C symbol gdswzd, src/gdswzd_c.F90:173-210 (prototype src/iplib_d.h:146); optional outputs may be NULL */
void gdswzd_d(
    int32_t igdtnum, const int32_t* igdtmpl, int32_t igdtlen, int32_t iopt, int32_t npts, double fill,
    double* xpts, double* ypts, double* rlon, double* rlat, int32_t* nret,
    double* crot, double* srot, double* xlon, double* xlat, double* ylon, double* ylat, double* area);
/* replaces the C symbol gdswzd_grib1, src/gdswzd_c.F90:259-292 (prototype src/iplib_d.h:187) */
void gdswzd_grib1_d(
    const int32_t* kgds, int32_t iopt, int32_t npts, double fill,
    double* xpts, double* ypts, double* rlon, double* rlat, int32_t* nret,
    double* crot, double* srot, double* xlon, double* xlat, double* ylon, double* ylat, double* area);
/* Grid expanders, every argument by reference like the Fortran routines they replace.
 * ipxwafs (src/ipxwafs.F90:81): thinned WAFS grid (3447 points) <-> full 73 x 73 grid (5329), linear along each row;
 * ipxwafs2 (src/ipxwafs2.F90:92): the same with bitmaps; ipxwafs3 (src/ipxwafs3.F90:95): nearest neighbour with bitmaps.
 * idir > 0 expands, < 0 contracts, 0 only checks; iret 1 = the template is not the WAFS grid.
 * ipxetas (src/ipxetas.F90:90): staggered Eta E grid (scan 68 / 72) <-> filled E grid (scan 64); iret 1 / 2 / 3 as there. */
void ipxwafs_d(
    const int32_t* idir, const int32_t* numpts_thin, const int32_t* numpts_full, const int32_t* km, const int32_t* num_opt, int32_t* opt_pts, const int32_t* igdtlen,
    int32_t* igdtmpl_thin, double* data_thin, int32_t* igdtmpl_full, double* data_full, int32_t* iret);
void ipxwafs2_d(
    const int32_t* idir, const int32_t* numpts_thin, const int32_t* numpts_full, const int32_t* km, const int32_t* num_opt, int32_t* opt_pts, const int32_t* igdtlen,
    int32_t* igdtmpl_thin, double* data_thin, int32_t* ib_thin, ipb_bool* bitmap_thin, int32_t* igdtmpl_full, double* data_full, int32_t* ib_full, ipb_bool* bitmap_full, int32_t* iret);
void ipxwafs3_d(
    const int32_t* idir, const int32_t* numpts_thin, const int32_t* numpts_full, const int32_t* km, const int32_t* num_opt, int32_t* opt_pts, const int32_t* igdtlen,
    int32_t* igdtmpl_thin, double* data_thin, int32_t* ib_thin, ipb_bool* bitmap_thin, int32_t* igdtmpl_full, double* data_full, int32_t* ib_full, ipb_bool* bitmap_full, int32_t* iret);
void ipxetas_d(
    const int32_t* idir, const int32_t* igdtnumi, const int32_t* igdtlen, const int32_t* igdtmpli, const int32_t* npts_input, const ipb_bool* bitmap_input, const double* data_input,
    int32_t* igdtnumo, int32_t* igdtmplo, const int32_t* npts_output, ipb_bool* bitmap_output, double* data_output, int32_t* iret);
/* Device-resident variants (new): li, gi/ui/vi, lo, go/uo/vo are DEVICE pointers (on the calling
 * thread's current device) and so are rlat, rlon, crot, srot (each may be NULL to skip that output; for station-point
 * outputs they are inputs).  Scalars, ipopt, the grid definition, ibi, ibo, no and iret stay host memory.  Work is queued
 * on `stream` (a cudaStream_t) and the call returns after that stream has drained (ibo is a host output); the
 * ipgpu_plan_* entry points below queue without waiting. */
void ipgpu_ipolates_grib2_dev_d(
    void* stream, const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* gi,
    int32_t* no, double* rlat, double* rlon, int32_t* ibo, ipb_bool* lo, double* go, int32_t* iret);
void ipgpu_ipolatev_grib2_dev_d(
    void* stream, const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int32_t* no, double* rlat, double* rlon, double* crot, double* srot, int32_t* ibo, ipb_bool* lo, double* uo, double* vo, int32_t* iret);
/* the same for the GRIB1 interface (kgds arrays of 200 words) */
void ipgpu_ipolates_grib1_dev_d(
    void* stream, const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* gi,
    int32_t* no, double* rlat, double* rlon, int32_t* ibo, ipb_bool* lo, double* go, int32_t* iret);
void ipgpu_ipolatev_grib1_dev_d(
    void* stream, const int32_t* ip, const int32_t* ipopt, const int32_t* kgdsi, const int32_t* kgdso,
    const int32_t* mi, const int32_t* mo, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int32_t* no, double* rlat, double* rlon, double* crot, double* srot, int32_t* ibo, ipb_bool* lo, double* uo, double* vo, int32_t* iret);
/* Explicit plan handles.  ipgpu_plan_get builds the plan of (grid in, grid out, ip, ipopt) on the
 * current device or fetches it from that device's cache and returns a handle (NULL on failure: iret as ipolates gives it;
 * station-point outputs have no cacheable plan).  *vector != 0 asks for the ipolatev plan.  ipgpu_plan_apply queues the
 * km-field apply stage on `stream` (a cudaStream_t) and RETURNS WITHOUT WAITING: li, gi/ui, vi (NULL for scalars), lo,
 * go/uo, vo are device pointers; ibi is read from the host before the call returns; ibo_dev (km int32 on the device, may
 * be NULL) receives ibo when the stream gets there.  ipgpu_plan_geometry copies rlat/rlon(/crot/srot) to host or device
 * arrays (any may be NULL), queued on `stream`.  Release every handle with ipgpu_plan_release. */
void* ipgpu_plan_get_d(
    const int32_t* vector, const int32_t* ip, const int32_t* ipopt, const int32_t* igdtnumi, const int32_t* igdtmpli, const int32_t* igdtleni, const int32_t* igdtnumo, const int32_t* igdtmplo, const int32_t* igdtleno,
    const int32_t* mi, const int32_t* mo, int32_t* no, int32_t* iret);
void ipgpu_plan_geometry_d(void* plan, void* stream, double* rlat, double* rlon, double* crot, double* srot, int32_t* iret);
void ipgpu_plan_apply_d(
    void* plan, void* stream, const int32_t* km, const int32_t* ibi, const ipb_bool* li, const double* gi, const double* vi,
    int32_t* ibo_dev, ipb_bool* lo, double* go, double* vo, int32_t* iret);

/* ======================== kind _8  (libip_8: INTEGER(8), REAL(8)) ======================== */
/* replaces ipolates_grib2, src/ipolates.F90:587-636 */
void ipolates_grib2_8(
    const int64_t* ip, const int64_t* ipopt, const int64_t* igdtnumi, const int64_t* igdtmpli, const int64_t* igdtleni, const int64_t* igdtnumo, const int64_t* igdtmplo, const int64_t* igdtleno,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* gi,
    int64_t* no, double* rlat, double* rlon, int64_t* ibo, ipb_bool* lo, double* go, int64_t* iret);
/* replaces ipolates_grib2_single_field, src/ipolates.F90:808-864 */
void ipolates_grib2_single_field_8(
    const int64_t* ip, const int64_t* ipopt, const int64_t* igdtnumi, const int64_t* igdtmpli, const int64_t* igdtleni, const int64_t* igdtnumo, const int64_t* igdtmplo, const int64_t* igdtleno,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* gi,
    int64_t* no, double* rlat, double* rlon, int64_t* ibo, ipb_bool* lo, double* go, int64_t* iret);
/* replaces ipolates_grib1, src/ipolates.F90:293-334 */
void ipolates_grib1_8(
    const int64_t* ip, const int64_t* ipopt, const int64_t* kgdsi, const int64_t* kgdso,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* gi,
    int64_t* no, double* rlat, double* rlon, int64_t* ibo, ipb_bool* lo, double* go, int64_t* iret);
/* replaces ipolates_grib1_single_field, src/ipolates.F90:158-206 */
void ipolates_grib1_single_field_8(
    const int64_t* ip, const int64_t* ipopt, const int64_t* kgdsi, const int64_t* kgdso,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* gi,
    int64_t* no, double* rlat, double* rlon, int64_t* ibo, ipb_bool* lo, double* go, int64_t* iret);
/* replaces ipolatev_grib2, src/ipolatev.F90:382-434 */
void ipolatev_grib2_8(
    const int64_t* ip, const int64_t* ipopt, const int64_t* igdtnumi, const int64_t* igdtmpli, const int64_t* igdtleni, const int64_t* igdtnumo, const int64_t* igdtmplo, const int64_t* igdtleno,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int64_t* no, double* rlat, double* rlon, double* crot, double* srot, int64_t* ibo, ipb_bool* lo, double* uo, double* vo, int64_t* iret);
/* replaces ipolatev_grib2_single_field, src/ipolatev.F90:832-891 */
void ipolatev_grib2_single_field_8(
    const int64_t* ip, const int64_t* ipopt, const int64_t* igdtnumi, const int64_t* igdtmpli, const int64_t* igdtleni, const int64_t* igdtnumo, const int64_t* igdtmplo, const int64_t* igdtleno,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int64_t* no, double* rlat, double* rlon, double* crot, double* srot, int64_t* ibo, ipb_bool* lo, double* uo, double* vo, int64_t* iret);
/* replaces ipolatev_grib1, src/ipolatev.F90:565-628 (the 203 -> wind-point switch of kgds(11) is applied internally; the caller's kgds is not modified) */
void ipolatev_grib1_8(
    const int64_t* ip, const int64_t* ipopt, const int64_t* kgdsi, const int64_t* kgdso,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int64_t* no, double* rlat, double* rlon, double* crot, double* srot, int64_t* ibo, ipb_bool* lo, double* uo, double* vo, int64_t* iret);
/* replaces ipolatev_grib1_single_field, src/ipolatev.F90:680-750 */
void ipolatev_grib1_single_field_8(
    const int64_t* ip, const int64_t* ipopt, const int64_t* kgdsi, const int64_t* kgdso,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int64_t* no, double* rlat, double* rlon, double* crot, double* srot, int64_t* ibo, ipb_bool* lo, double* uo, double* vo, int64_t* iret);
/* replaces the C symbol gdswzd, src/gdswzd_c.F90:173-210 (prototype src/iplib_8.h:146); optional outputs may be NULL */
void gdswzd_8(
    int64_t igdtnum, const int64_t* igdtmpl, int64_t igdtlen, int64_t iopt, int64_t npts, double fill,
    double* xpts, double* ypts, double* rlon, double* rlat, int64_t* nret,
    double* crot, double* srot, double* xlon, double* xlat, double* ylon, double* ylat, double* area);
/* replaces the C symbol gdswzd_grib1, src/gdswzd_c.F90:259-292 (prototype src/iplib_8.h:187) */
void gdswzd_grib1_8(
    const int64_t* kgds, int64_t iopt, int64_t npts, double fill,
    double* xpts, double* ypts, double* rlon, double* rlat, int64_t* nret,
    double* crot, double* srot, double* xlon, double* xlat, double* ylon, double* ylat, double* area);
/* Grid expanders, every argument by reference like the Fortran routines they replace.
 * ipxwafs (src/ipxwafs.F90:81): thinned WAFS grid (3447 points) <-> full 73 x 73 grid (5329), linear along each row;
 * ipxwafs2 (src/ipxwafs2.F90:92): the same with bitmaps; ipxwafs3 (src/ipxwafs3.F90:95): nearest neighbour with bitmaps.
 * idir > 0 expands, < 0 contracts, 0 only checks; iret 1 = the template is not the WAFS grid.
 * ipxetas (src/ipxetas.F90:90): staggered Eta E grid (scan 68 / 72) <-> filled E grid (scan 64); iret 1 / 2 / 3 as there. */
void ipxwafs_8(
    const int64_t* idir, const int64_t* numpts_thin, const int64_t* numpts_full, const int64_t* km, const int64_t* num_opt, int64_t* opt_pts, const int64_t* igdtlen,
    int64_t* igdtmpl_thin, double* data_thin, int64_t* igdtmpl_full, double* data_full, int64_t* iret);
void ipxwafs2_8(
    const int64_t* idir, const int64_t* numpts_thin, const int64_t* numpts_full, const int64_t* km, const int64_t* num_opt, int64_t* opt_pts, const int64_t* igdtlen,
    int64_t* igdtmpl_thin, double* data_thin, int64_t* ib_thin, ipb_bool* bitmap_thin, int64_t* igdtmpl_full, double* data_full, int64_t* ib_full, ipb_bool* bitmap_full, int64_t* iret);
void ipxwafs3_8(
    const int64_t* idir, const int64_t* numpts_thin, const int64_t* numpts_full, const int64_t* km, const int64_t* num_opt, int64_t* opt_pts, const int64_t* igdtlen,
    int64_t* igdtmpl_thin, double* data_thin, int64_t* ib_thin, ipb_bool* bitmap_thin, int64_t* igdtmpl_full, double* data_full, int64_t* ib_full, ipb_bool* bitmap_full, int64_t* iret);
void ipxetas_8(
    const int64_t* idir, const int64_t* igdtnumi, const int64_t* igdtlen, const int64_t* igdtmpli, const int64_t* npts_input, const ipb_bool* bitmap_input, const double* data_input,
    int64_t* igdtnumo, int64_t* igdtmplo, const int64_t* npts_output, ipb_bool* bitmap_output, double* data_output, int64_t* iret);
/* Device-resident variants (new): li, gi/ui/vi, lo, go/uo/vo are DEVICE pointers (on the calling
 * thread's current device) and so are rlat, rlon, crot, srot (each may be NULL to skip that output; for station-point
 * outputs they are inputs).  Scalars, ipopt, the grid definition, ibi, ibo, no and iret stay host memory.  Work is queued
 * on `stream` (a cudaStream_t) and the call returns after that stream has drained (ibo is a host output); the
 * ipgpu_plan_* entry points below queue without waiting. */
void ipgpu_ipolates_grib2_dev_8(
    void* stream, const int64_t* ip, const int64_t* ipopt, const int64_t* igdtnumi, const int64_t* igdtmpli, const int64_t* igdtleni, const int64_t* igdtnumo, const int64_t* igdtmplo, const int64_t* igdtleno,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* gi,
    int64_t* no, double* rlat, double* rlon, int64_t* ibo, ipb_bool* lo, double* go, int64_t* iret);
void ipgpu_ipolatev_grib2_dev_8(
    void* stream, const int64_t* ip, const int64_t* ipopt, const int64_t* igdtnumi, const int64_t* igdtmpli, const int64_t* igdtleni, const int64_t* igdtnumo, const int64_t* igdtmplo, const int64_t* igdtleno,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int64_t* no, double* rlat, double* rlon, double* crot, double* srot, int64_t* ibo, ipb_bool* lo, double* uo, double* vo, int64_t* iret);
/* the same for the GRIB1 interface (kgds arrays of 200 words) */
void ipgpu_ipolates_grib1_dev_8(
    void* stream, const int64_t* ip, const int64_t* ipopt, const int64_t* kgdsi, const int64_t* kgdso,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* gi,
    int64_t* no, double* rlat, double* rlon, int64_t* ibo, ipb_bool* lo, double* go, int64_t* iret);
void ipgpu_ipolatev_grib1_dev_8(
    void* stream, const int64_t* ip, const int64_t* ipopt, const int64_t* kgdsi, const int64_t* kgdso,
    const int64_t* mi, const int64_t* mo, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* ui, const double* vi,
    int64_t* no, double* rlat, double* rlon, double* crot, double* srot, int64_t* ibo, ipb_bool* lo, double* uo, double* vo, int64_t* iret);
/* Explicit plan handles.  ipgpu_plan_get builds the plan of (grid in, grid out, ip, ipopt) on the
 * current device or fetches it from that device's cache and returns a handle (NULL on failure: iret as ipolates gives it;
 * station-point outputs have no cacheable plan).  *vector != 0 asks for the ipolatev plan.  ipgpu_plan_apply queues the
 * km-field apply stage on `stream` (a cudaStream_t) and RETURNS WITHOUT WAITING: li, gi/ui, vi (NULL for scalars), lo,
 * go/uo, vo are device pointers; ibi is read from the host before the call returns; ibo_dev (km int32 on the device, may
 * be NULL) receives ibo when the stream gets there.  ipgpu_plan_geometry copies rlat/rlon(/crot/srot) to host or device
 * arrays (any may be NULL), queued on `stream`.  Release every handle with ipgpu_plan_release. */
void* ipgpu_plan_get_8(
    const int64_t* vector, const int64_t* ip, const int64_t* ipopt, const int64_t* igdtnumi, const int64_t* igdtmpli, const int64_t* igdtleni, const int64_t* igdtnumo, const int64_t* igdtmplo, const int64_t* igdtleno,
    const int64_t* mi, const int64_t* mo, int64_t* no, int64_t* iret);
void ipgpu_plan_geometry_8(void* plan, void* stream, double* rlat, double* rlon, double* crot, double* srot, int64_t* iret);
void ipgpu_plan_apply_8(
    void* plan, void* stream, const int64_t* km, const int64_t* ibi, const ipb_bool* li, const double* gi, const double* vi,
    int32_t* ibo_dev, ipb_bool* lo, double* go, double* vo, int64_t* iret);

/* ======================== library management (new; no reference counterpart) ======================== */
int ipgpu_device_count(void);                     /* CUDA devices visible to the process */
int ipgpu_set_device(int device);                 /* cudaSetDevice for the calling thread; every call works on the device that is current when it enters */
int ipgpu_set_devices(int n);                     /* host-pointer calls shard their field batch over the first n devices (n <= 1: the current device); returns the count in use */
void ipgpu_plan_release(void* plan);              /* drops a handle of ipgpu_plan_get_<k> (the plan stays cached) */
void ipgpu_plan_clear(void);                      /* drop every cached plan and the staging buffers, on every device */
long long ipgpu_plan_count(void);                 /* plans resident in HBM */
long long ipgpu_plan_bytes(void);                 /* bytes they hold */
void ipgpu_set_plan_cache_bytes(long long bytes); /* LRU capacity of the plan cache (default 24 GiB) */
void ipgpu_set_chunk_bytes(long long bytes);      /* staging size per in-flight chunk of the host-pointer API (default 768 MiB) */
long long ipgpu_launch_count(void);               /* kernels launched by this library since load */
void ipgpu_last_timing(double* plan_ms, double* apply_ms, int* plan_cache_hit); /* CUDA-event times of the last call */
void ipgpu_last_transfer(long long* h2d_bytes, long long* d2h_bytes); /* field bytes the last host-pointer call moved over PCIe */
void* ipgpu_host_alloc(long long bytes);          /* pinned host memory on the NUMA node of the current device's PCIe link, so the host API copies asynchronously at link rate */
void ipgpu_host_free(void* p);
int ipgpu_host_register(void* p, long long bytes); /* page-lock an array the caller owns (Fortran allocatable, malloc): copied directly at link rate afterwards; 0 or the cudaError */
int ipgpu_host_unregister(void* p);
int ipgpu_device_numa_node(void);                 /* NUMA node of the current device's PCIe link, -1 unknown */
/* host half of the host-pointer pipeline, no GPU needed (tests): what = 0 rebuilds lo (dst, pitch dpitch bytes per field, the first
 * `width` points of each of kn fields) from packed bits (src: kn rows of ceil(width/32) words, bit i of word w = point 32 w + i) and
 * hasfalse[kn]; what = 1 copies kn rows of `width` bytes between pitched arrays.  Returns 0. */
int ipgpu_selftest_host(int what, int threads, unsigned char* dst, long long dpitch, const unsigned char* src, long long spitch,
                        long long width, long long kn, const int* hasfalse);
void ipgpu_link_probe(long long bytes, double* h2d_gbs, double* d2h_gbs, double* both_gbs); /* measured host-link ceiling (GB/s) with that allocator */
const char* ipgpu_version(void);
/* tile geometry of the point-staged kernels (host-side, for tests): mode 0 = 128 consecutive points, 1 = 32 x 4 patch,
   2 = 16 x 8 patch of an output grid whose rows hold `rowlen` points; the point of (tile, r < 4, lane < 32) or -1 */
int ipgpu_tile_point(int mode, int rowlen, int no, int tile, int r, int lane);
int ipgpu_tile_count(int mode, int rowlen, int no);

#ifdef __cplusplus
}
#endif
#endif /* IPB200_H */
