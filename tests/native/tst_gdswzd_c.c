/* C caller of the drop-in libip_<kind> (shims/ip_shim.c over libipb200.so), with the checks of the reference's own C
 * tests tests/tst_gdswzd_grib2.c and tests/tst_gdswzd_grib1.c: gdswzd / gdswzd_grib1 on a 251 x 201 rotated lat-lon
 * "B" grid, iopt = 1, must return 50451 valid points and corner points within 1 % of (-7.491, -144.134) and
 * (44.539, 14.802).  Also calls ipolates_grib2 through the unsuffixed symbol (1-degree global -> the same B grid,
 * bilinear, one field) and checks a constant field comes back constant.
 * build: gcc -std=c99 -DLSIZE=<4|D|8> -Iinclude tests/native/tst_gdswzd_c.c -L<dir> -lip_<kind> -lipb200 -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "iplib.h"

#define IM 251
#define JM 201

static int corners_ok(const ip_real* rlat, const ip_real* rlon, ip_int nret) {
  const double first_lat = rlat[0], first_lon = rlon[0] - 360.0, last_lat = rlat[nret - 1], last_lon = rlon[nret - 1];
  printf("  first corner %f %f (expected -7.491 -144.134), last corner %f %f (expected 44.539 14.802)\n", first_lat, first_lon, last_lat,
         last_lon);
  return fabs(first_lat + 7.491) / 7.491 <= 0.01 && fabs(first_lon + 144.134) / 144.134 <= 0.01 && fabs(last_lat - 44.539) / 44.539 <= 0.01 &&
         fabs(last_lon - 14.802) / 14.802 <= 0.01;
}

int main(void) {
  const ip_int npts = IM * JM;
  ip_real* a[11];
  for (int q = 0; q < 11; q++) a[q] = (ip_real*)malloc((size_t)npts * sizeof(ip_real));
  ip_real *xpts = a[0], *ypts = a[1], *rlon = a[2], *rlat = a[3];
  for (int j = 0; j < JM; j++)
    for (int i = 0; i < IM; i++) { xpts[j * IM + i] = (ip_real)(i + 1); ypts[j * IM + i] = (ip_real)(j + 1); }

  /* GRIB2 template 3.1, as tests/tst_gdswzd_grib2.c */
  ip_int t[22] = {6, 255, -1, 255, -1, 255, -1, IM, JM, 0, -1, -45036000, 299961000, 56, 45036000, 60039000, 0, 0, 64, -36000000, 254000000, 0};
  ip_int nret = 0;
  gdswzd(1, t, 22, 1, npts, (ip_real)-9999.0, xpts, ypts, rlon, rlat, &nret, a[4], a[5], a[6], a[7], a[8], a[9], a[10]);
  printf("gdswzd: points returned %ld (expected 50451)\n", (long)nret);
  if (nret != 50451 || !corners_ok(rlat, rlon, nret)) return 1;

  /* GRIB1 KGDS, as tests/tst_gdswzd_grib1.c */
  ip_int kgds[200] = {205, IM, JM, -7446, -144139, 8, 54000, -106000, 0, 0, 64, 44560, 14744};
  nret = 0;
  gdswzd_grib1(kgds, 1, npts, (ip_real)-9999.0, xpts, ypts, rlon, rlat, &nret, a[4], a[5], a[6], a[7], a[8], a[9], a[10]);
  printf("gdswzd_grib1: points returned %ld (expected 50451)\n", (long)nret);
  if (nret != 50451 || !corners_ok(rlat, rlon, nret)) return 2;

  /* ipolates through the unsuffixed symbol */
  ip_int tin[19] = {6, 255, -1, 255, -1, 255, -1, 360, 181, 0, -1, 90000000, 0, 48, -90000000, 359000000, 1000000, 1000000, 0};
  const ip_int ip = 0, numi = 0, leni = 19, numo = 1, leno = 22, mi = 360 * 181, mo = npts, km = 1, ibi = 0;
  ip_int ipopt[20] = {0}, no = 0, ibo = -1, iret = -1;
  ip_real* gi = (ip_real*)malloc((size_t)mi * sizeof(ip_real));
  ip_logical* li = (ip_logical*)malloc((size_t)mi);
  ip_logical* lo = (ip_logical*)malloc((size_t)mo);
  for (ip_int n = 0; n < mi; n++) { gi[n] = (ip_real)3.25; li[n] = 1; }
  ipolates_grib2(&ip, ipopt, &numi, tin, &leni, &numo, t, &leno, &mi, &mo, &km, &ibi, li, gi, &no, rlat, rlon, &ibo, lo, a[4], &iret);
  printf("ipolates_grib2: iret %ld no %ld ibo %ld\n", (long)iret, (long)no, (long)ibo);
  if (iret != 0 || no != npts || ibo != 0) return 3;
  for (ip_int n = 0; n < no; n++)
    if (!lo[n] || fabs((double)a[4][n] - 3.25) > 1e-5) { printf("  point %ld: lo %d value %f\n", (long)n, lo[n], (double)a[4][n]); return 4; }
  printf("ok\n");
  return 0;
}
