/*
 * oracle/shim/proj_api.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Declaration-only stand-in for PROJ.4's classic <proj_api.h> (PROJ 4.4.9 .. 7.x; the reference
 * accepts any version, /root/reference/CMakeLists.txt:44, README.md:16).  PROJ is neither vendored
 * under /root/reference nor installed in this image, so the six symbols that
 * /root/reference/src/interpolation.c needs (:355,:366,:623-644,:1168-1233) are declared here and
 * DEFINED by oracle/pj_oracle.c, a restatement of PROJ 4.9.x's published algorithm for the
 * projections the hot path names.  With this header the reference file compiles unmodified.
 */
#ifndef FIMEX_B200_ORACLE_PROJ_API_SHIM_H
#define FIMEX_B200_ORACLE_PROJ_API_SHIM_H

#include <math.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAD_TO_DEG 57.29577951308232
#define DEG_TO_RAD .0174532925199432958

typedef void* projPJ;

extern int pj_errno;

projPJ pj_init_plus(const char* definition);
int pj_transform(projPJ src, projPJ dst, long point_count, int point_offset, double* x, double* y, double* z);
void pj_free(projPJ pj);
char* pj_strerrno(int err);
int pj_is_latlong(projPJ pj);

#ifdef __cplusplus
}
#endif

#endif
