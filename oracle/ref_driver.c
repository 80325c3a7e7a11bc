/*
 * oracle/ref_driver.c -- TEST / BASELINE INFRASTRUCTURE ONLY.  Linked into oracle/_ref/libmifi_ref.so next to the
 * reference's own, unmodified src/interpolation.c.
 *
 * The reference's per-slice loop lives in C++ (CachedInterpolation::interpolateValues,
 * /root/reference/src/CachedInterpolation.cc:118-147) and cannot be compiled here (Boost).  This file restates
 * exactly that loop -- "#pragma omp parallel" with a per-thread zValues buffer, "#pragma omp for" over the
 * target points, z-strided scatter -- around the REFERENCE's function pointers mifi_get_values_f /
 * _bilinear_f / _bicubic_f, so that the CPU baseline times the reference's own kernels.
 */
#include <stdlib.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

extern int mifi_get_values_f(const float*, float*, const double, const double, const int, const int, const int);
extern int mifi_get_values_bilinear_f(const float*, float*, const double, const double, const int, const int, const int);
extern int mifi_get_values_bicubic_f(const float*, float*, const double, const double, const int, const int, const int);
extern int mifi_vector_reproject_values_by_matrix_f(int, const double*, float*, float*, int, int, int);

int ref_cached_interpolate(int method, const double* px, const double* py, size_t inX, size_t inY, size_t outX, size_t outY,
                           const float* in, size_t size, float* out, int nthreads)
{
    int (*func)(const float*, float*, const double, const double, const int, const int, const int);
    switch (method) { /* constructor switch, CachedInterpolation.cc:108-115 */
    case 1: func = mifi_get_values_bilinear_f; break;
    case 2: func = mifi_get_values_bicubic_f; break;
    case 0: case 3: case 4: func = mifi_get_values_f; break;
    default: return -1;
    }
    const size_t outLayerSize = outX * outY;
    const size_t inZ = size / (inX * inY);
    int failed = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel default(shared) num_threads(nthreads)
    {
#endif
        float* zValues = (float*)malloc(sizeof(float) * (inZ ? inZ : 1));
#ifdef _OPENMP
#pragma omp for
#endif
        for (long long xy = 0; xy < (long long)outLayerSize; ++xy) {
            float* outPos = &out[xy];
            if (func(in, zValues, px[xy], py[xy], (int)inX, (int)inY, (int)inZ) != -1) {
                for (size_t z = 0; z < inZ; ++z) {
                    *outPos = zValues[z];
                    outPos += outLayerSize;
                }
            } else {
                failed = 1;
            }
        }
        free(zValues);
#ifdef _OPENMP
    }
#endif
    (void)nthreads;
    return failed ? -1 : 1;
}

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
