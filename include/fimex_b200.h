/*
 * fimex_b200.h -- C ABI of libfimex_b200.so: Fimex's horizontal-regridding hot path on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  Two groups of entry points:
 *
 *  (1) HANDLE API -- what the three C++ classes of the reference forward to after a one-line patch
 *      (INTEGRATION.md):  CachedInterpolation / CachedForwardInterpolation (both behind
 *      CachedInterpolationInterface::interpolateValues) and CachedVectorReprojection::reprojectValues.
 *  (2) mifi_* FUNCTIONS -- the reference's own C symbols for this path, same names, same signatures, same
 *      MIFI_OK / MIFI_ERROR convention (reference include/fimex/interpolation.h), so a program linked
 *      against libfimex's C interface can link against this library for them instead.
 *
 * All file:line citations are relative to the reference tree (arebru/fimex 0.67.2).
 *
 * Conventions:
 *  - plain pointers and sizes only; fp32 data is C-ordered [z][y][x], x fastest
 *    (mifi_3d_array_position, include/fimex/interpolation.h:423-426); tables are fp64 [outY][outX]
 *    (backward methods) or [inY][inX] (forward methods); the rotation matrix is fp64 [oy][ox][4]
 *    = (cos, sin, -sin, phi) (src/interpolation.c:429-432).
 *  - the caller owns every host buffer; nothing is retained after a call returns.
 *  - return MIFI_OK (1) or MIFI_ERROR (-1); the message is available from fb200_last_error() (per
 *    thread) and is also printed to stderr like the reference does; no C++ exception crosses the ABI.
 *  - handles are immutable after creation (createReducedDomain is part of creation) and every
 *    interpolate/reproject call is re-entrant: concurrent calls on one handle from several host threads are
 *    safe (the reference calls them from OpenMP tasks, src/NetCDF_CDMWriter.cc:749-753).
 *  - "_device" variants take DEVICE pointers (and a cudaStream_t passed as void*, NULL = per-thread
 *    default stream), do no host<->device copies and do not synchronise.
 *  - there is NO CPU fallback: every call fails with MIFI_ERROR when no CUDA device is usable.
 *  - device selection: environment variable FIMEX_B200_DEVICE (default: the current CUDA device) or
 *    fb200_set_device(); no new user option is needed (the fimex --interpolate.* options are untouched).
 */
#ifndef FIMEX_B200_H_
#define FIMEX_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MIFI_OK
#define MIFI_ERROR -1 /* include/fimex/mifi_constants.h:259 */
#define MIFI_OK 1     /* include/fimex/mifi_constants.h:261 */
#endif
#ifndef MIFI_PROJ_AXIS
#define MIFI_PROJ_AXIS 0 /* include/fimex/mifi_constants.h:264-268 */
#define MIFI_LONGITUDE 1
#define MIFI_LATITUDE 2
#endif
#ifndef MIFI_VECTOR_KEEP_SIZE
#define MIFI_VECTOR_KEEP_SIZE 0 /* include/fimex/mifi_constants.h:154-160 */
#define MIFI_VECTOR_RESIZE 1
#endif

/* enum mifi_interpol_method, include/fimex/mifi_constants.h:52-147 (same numbering) */
#ifndef FIMEX_B200_NO_METHOD_ENUM
enum fb200_interpol_method {
    FB200_INTERPOL_UNKNOWN = -1,
    FB200_INTERPOL_NEAREST_NEIGHBOR = 0,
    FB200_INTERPOL_BILINEAR,
    FB200_INTERPOL_BICUBIC,
    FB200_INTERPOL_COORD_NN,
    FB200_INTERPOL_COORD_NN_KD,
    FB200_INTERPOL_FORWARD_SUM,
    FB200_INTERPOL_FORWARD_MEAN,
    FB200_INTERPOL_FORWARD_MEDIAN,
    FB200_INTERPOL_FORWARD_MAX,
    FB200_INTERPOL_FORWARD_MIN,
    FB200_INTERPOL_FORWARD_UNDEF_SUM,
    FB200_INTERPOL_FORWARD_UNDEF_MEAN,
    FB200_INTERPOL_FORWARD_UNDEF_MEDIAN,
    FB200_INTERPOL_FORWARD_UNDEF_MAX,
    FB200_INTERPOL_FORWARD_UNDEF_MIN
};
#endif

/* enum CDMDataType, include/fimex/CDMDataType.h:35-49 (same numbering) */
enum fb200_datatype {
    FB200_NAT = 0,
    FB200_CHAR,
    FB200_SHORT,
    FB200_INT,
    FB200_FLOAT,
    FB200_DOUBLE,
    FB200_STRING,
    FB200_UCHAR,
    FB200_USHORT,
    FB200_UINT,
    FB200_INT64,
    FB200_UINT64
};

/* ------------------------------------------------------------------------------------------------
 * library state
 * ---------------------------------------------------------------------------------------------- */
const char* fb200_version(void);
/* message of the last failed call made by the calling thread ("" if none) */
const char* fb200_last_error(void);
/* choose the CUDA device for handles created afterwards by this thread (default: $FIMEX_B200_DEVICE or current) */
int fb200_set_device(int device);
int fb200_get_device(void);
/* number of kernels this library has launched so far in the process (bench.py's gpu_launches) */
unsigned long long fb200_kernel_launches(void);
/* Page-locked host buffers for full-rate PCIe copies in the host-buffer calls (pageable buffers work too, at a third of
 * the rate).  Freed buffers stay page-locked in a size-keyed cache and are handed out again, because locking pages costs
 * about as much as copying them and the reference allocates a new output array per interpolateValues call
 * (src/CachedInterpolation.cc:123).  The cache keeps at most FIMEX_B200_PINNED_CACHE_MB (parsed once; default: an eighth of
 * the physical RAM, at most 8192 MB; 0 disables caching); fb200_host_trim releases it.  The buffers are portable
 * (usable from every device); the calls never change the calling thread's current device.  Thread-safe. */
void* fb200_host_alloc(size_t bytes);
void fb200_host_free(void* p);
void fb200_host_trim(void);

/* ------------------------------------------------------------------------------------------------
 * (1a) CachedInterpolationInterface: CachedInterpolation and CachedForwardInterpolation
 * ---------------------------------------------------------------------------------------------- */
typedef struct fb200_interp fb200_interp; /* opaque */

/* CachedInterpolation::CachedInterpolation(xDimName, yDimName, funcType, pointsOnXAxis, pointsOnYAxis, inX, inY,
 * outX, outY) -- include/fimex/CachedInterpolation.h:126-128, src/CachedInterpolation.cc:93-116.
 * funcType: NEAREST_NEIGHBOR, BILINEAR, BICUBIC, COORD_NN, COORD_NN_KD (the last two gather like NN).
 * pointsOn?Axis: outX*outY fractional source positions (host memory).  The dimension names are host
 * metadata and stay with the caller. */
int fb200_cached_interpolation_create(int funcType, const double* pointsOnXAxis, const double* pointsOnYAxis, size_t inX, size_t inY,
                                      size_t outX, size_t outY, fb200_interp** handle);
/* same, positions already on the device (used when the tables are computed or broadcast on the GPU) */
int fb200_cached_interpolation_create_device(int funcType, const double* d_pointsOnXAxis, const double* d_pointsOnYAxis, size_t inX,
                                             size_t inY, size_t outX, size_t outY, fb200_interp** handle);

/* The table-producing part of CDMInterpolator::changeProjectionByProjectionParameters
 * (src/CDMInterpolator.cc:1440-1481) on the device: target axes -> source CRS (mifi_project_axes :1458) ->
 * fractional source positions (mifi_points2position x2, :1475-1476) -> CachedInterpolation.
 * Axis values are in the projection's unit or DEGREES; *_is_degree mirrors the reference's unit regex
 * ".*degree.*" (:1443-1451) for the target and Projection::isDegree() (:1468-1473) for the source. */
int fb200_cached_interpolation_create_from_projection(int funcType, const char* proj_target, const double* out_x_axis,
                                                      const double* out_y_axis, size_t outX, size_t outY, int out_x_is_degree,
                                                      int out_y_is_degree, const char* proj_source, const double* in_x_axis,
                                                      const double* in_y_axis, size_t inX, size_t inY, int in_is_degree,
                                                      fb200_interp** handle);

/* The table-producing part of CDMInterpolator::changeProjectionByProjectionParametersToLatLonTemplate
 * (src/CDMInterpolator.cc:1755-1803): the target is a 2-D lon/lat template (DEGREES, outX*outY values each, any curvilinear
 * grid or a 1-D point list with outY = 1) in CRS proj_template (normally MIFI_WGS84_LATLON_PROJ4): deg -> rad,
 * mifi_project_values(template -> source) (:1770), mifi_points2position on both source axes (:1789-1790), CachedInterpolation. */
int fb200_cached_interpolation_create_from_template(int funcType, const char* proj_template, const double* tmplLon, const double* tmplLat,
                                                    size_t outX, size_t outY, const char* proj_source, const double* in_x_axis,
                                                    const double* in_y_axis, size_t inX, size_t inY, int in_is_degree, fb200_interp** handle);

/* The table-producing part of CDMInterpolator::changeProjectionByCoordinates for MIFI_INTERPOL_COORD_NN
 * (src/CDMInterpolator.cc:1387-1412 with fastTranslatePointsToClosestInputCell :1141-1220): target axes ->
 * WGS84 lat/lon -> nearest source cell by great-circle distance.  lon2d/lat2d: source coordinates in
 * DEGREES, inX*inY values each, index ix + iy*inX. */
int fb200_cached_interpolation_create_from_coordinates(int funcType, const char* proj_target, const double* out_x_axis,
                                                       const double* out_y_axis, size_t outX, size_t outY, int out_x_is_degree,
                                                       int out_y_is_degree, const double* lon2d, const double* lat2d, size_t inX,
                                                       size_t inY, fb200_interp** handle);

/* The same for MIFI_INTERPOL_COORD_NN_KD (coord_kdtree; flannTranslatePointsToClosestInputCell, src/CDMInterpolator.cc:991-1062):
 * the source point with the smallest squared chord distance on the unit sphere inside (maxDistance / 6371000)^2, else
 * (-1000, -1000).  maxDistance (metres) > 0 is the reference's setDistanceOfInterest; <= 0 derives it from the output
 * axes like getMaxDistanceOfInterest (:304-326).  funcType COORD_NN is accepted too (maxDistance is ignored then), and
 * fb200_cached_interpolation_create_from_coordinates accepts COORD_NN_KD with the derived distance. */
int fb200_cached_interpolation_create_from_coordinates_kd(int funcType, const char* proj_target, const double* out_x_axis,
                                                          const double* out_y_axis, size_t outX, size_t outY, int out_x_is_degree,
                                                          int out_y_is_degree, const double* lon2d, const double* lat2d, size_t inX,
                                                          size_t inY, double maxDistance, fb200_interp** handle);

/* CachedForwardInterpolation::CachedForwardInterpolation(..., funcType, pOnX, pOnY, inX, inY, outX, outY) --
 * src/CachedForwardInterpolation.h:51-53, src/CachedForwardInterpolation.cc:62-90.  funcType: FORWARD_*.
 * pOnX/pOnY: inX*inY fractional TARGET positions of every source point. */
int fb200_cached_forward_interpolation_create(int funcType, const double* pOnX, const double* pOnY, size_t inX, size_t inY, size_t outX,
                                              size_t outY, fb200_interp** handle);
/* The table-producing part of CDMInterpolator::changeProjectionByForwardInterpolation
 * (src/CDMInterpolator.cc:1289-1332): source lon/lat (DEGREES, inX*inY each) -> target CRS
 * (mifi_project_values :1311) -> positions on the target axes (:1316-1317) -> CachedForwardInterpolation. */
int fb200_cached_forward_interpolation_create_from_coordinates(int funcType, const char* proj_target, const double* out_x_axis,
                                                               const double* out_y_axis, size_t outX, size_t outY, int out_x_is_degree,
                                                               int out_y_is_degree, const double* lon2d, const double* lat2d,
                                                               size_t inX, size_t inY, fb200_interp** handle);

/* CachedInterpolation::createReducedDomain -- src/CachedInterpolation.cc:159-200.  Crops the expected input to
 * the target footprint (+2 cells).  Returns MIFI_OK; *reduced tells whether a crop was applied, in which case
 * xMin and yMin receive the offsets into the original source grid and fb200_interp_in_x/y() shrink.
 * Must be called before the handle is shared between threads (it is part of construction). */
int fb200_interp_create_reduced_domain(fb200_interp* handle, int* reduced, long long* xMin, long long* yMin);

size_t fb200_interp_in_x(const fb200_interp* handle);  /* getInX()  */
size_t fb200_interp_in_y(const fb200_interp* handle);  /* getInY()  */
size_t fb200_interp_out_x(const fb200_interp* handle); /* getOutX() */
size_t fb200_interp_out_y(const fb200_interp* handle); /* getOutY() */
int fb200_interp_method(const fb200_interp* handle);
/* copy the fractional positions (after any crop) to host arrays / expose the device arrays */
int fb200_interp_get_points(const fb200_interp* handle, double* pointsOnXAxis, double* pointsOnYAxis);
int fb200_interp_device_points(const fb200_interp* handle, const double** d_pointsOnXAxis, const double** d_pointsOnYAxis, size_t* n);

/* CachedInterpolationInterface::interpolateValues(inData, size, newSize) -- include/fimex/CachedInterpolation.h:66,
 * src/CachedInterpolation.cc:118-147 and src/CachedForwardInterpolation.cc:92-131.
 * inData: `size` floats = [inZ][inY][inX]; outData: caller-allocated, outX*outY*inZ floats (query with
 * fb200_interp_new_size); *newSize receives that count.  Host buffers; copies are pipelined with the kernels. */
size_t fb200_interp_new_size(const fb200_interp* handle, size_t size);
int fb200_interp_interpolate_values(const fb200_interp* handle, const float* inData, size_t size, float* outData, size_t* newSize);
int fb200_interp_interpolate_values_device(const fb200_interp* handle, const float* d_inData, size_t size, float* d_outData,
                                           size_t* newSize, void* cuda_stream);
void fb200_interp_destroy(fb200_interp* handle);

/* ------------------------------------------------------------------------------------------------
 * (1b) CachedVectorReprojection
 * ---------------------------------------------------------------------------------------------- */
typedef struct fb200_vector fb200_vector; /* opaque */

/* CachedVectorReprojection(int method, shared_array<double> matrix, int ox, int oy) --
 * include/fimex/CachedVectorReprojection.h:36-44.  matrix: ox*oy*4 doubles (host). */
int fb200_vector_create(int method, const double* matrix, int ox, int oy, fb200_vector** handle);
/* mifi_get_vector_reproject_matrix (src/interpolation.c:719-788) evaluated on the device, as called from
 * src/CDMInterpolator.cc:1492-1500; axis values in the projection's unit or degrees (axis types as there) */
int fb200_vector_create_from_projection(int method, const char* proj_input, const char* proj_output, const double* out_x_axis,
                                        const double* out_y_axis, int out_x_axis_type, int out_y_axis_type, int ox, int oy,
                                        fb200_vector** handle);
/* mifi_get_vector_reproject_matrix_points (src/interpolation.c:719-788) on the device for a list of `on` points given as
 * lon/lat in DEGREES, as the lat/lon-template path builds it (src/CDMInterpolator.cc:1805-1823: proj_output =
 * MIFI_WGS84_LATLON_PROJ4, inputIsMetric = !isDegree(source)); the handle has ox = on, oy = 1 */
int fb200_vector_create_from_points(int method, const char* proj_input, const char* proj_output, int inputIsMetric, const double* lon,
                                    const double* lat, int on, fb200_vector** handle);
/* makeCachedVectorReprojection(dataReader, cs, toLatLon) (src/CDMProcessor.cc:99-145), the matrix behind
 * CDMProcessor::rotateVectorToLatLon / rotateDirectionToLatLon: the grid's OWN axes (degrees when isDegree, else metres).
 * toLatLon != 0: grid directions -> geographic (mifi_get_vector_reproject_matrix_field on the expanded mesh, :124-135);
 * toLatLon == 0: geographic -> grid directions (mifi_get_vector_reproject_matrix from MIFI_WGS84_LATLON_PROJ4, :137-140). */
int fb200_vector_create_from_grid(int method, const char* proj, const double* x_axis, const double* y_axis, int nx, int ny, int isDegree,
                                  int toLatLon, fb200_vector** handle);
/* reprojectValues(uValues, vValues, size): rotate in place -- src/CachedVectorReprojection.cc:35-44 */
int fb200_vector_reproject_values(const fb200_vector* handle, float* uValues, float* vValues, size_t size);
int fb200_vector_reproject_values_device(const fb200_vector* handle, float* d_uValues, float* d_vValues, size_t size, void* cuda_stream);
/* reprojectDirectionValues(angles, size) -- src/CachedVectorReprojection.cc:46-55 */
int fb200_vector_reproject_direction_values(const fb200_vector* handle, float* angles, size_t size);
int fb200_vector_get_matrix(const fb200_vector* handle, double* matrix);
/* The rotation branch of CDMProcessor::getDataSlice (src/CDMProcessor.cc:579-617) for one x/y pair: both components
 * fill -> NaN as float (data2InterpolationArray), reprojectValues, NaN -> fill and cast back (interpolationArray2Data).
 * inType/outType: enum fb200_datatype; `size` values per component; the caller keeps the component it asked for. */
int fb200_vector_get_slice(const fb200_vector* handle, int inType, const void* uIn, const void* vIn, size_t size, double badU, double badV,
                           int outType, void* uOut, void* vOut);
int fb200_vector_get_slice_device(const fb200_vector* handle, int inType, const void* d_uIn, const void* d_vIn, size_t size, double badU,
                                  double badV, int outType, void* d_uOut, void* d_vOut, void* cuda_stream);
void fb200_vector_destroy(fb200_vector* handle);

/* Fused form of src/CDMInterpolator.cc:255-276: interpolate both components of an x/y vector with ONE table
 * pass and rotate them in the same kernel (vector may be NULL: no rotation).  Outputs are outX*outY*inZ each. */
int fb200_interp_interpolate_vector(const fb200_interp* handle, const fb200_vector* vector, const float* uIn, const float* vIn,
                                    size_t size, float* uOut, float* vOut, size_t* newSize);
int fb200_interp_interpolate_vector_device(const fb200_interp* handle, const fb200_vector* vector, const float* d_uIn, const float* d_vIn,
                                           size_t size, float* d_uOut, float* d_vOut, size_t* newSize, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------
 * (1c) the whole per-slice body of CDMInterpolator::getDataSlice in ONE call (src/CDMInterpolator.cc:250-285)
 *
 *   data2InterpolationArray  (:115-119)  Data::asFloat() + mifi_bad2nanf(badValue)      any CDM numeric type in
 *   interpolateValues        (:258)
 *   [reprojectValues         (:259-283)  with the counterpart component]
 *   interpolationArray2Data  (:121-124)  NaN -> badValue, round + cast to the variable's type
 *
 * inType / outType: enum fb200_datatype of the input slice and of the variable; badValue: CDM::getFillValue(varName)
 * (src/CDM.cc:505-522; NaN switches the input pass off, src/interpolation.c:1776).  On the staged gathers (bilinear,
 * nearest neighbour / coord_nn, bicubic) both adapter passes run INSIDE the gather kernel: the fill -> NaN compare is
 * applied once per staged source value and the output is written once, in the variable's type (a `short` variable
 * halves the bytes stored).  Other methods / 64-bit integer outputs take a float slab and one conversion pass.
 * outData: outX*outY*inZ elements of outType.  Pre/post-processes added with fb200_interp_add_pre/postprocess run inside the call.
 * ---------------------------------------------------------------------------------------------- */
/* CDMInterpolator::addPreprocess / addPostprocess (src/CDMInterpolator.cc:1886-1896) with the option strings of
 * --interpolate.preprocess / --interpolate.postprocess as parseProcess reads them (src/binSrc/fimex.cc:644-671):
 * "fill2d(critx,cor,maxLoop)", "creepfill2d(repeat,weight)" or "creepfill2d(repeat,weight,defaultValue)" (weight is ONE
 * CHARACTER and its code is the weight, as in the reference: "2" means 50).  They run on the device inside the slice calls
 * below (before / after the gather, :256, :284) and switch the fused adapters to the three-pass form.  Part of construction:
 * call before the handle is shared between threads. */
int fb200_interp_add_preprocess(fb200_interp* handle, const char* procString);
int fb200_interp_add_postprocess(fb200_interp* handle, const char* procString);

int fb200_interp_get_data_slice(const fb200_interp* handle, int inType, const void* inData, size_t size, double badValue, int outType,
                                void* outData, size_t* newSize);
int fb200_interp_get_data_slice_device(const fb200_interp* handle, int inType, const void* d_inData, size_t size, double badValue,
                                       int outType, void* d_outData, size_t* newSize, void* cuda_stream);
/* the vector branch (:259-283): both components, each with its own fill value, rotated (vector may be NULL), converted */
int fb200_interp_get_vector_slice(const fb200_interp* handle, const fb200_vector* vector, int inType, const void* uIn, const void* vIn,
                                  size_t size, double badValueU, double badValueV, int outType, void* uOut, void* vOut, size_t* newSize);
int fb200_interp_get_vector_slice_device(const fb200_interp* handle, const fb200_vector* vector, int inType, const void* d_uIn,
                                         const void* d_vIn, size_t size, double badValueU, double badValueV, int outType, void* d_uOut,
                                         void* d_vOut, size_t* newSize, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------
 * (2) the reference's C symbols for this path -- include/fimex/interpolation.h (line of each prototype given)
 * ---------------------------------------------------------------------------------------------- */
int mifi_string_to_interpolation_method(const char* stringMethod); /* :44 */

int mifi_interpolate_f(int method, const char* proj_input, const float* infield, const double* in_x_axis, const double* in_y_axis,
                       const int in_x_axis_type, const int in_y_axis_type, const int ix, const int iy, const int iz,
                       const char* proj_output, float* outfield, const double* out_x_axis, const double* out_y_axis,
                       const int out_x_axis_type, const int out_y_axis_type, const int ox, const int oy); /* :71-75 */

int mifi_vector_reproject_values_f(int method, const char* proj_input, const char* proj_output, float* u_out, float* v_out,
                                   const double* out_x_axis, const double* out_y_axis, int out_x_axis_type, int out_y_axis_type, int ox,
                                   int oy, int oz); /* :124-130 */
int mifi_vector_reproject_values_by_matrix_f(int method, const double* matrix, float* u_out, float* v_out, int ox, int oy,
                                             int oz); /* :142-145 */
int mifi_vector_reproject_direction_by_matrix_f(int method, const double* matrix, float* angle_out, int ox, int oy, int oz); /* :158-161 */
int mifi_get_vector_reproject_matrix(const char* proj_input, const char* proj_output, const double* out_x_axis, const double* out_y_axis,
                                     int out_x_axis_type, int out_y_axis_type, int ox, int oy, double* matrix); /* :177-182 */
int mifi_get_vector_reproject_matrix_field(const char* proj_input, const char* proj_output, const double* in_x_field,
                                           const double* in_y_field, int ox, int oy, double* matrix); /* :197-201 */
int mifi_get_vector_reproject_matrix_points(const char* proj_input, const char* proj_output, int inputIsMetric, const double* out_x_points,
                                            const double* out_y_points, int on, double* matrix); /* :216-222 */

/* per-point gathers (:231, :274, :291).  Exported for link compatibility; each call moves the whole field to
 * the device, so use the handle API for anything but spot checks. */
int mifi_get_values_f(const float* infield, float* outfield, const double x, const double y, const int ix, const int iy, const int iz);
int mifi_get_values_bilinear_f(const float* infield, float* outvalues, const double x, const double y, const int ix, const int iy,
                               const int iz);
int mifi_get_values_bicubic_f(const float* infield, float* outvalues, const double x, const double y, const int ix, const int iy,
                              const int iz);

int mifi_points2position(double* points, const int n, const double* axis, const int num, const int axis_type); /* :415 */
int mifi_project_values(const char* proj_input, const char* proj_output, double* in_out_x_vals, double* in_out_y_vals,
                        const int num); /* :443 */
int mifi_project_axes(const char* proj_input, const char* proj_output, const double* in_x_axis, const double* in_y_axis, const int ix,
                      const int iy, double* out_xproj_axis, double* out_yproj_axis); /* :461 */

/* bad <-> NaN adapters on either side of the path (src/interpolation.c:1775-1793); :586-587 */
size_t mifi_bad2nanf(float* posPtr, float* endPtr, float badVal);
size_t mifi_nanf2bad(float* posPtr, float* endPtr, float badVal);

/* The 2-D pre/post-processes of getDataSlice (--interpolate.preprocess / postprocess, src/CDMInterpolator.cc:126-159, 256, 284;
 * include/fimex/interpolation.h:520-560): same prototypes as the reference, one level of host data per call.  nx, ny >= 2. */
int mifi_fill2d_f(size_t nx, size_t ny, float* field, float relaxCrit, float corrEff, size_t maxLoop, size_t* nChanged);
int mifi_creepfill2d_f(size_t nx, size_t ny, float* field, unsigned short repeat, char setWeight, size_t* nChanged);
int mifi_creepfillval2d_f(size_t nx, size_t ny, float* field, float defaultVal, unsigned short repeat, char setWeight, size_t* nChanged);
/* the same for nz levels resident on the device, in place (processArray_, src/CDMInterpolator.cc:136-159): one CTA per level
 * walks the anti-diagonals of the lexicographic Gauss-Seidel sweeps, so results are bit-identical to the reference.
 * useDefaultVal != 0: creepfillval2d(defaultVal), else creepfill2d (mean of the defined values as first guess). */
int fb200_fill2d_device(float* d_field, size_t nx, size_t ny, size_t nz, float relaxCrit, float corrEff, size_t maxLoop, void* cuda_stream);
int fb200_creepfill2d_device(float* d_field, size_t nx, size_t ny, size_t nz, int useDefaultVal, float defaultVal, unsigned short repeat,
                             char setWeight, void* cuda_stream);

/* ThreadPool.c's knob (src/ThreadPool.c:33-46): kept as a symbol, has no effect on this path */
int mifi_setNumThreads(int n);

#ifdef __cplusplus
}
#endif

#endif /* FIMEX_B200_H_ */
