// tests/cpp/testInterpolation_b200.cc -- the reference's own unit tests of this path, restated against libfimex_b200.so.
//
// Source of every case: /root/reference/test/testInterpolation.cc (Boost.Test there; a tiny CHECK macro here because Boost
// is not in the image).  The calls are the SAME C symbols (mifi_*) the reference's tests call, resolved from our
// library, plus the C++ mirror classes of include/fimex_b200/Cached.h.  Needs a GPU; driven by tests/test_cpp_harness.py.
#include "fimex_b200/Cached.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static int failures = 0;
static int checks = 0;
#define CHECK(cond)                                                                      \
    do {                                                                                 \
        ++checks;                                                                        \
        if (!(cond)) {                                                                   \
            ++failures;                                                                  \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
        }                                                                                \
    } while (0)
#define CHECK_CLOSE(a, b, pct) CHECK(std::fabs((a) - (b)) <= (pct) / 100. * std::fmax(std::fabs(a), std::fabs(b)))

static bool near(float a, float b, float eps) { return std::fabs(a - b) < eps; }
static const double RAD_TO_DEG = 57.29577951308232;

// test/testInterpolation.cc:49-58
static void test_mifi_points2position()
{
    double axis[5] = {1., 2., 3., 4., 5.};
    double points[5] = {-3., 5., 1.3, 2., 6.};
    double apoints[5] = {-4., 4., 0.3, 1., 5.};
    CHECK(mifi_points2position(points, 5, axis, 5, MIFI_PROJ_AXIS) == MIFI_OK);
    for (int i = 0; i < 5; ++i)
        CHECK(near(apoints[i], points[i], 1e-10));
}

// :61-70
static void test_mifi_points2position_reverse()
{
    double axis[5] = {5., 4., 3., 2., 1.};
    double points[5] = {-3., 5., 1.3, 2., 6.};
    double apoints[5] = {8., 0., 3.7, 3., -1.};
    mifi_points2position(points, 5, axis, 5, MIFI_PROJ_AXIS);
    for (int i = 0; i < 5; ++i)
        CHECK(near(apoints[i], points[i], 1e-10));
}

// :73-80
static void test_mifi_get_values_f()
{
    float infield[4] = {1., 2., 1., 2.};
    float outvalues[1];
    mifi_get_values_f(infield, outvalues, 0.3, 0.3, 2, 2, 1);
    CHECK(near(outvalues[0], 1, 1e-10));
}

// :83-112
static void test_mifi_get_values_bilinear_f()
{
    float infield[4] = {1., 2., 2., 1 + std::sqrt(2.0f)};
    float outvalues[1];
    mifi_get_values_bilinear_f(infield, outvalues, 0.3, 0., 2, 2, 1);
    CHECK(near(outvalues[0], 1.3, 1e-6));
    mifi_get_values_bilinear_f(infield, outvalues, 0.3, 0.0001, 2, 2, 1);
    CHECK(near(outvalues[0], 1.3, 1e-4));
    mifi_get_values_bilinear_f(infield, outvalues, 0., 0.3, 2, 2, 1);
    CHECK(near(outvalues[0], 1.3, 1e-6));
    mifi_get_values_bilinear_f(infield, outvalues, 0.0001, 0.3, 2, 2, 1);
    CHECK(near(outvalues[0], 1.3, 1e-4));
    mifi_get_values_bilinear_f(infield, outvalues, 0, 0, 2, 2, 1);
    CHECK(!std::isnan(outvalues[0]));
    mifi_get_values_bilinear_f(infield, outvalues, 1, 1, 2, 2, 1);
    CHECK(!std::isnan(outvalues[0]));
    mifi_get_values_bilinear_f(infield, outvalues, 1.5, 0.5, 2, 2, 1);
    CHECK(std::isnan(outvalues[0]));
    mifi_get_values_bilinear_f(infield, outvalues, 0.5, 1.5, 2, 2, 1);
    CHECK(std::isnan(outvalues[0]));
    mifi_get_values_bilinear_f(infield, outvalues, 0.5, -0.5, 2, 2, 1);
    CHECK(std::isnan(outvalues[0]));
    mifi_get_values_bilinear_f(infield, outvalues, -0.5, 0.5, 2, 2, 1);
    CHECK(std::isnan(outvalues[0]));
}

// :115-155
static void test_mifi_get_values_bicubic_f()
{
    float infield[16] = {1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1};
    float outvalues[1];
    float infield_t[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            infield_t[i + 4 * j] = infield[j + 4 * i];
    mifi_get_values_bicubic_f(infield, outvalues, 1, 1, 4, 4, 1);
    CHECK_CLOSE(2.f, outvalues[0], 1e-3);
    mifi_get_values_bicubic_f(infield, outvalues, 1, 1.99999, 4, 4, 1);
    CHECK_CLOSE(2.f, outvalues[0], 1e-3);
    mifi_get_values_bicubic_f(infield, outvalues, 1, 1.5, 4, 4, 1);
    CHECK_CLOSE(2.125f, outvalues[0], 1e-3);
    mifi_get_values_bicubic_f(infield, outvalues, 1.5, 1, 4, 4, 1);
    CHECK_CLOSE(2.f, outvalues[0], 1e-3);
    mifi_get_values_bicubic_f(infield_t, outvalues, 1, 1, 4, 4, 1);
    CHECK_CLOSE(2.f, outvalues[0], 1e-3);
    mifi_get_values_bicubic_f(infield_t, outvalues, 1.99999, 1, 4, 4, 1);
    CHECK_CLOSE(2.f, outvalues[0], 1e-3);
    mifi_get_values_bicubic_f(infield_t, outvalues, 1.5, 1, 4, 4, 1);
    CHECK_CLOSE(2.125f, outvalues[0], 1e-3);
    mifi_get_values_bicubic_f(infield_t, outvalues, 1, 1.5, 4, 4, 1);
    CHECK_CLOSE(2.f, outvalues[0], 1e-3);
    mifi_get_values_bicubic_f(infield, outvalues, .5, 1, 4, 4, 1);
    CHECK(std::isnan(outvalues[0]));
    mifi_get_values_bicubic_f(infield, outvalues, 1, .5, 4, 4, 1);
    CHECK(std::isnan(outvalues[0]));
    mifi_get_values_bicubic_f(infield, outvalues, 2.5, 1, 4, 4, 1);
    CHECK(std::isnan(outvalues[0]));
    mifi_get_values_bicubic_f(infield, outvalues, 1, 2.5, 4, 4, 1);
    CHECK(std::isnan(outvalues[0]));
}

// :265-278
static void test_mifi_project_axes()
{
    std::string emepProj("+ellps=sphere +a=127.4 +e=0 +proj=stere +lat_0=90 +lon_0=-32 +lat_ts=60 +x_0=7 +y_0=109");
    std::string latlongProj("+ellps=sphere +a=6370 +e=0 +proj=latlong");
    double emepX[] = {6, 7, 8};
    double emepY[] = {108, 109, 110};
    double outX[9];
    double outY[9];
    CHECK(MIFI_OK == mifi_project_axes(emepProj.c_str(), latlongProj.c_str(), &emepX[0], &emepY[0], 3, 3, &outX[0], &outY[0]));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            CHECK((RAD_TO_DEG * outY[j + 3 * i]) > 89);
}

// :396-512 (rotate 90 / 180 degrees between two polar-stereographic grids)
static void test_mifi_vector_reproject_values_rotate(int lon0, double tol)
{
    std::string emepProj("+ellps=sphere +a=127.4 +e=0 +proj=stere +lat_0=90 +lon_0=0 +lat_ts=60");
    std::string emepProj2 = "+ellps=sphere +a=127.4 +e=0 +proj=stere +lat_0=90 +lon_0=" + std::to_string(lon0) + " +lat_ts=60";
    double ax[5];
    float u[25], v[25], uOut[25], vOut[25], uRot[25], vRot[25];
    for (int i = 0; i < 5; ++i)
        ax[i] = i - 2;
    for (int i = 0; i < 25; ++i) {
        u[i] = i;
        v[i] = 25 - i;
    }
    mifi_interpolate_f(0, emepProj.c_str(), u, ax, ax, MIFI_PROJ_AXIS, MIFI_PROJ_AXIS, 5, 5, 1, emepProj2.c_str(), uOut, ax, ax, MIFI_PROJ_AXIS,
                       MIFI_PROJ_AXIS, 5, 5);
    mifi_interpolate_f(0, emepProj.c_str(), v, ax, ax, MIFI_PROJ_AXIS, MIFI_PROJ_AXIS, 5, 5, 1, emepProj2.c_str(), vOut, ax, ax, MIFI_PROJ_AXIS,
                       MIFI_PROJ_AXIS, 5, 5);
    for (int i = 0; i < 25; ++i) {
        uRot[i] = uOut[i];
        vRot[i] = vOut[i];
    }
    CHECK(mifi_vector_reproject_values_f(MIFI_VECTOR_KEEP_SIZE, emepProj.c_str(), emepProj2.c_str(), uOut, vOut, ax, ax, MIFI_PROJ_AXIS,
                                         MIFI_PROJ_AXIS, 5, 5, 1) == MIFI_OK);
    for (int i = 0; i < 25; ++i) {
        if (lon0 == 90) {
            CHECK(std::fabs(vRot[i] - uOut[i]) < tol);
            CHECK(std::fabs(uRot[i] + vOut[i]) < tol);
        } else {
            CHECK(std::fabs(vRot[i] + vOut[i]) < tol);
            CHECK(std::fabs(uRot[i] + uOut[i]) < tol);
        }
    }
}

// the C++ mirror classes: CachedInterpolation / createReducedDomain / CachedVectorReprojection / CachedForwardInterpolation
static void test_cached_classes()
{
    using namespace MetNoFimexB200;
    const size_t inX = 6, inY = 5, outX = 3, outY = 2;
    std::vector<double> px = {2.25, 2.5, 2.75, 2.25, 2.5, 2.75};
    std::vector<double> py = {1.5, 1.5, 1.5, 2.0, 2.0, 2.0};
    CachedInterpolation ci("x", "y", FB200_INTERPOL_BILINEAR, px, py, inX, inY, outX, outY);
    CHECK(ci.getInX() == inX && ci.getOutY() == outY);
    shared_float_array in(new float[inX * inY * 2]);
    for (size_t i = 0; i < inX * inY * 2; ++i)
        in[i] = (float)(i % (inX * inY)) + 100.f * (float)(i / (inX * inY)); // value = y*inX + x (+100 per level)
    size_t newSize = 0;
    shared_float_array out = ci.interpolateValues(in, inX * inY * 2, newSize);
    CHECK(newSize == outX * outY * 2);
    for (size_t z = 0; z < 2; ++z)
        for (size_t i = 0; i < outX * outY; ++i)
            CHECK(near(out[z * outX * outY + i], (float)(py[i] * inX + px[i]) + 100.f * z, 1e-4)); // bilinear is exact on a plane
    // createReducedDomain (src/CachedInterpolation.cc:159-200): bbox 2..3 x 1..2, +-2 cells, clamped
    ci.createReducedDomain("x", "y");
    CHECK(ci.reducedDomain().get() != 0);
    CHECK(ci.reducedDomain()->xMin == 0 && ci.reducedDomain()->yMin == 0 && ci.reducedDomain()->xOrg == inX);
    CHECK(ci.getInX() == 6 && ci.getInY() == 5); // ceil(2.75)+2 = 5 -> cols 0..5, rows 0..4
    bool thrown = false;
    try {
        CachedInterpolation bad("x", "y", 99, px, py, inX, inY, outX, outY);
    } catch (CDMException&) {
        thrown = true;
    }
    CHECK(thrown); // "unknown interpolation function", src/CachedInterpolation.cc:114

    // rotation by 90 degrees everywhere: u' = -v, v' = u
    std::shared_ptr<double[]> m(new double[4 * outX * outY]);
    for (size_t i = 0; i < outX * outY; ++i) {
        m[4 * i] = 0.;
        m[4 * i + 1] = 1.;
        m[4 * i + 2] = -1.;
        m[4 * i + 3] = 1.5707963267948966;
    }
    CachedVectorReprojection cvr(MIFI_VECTOR_KEEP_SIZE, m, (int)outX, (int)outY);
    shared_float_array u(new float[outX * outY]), v(new float[outX * outY]);
    for (size_t i = 0; i < outX * outY; ++i) {
        u[i] = (float)i;
        v[i] = 10.f + (float)i;
    }
    cvr.reprojectValues(u, v, outX * outY);
    for (size_t i = 0; i < outX * outY; ++i) {
        CHECK(u[i] == -(10.f + (float)i));
        CHECK(v[i] == (float)i);
    }
    CachedVectorReprojection ident; // not initialised: identity (src/CachedVectorReprojection.cc:37-40)
    ident.reprojectValues(u, v, outX * outY);
    CHECK(v[0] == 0.f);

    // forward mean: two source points into cell (0,0), one into (1,0), cell (2,0) empty
    std::vector<double> fx = {0.2, -0.3, 1.4, 9.0};
    std::vector<double> fy = {0.1, 0.4, 0.0, 0.0};
    CachedForwardInterpolation cfi("x", "y", FB200_INTERPOL_FORWARD_MEAN, fx, fy, 4, 1, 3, 1);
    shared_float_array fin(new float[4]);
    fin[0] = 1.f;
    fin[1] = 2.f;
    fin[2] = 5.f;
    fin[3] = 7.f;
    shared_float_array fout = cfi.interpolateValues(fin, 4, newSize);
    CHECK(newSize == 3);
    CHECK(fout[0] == 1.5f && fout[1] == 5.f && std::isnan(fout[2]));
}

// the per-slice body of CDMInterpolator::getDataSlice (src/CDMInterpolator.cc:250-285) through the mirror: a packed short
// variable with _FillValue -32767, nearest neighbour -> short out; the fill value survives both adapters
static void test_get_data_slice()
{
    using namespace MetNoFimexB200;
    const size_t inX = 4, inY = 3, outX = 3, outY = 2;
    std::vector<double> px = {0.2, 1.6, 3.0, 0.0, 2.4, 9.0};
    std::vector<double> py = {0.0, 0.4, 1.6, 2.0, 1.0, 1.0};
    CachedInterpolation ci("x", "y", FB200_INTERPOL_NEAREST_NEIGHBOR, px, py, inX, inY, outX, outY);
    std::vector<short> in(inX * inY);
    for (size_t i = 0; i < in.size(); ++i)
        in[i] = (short)(10 * i);
    in[6] = -32767; // (x=2, y=1)
    std::vector<short> out(outX * outY, 0);
    size_t newSize = 0;
    ci.getDataSlice(FB200_SHORT, in.data(), in.size(), -32767., FB200_SHORT, out.data(), newSize);
    CHECK(newSize == outX * outY);
    CHECK(out[0] == 0 && out[1] == 20 && out[2] == 110 && out[3] == 80);
    CHECK(out[4] == -32767); // the undefined source value
    CHECK(out[5] == -32767); // outside the source grid
    // bilinear to float: 0.5 * (in[0] + in[1]) and rounding to short (half away from zero): (0 + 10) / 2 = 5
    std::vector<double> bx = {0.5, 0.25}, by = {0.0, 0.0};
    CachedInterpolation cb("x", "y", FB200_INTERPOL_BILINEAR, bx, by, inX, inY, 2, 1);
    std::vector<short> bo(2, 0);
    cb.getDataSlice(FB200_SHORT, in.data(), in.size(), -32767., FB200_SHORT, bo.data(), newSize);
    CHECK(bo[0] == 5 && bo[1] == 3); // 2.5 -> 3
    std::vector<float> bf(2, 0.f);
    cb.getDataSlice(FB200_SHORT, in.data(), in.size(), -32767., FB200_FLOAT, bf.data(), newSize);
    CHECK(bf[0] == 5.f && bf[1] == 2.5f);
}

// CDMProcessor::rotateVectorToLatLon on the mirror (makeCachedVectorReprojection, src/CDMProcessor.cc:99-145; test_rotate,
// test/testProcessor.cc:71-93): on a polar-stereographic grid the components change, the speed does not, fill values survive
static void test_rotate_vector_to_latlon()
{
    using namespace MetNoFimexB200;
    const std::string proj = "+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +units=m +a=6.371e+06 +e=0 +no_defs";
    std::vector<double> x(11), y(11);
    for (int i = 0; i < 11; ++i) {
        x[i] = -1705516. + 50162. * i;
        y[i] = -6872225. + 50162. * i;
    }
    CachedVectorReprojection cvr(MIFI_VECTOR_KEEP_SIZE, proj, x, y, false, true);
    CHECK(cvr.getXSize() == 11 && cvr.getYSize() == 11);
    std::vector<float> u(121), v(121), un(121), vn(121);
    for (int i = 0; i < 121; ++i) {
        u[i] = -3.f + 0.05f * i;
        v[i] = -8.f - 0.02f * i;
    }
    u[7] = 9.96921e+36f;
    cvr.getVectorSlice(FB200_FLOAT, u.data(), v.data(), 121, 9.9692099683868690e+36, 9.9692099683868690e+36, FB200_FLOAT, un.data(),
                       vn.data());
    CHECK(un[3] != u[3] && vn[3] != v[3]);
    for (int i = 0; i < 121; ++i) {
        if (i == 7)
            continue;
        const double a = (double)un[i] * un[i] + (double)vn[i] * vn[i], b = (double)u[i] * u[i] + (double)v[i] * v[i];
        CHECK(std::fabs(a - b) < 1e-5 * b);
    }
    CHECK(un[7] == 9.96921e+36f && vn[7] == 9.96921e+36f);
    // and back again
    CachedVectorReprojection back(MIFI_VECTOR_KEEP_SIZE, proj, x, y, false, false);
    std::vector<float> ub(121), vb(121);
    back.getVectorSlice(FB200_FLOAT, un.data(), vn.data(), 121, 9.9692099683868690e+36, 9.9692099683868690e+36, FB200_FLOAT, ub.data(),
                        vb.data());
    for (int i = 0; i < 121; ++i)
        if (i != 7)
            CHECK(std::fabs(ub[i] - u[i]) < 2e-3 && std::fabs(vb[i] - v[i]) < 2e-3);
}

// both components of an x/y pair are interpolated and rotated ONCE; the counterpart's getDataSlice call
// is served from the pair cache (the reference computes the pair twice, src/CDMInterpolator.cc:259-276)
static void test_vector_pair_cache()
{
    using namespace MetNoFimexB200;
    const size_t inX = 6, inY = 5, outX = 3, outY = 2, n = outX * outY;
    std::vector<double> px = {2.25, 2.5, 2.75, 2.25, 2.5, 2.75};
    std::vector<double> py = {1.5, 1.5, 1.5, 2.0, 2.0, 2.0};
    CachedInterpolation ci("x", "y", FB200_INTERPOL_BILINEAR, px, py, inX, inY, outX, outY);
    std::shared_ptr<double[]> m(new double[4 * n]);
    for (size_t i = 0; i < n; ++i) { // 90 degrees: u' = -v, v' = u
        m[4 * i] = 0.;
        m[4 * i + 1] = 1.;
        m[4 * i + 2] = -1.;
        m[4 * i + 3] = 1.5707963267948966;
    }
    CachedVectorReprojection cvr(MIFI_VECTOR_KEEP_SIZE, m, (int)outX, (int)outY);
    std::vector<float> u(inX * inY), v(inX * inY);
    for (size_t i = 0; i < u.size(); ++i) {
        u[i] = (float)i;
        v[i] = 100.f - (float)i;
    }
    u[8] = 9.96921e+36f; // a tap of none of the targets' cells? (x=2, y=1): tap of the first row of targets -> fill values there
    const double fill = 9.9692099683868690e+36;
    std::vector<float> wantU(n), wantV(n);
    size_t ns = 0;
    ci.getVectorSlice(cvr.handle(), FB200_FLOAT, u.data(), v.data(), u.size(), fill, fill, FB200_FLOAT, wantU.data(), wantV.data(), ns);
    CHECK(ns == n);
    VectorPairCache cache(2);
    const unsigned long long before = fb200_kernel_launches();
    std::vector<float> gotU(n), gotV(n);
    cache.getDataSlice(ci, cvr.handle(), "x_wind|y_wind|0", 0, FB200_FLOAT, u.data(), v.data(), u.size(), fill, fill, FB200_FLOAT, sizeof(float),
                       gotU.data(), ns);
    const unsigned long long mid = fb200_kernel_launches();
    CHECK(mid > before && cache.size() == 1);
    cache.getDataSlice(ci, cvr.handle(), "x_wind|y_wind|0", 1, FB200_FLOAT, v.data(), u.data(), u.size(), fill, fill, FB200_FLOAT, sizeof(float),
                       gotV.data(), ns);
    CHECK(fb200_kernel_launches() == mid); // the second half cost no kernel at all
    CHECK(cache.size() == 0);              // handed out once
    for (size_t i = 0; i < n; ++i) {
        CHECK(std::memcmp(&gotU[i], &wantU[i], 4) == 0);
        CHECK(std::memcmp(&gotV[i], &wantV[i], 4) == 0);
    }
    CHECK(gotU[0] == 9.96921e+36f); // the undefined tap reaches both rotated components as the fill value
    // a different slice of the same pair is a different key; asking for y first works the same way
    cache.getDataSlice(ci, cvr.handle(), "x_wind|y_wind|1", 1, FB200_FLOAT, v.data(), u.data(), u.size(), fill, fill, FB200_FLOAT, sizeof(float),
                       gotV.data(), ns);
    CHECK(cache.size() == 1);
    std::vector<float> again(n);
    CHECK(!cache.take("x_wind|y_wind|0", 0, again.data(), n * sizeof(float)));
    CHECK(cache.take("x_wind|y_wind|1", 0, again.data(), n * sizeof(float)));
    for (size_t i = 0; i < n; ++i)
        CHECK(std::memcmp(&again[i], &wantU[i], 4) == 0);
    cache.clear();
}

int main()
{
    test_mifi_points2position();
    test_mifi_points2position_reverse();
    test_mifi_get_values_f();
    test_mifi_get_values_bilinear_f();
    test_mifi_get_values_bicubic_f();
    test_mifi_project_axes();
    test_mifi_vector_reproject_values_rotate(90, 1e-4);
    test_mifi_vector_reproject_values_rotate(180, 1e-5);
    test_cached_classes();
    test_get_data_slice();
    test_rotate_vector_to_latlon();
    test_vector_pair_cache();
    std::printf("%d checks, %d failures\n", checks, failures);
    return failures == 0 ? 0 : 1;
}
