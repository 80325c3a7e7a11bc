"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/fimex_b200.h declares,
refuses to compute without a GPU (no silent CPU path), and the host-side logic (axis strings, method strings)
matches the reference.  No compute calls need a GPU here.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import fimex_b200
from fimex_b200 import capi
from fimex_b200.interpolator import lon_lat_vals_to_matrix, spatial_axis_spec, tokenize_dotted


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.load()
    names = capi.declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fimex_b200.h but not exported by libfimex_b200.so"
    # and the dynamic symbol table agrees (not just ctypes' lazy lookup)
    out = subprocess.run(["nm", "-D", "--defined-only", capi.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in names if n not in exported]
    assert not missing, missing


def test_reference_symbols_of_the_path_are_all_there():
    # the prototypes of reference include/fimex/interpolation.h that belong to this path
    want = ["mifi_interpolate_f", "mifi_vector_reproject_values_f", "mifi_vector_reproject_values_by_matrix_f",
            "mifi_vector_reproject_direction_by_matrix_f", "mifi_get_vector_reproject_matrix", "mifi_get_vector_reproject_matrix_field",
            "mifi_get_vector_reproject_matrix_points", "mifi_get_values_f", "mifi_get_values_bilinear_f", "mifi_get_values_bicubic_f",
            "mifi_points2position", "mifi_project_values", "mifi_project_axes", "mifi_string_to_interpolation_method"]
    lib = capi.load()
    for n in want:
        assert hasattr(lib, n)


def test_library_is_built_for_sm_100a_only():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", capi.lib_path()], capture_output=True, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_method_strings_match_reference():
    # src/interpolation.c:66-101 (pure host logic, no device needed)
    f = fimex_b200.mifi_string_to_interpolation_method
    assert f("nearestneighbor") == 0 and f("bilinear") == 1 and f("bicubic") == 2
    assert f("coord_nearestneighbor") == 3 and f("coord_kdtree") == 4
    assert f("forward_sum") == 5 and f("forward_mean") == 6 and f("forward_median") == 7 and f("forward_max") == 8
    assert f("forward_min") == 9 and f("forward_undef_sum") == 10 and f("forward_undef_max") == 13
    assert f("forward_undef_min") == 9  # the reference's quirk, :97-98
    assert f("Bilinear") == -1 and f("") == -1


def test_no_cpu_fallback_without_gpu():
    if _has_gpu():
        pytest.skip("a GPU is present; the refusal path cannot be exercised")
    os.environ["FIMEX_B200_QUIET"] = "1"
    try:
        rc, _ = fimex_b200.mifi_points2position([1.5], [1.0, 2.0, 3.0], 0)
        assert rc == fimex_b200.MIFI_ERROR
        assert "no usable CUDA device" in fimex_b200.last_error()
        with pytest.raises(fimex_b200.FimexB200Error):
            fimex_b200.CachedInterpolation("x", "y", fimex_b200.Method.BILINEAR, np.zeros(4), np.zeros(4), 3, 3, 2, 2)
        rc, _ = fimex_b200.mifi_interpolate_f(1, "+proj=latlong +R=1", np.zeros(4, np.float32), [0.0, 1], [0.0, 1], 1, 2, 1,
                                              "+proj=latlong +R=1", [0.5], [0.5], 1, 2)
        assert rc == fimex_b200.MIFI_ERROR
    finally:
        os.environ.pop("FIMEX_B200_QUIET", None)


def test_bad_projection_string_is_an_error_not_a_crash():
    os.environ["FIMEX_B200_QUIET"] = "1"
    try:
        rc, _, _ = fimex_b200.mifi_project_values("+proj=doesnotexist +R=1", "+proj=latlong +R=1", [0.0], [0.0])
        assert rc == fimex_b200.MIFI_ERROR
        assert "unknown projection id" in fimex_b200.last_error()
        rc, _, _ = fimex_b200.mifi_project_values("+proj=latlong", "+proj=latlong +no_defs", [0.0], [0.0])
        assert rc == fimex_b200.MIFI_ERROR  # no ellipsoid with +no_defs: "major axis or radius = 0 or not given"
    finally:
        os.environ.pop("FIMEX_B200_QUIET", None)


def test_tokenize_dotted_like_reference():
    # include/fimex/Utils.h:373-407
    assert tokenize_dotted("5,5.5,6,6.5") == [5, 5.5, 6, 6.5]
    v = tokenize_dotted("3.5,4.5,...,17.5")
    assert v[0] == 3.5 and v[-1] == 17.5 and len(v) == 15
    v = tokenize_dotted("10,8,...,0")
    assert v == [10, 8, 6, 4, 2, 0]
    v = tokenize_dotted("0,0.1,...,1")
    assert len(v) == 11 and v[-1] == 1.0
    with pytest.raises(fimex_b200.FimexB200Error):
        tokenize_dotted("1,...,5")
    assert np.array_equal(spatial_axis_spec("61.5,62,62.5"), [61.5, 62, 62.5])
    with pytest.raises(fimex_b200.FimexB200Error):
        spatial_axis_spec("0,1,...,x;relativeStart=0")


def test_lon_lat_vals_to_matrix_layout():
    # CDMInterpolator.cc:1226-1239: pos = ix + iy*lonSize
    lon2d, lat2d = lon_lat_vals_to_matrix([10.0, 20, 30], [1.0, 2])
    assert np.array_equal(lon2d, [10, 20, 30, 10, 20, 30])
    assert np.array_equal(lat2d, [1, 1, 1, 2, 2, 2])


def test_header_is_plain_c():
    """include/fimex_b200.h must compile as C (the boundary is a C ABI, no C++/torch types)"""
    src = '#include "fimex_b200.h"\nint main(void){ return fb200_interp_in_x(0) != 0; }\n'
    inc = os.path.join(os.path.dirname(capi.HEADER))
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c", "-"], input=src, text=True,
                       capture_output=True)
    assert r.returncode == 0, r.stderr


def test_interpolator_host_logic_without_gpu():
    # argument checks and string handling of the CDMInterpolator mirror that run before any device call
    from fimex_b200 import FimexB200Error, Interpolator, Method

    ip = Interpolator("+proj=stere +lat_0=90 +lon_0=0 +lat_ts=60 +units=m +a=6.371e+06 +e=0 +no_defs", [0.0, 1000.0], [0.0, 1000.0], False)
    assert ip._latlong_of_source() == "+proj=latlong +a=6.371e+06 +e=0"  # "+proj=latlong " + getProj4EarthString()
    assert Interpolator("+proj=ob_tran +o_proj=longlat +lon_0=-40 +o_lat_p=22 +R=6.371e+06 +no_defs", [0.0, 1.0], [0.0, 1.0],
                        True)._latlong_of_source() == "+proj=latlong +R=6.371e+06"
    assert Interpolator("+proj=lcc +lat_1=63 +lat_0=63 +lon_0=15", [0.0, 1.0], [0.0, 1.0], False)._latlong_of_source() == \
        "+proj=latlong +a=6371000 +e=0"  # default earth (ProjectionImpl.cc:152)
    with pytest.raises(FimexB200Error):  # CDMInterpolator.cc:465-469
        ip.changeProjectionToLonLatValues(Method.BILINEAR, [1.0, 2.0], [1.0])
    with pytest.raises(FimexB200Error):  # :706-714: coordinate and forward methods cannot take a template
        ip.changeProjectionToTemplate(Method.COORD_NN_KD, np.zeros((2, 2)), np.zeros((2, 2)))
    with pytest.raises(FimexB200Error):
        ip.changeProjectionToTemplate("forward_mean", np.zeros((2, 2)), np.zeros((2, 2)))
    with pytest.raises(FimexB200Error):  # :533-535
        Interpolator("+proj=latlong +R=6371000", [0.0], [0.0, 1.0], True).changeProjectionToCrossSections(Method.BILINEAR, [("a", [(0, 0)])])
    with pytest.raises(FimexB200Error):  # :541-543
        Interpolator("+proj=latlong +R=6371000", [1.0, 1.0], [0.0, 1.0], True).changeProjectionToCrossSections(Method.BILINEAR, [("a", [(0, 0)])])
    with pytest.raises(FimexB200Error):
        ip.getDataSlice(np.zeros((2, 2), np.float32))  # no changeProjection yet


def test_spatial_axis_spec_relative():
    # test/testSpatialAxisSpec.cc:36-66
    vals = spatial_axis_spec("-450000,-400000,...,50000", 30000, 230000)
    assert vals.size == 11 and vals[0] == -450000 and vals[10] == 50000
    rel = spatial_axis_spec("0,50000,...,x,x+50000;relativeStart=0", 30000, 230000)
    assert rel.size == 6 and rel[0] == 0 and rel[5] == 250000
    # a negative start moves on to the next multiple of the step (src/SpatialAxisSpec.cc:128-133)
    rel = spatial_axis_spec("0,1000,...,x;relativeStart=0", -2500.0, 3999.0)
    assert rel[0] == -1000 and rel[-1] == 3000 and np.all(np.diff(rel) == 1000)
    with pytest.raises(fimex_b200.FimexB200Error):
        spatial_axis_spec("0,50000,...,x;relativeStart=0")  # "require start and end for axisSpec"
    with pytest.raises(fimex_b200.FimexB200Error):
        spatial_axis_spec("0;relativeStart=0", 0.0, 1.0)
    with pytest.raises(fimex_b200.FimexB200Error):
        spatial_axis_spec("0,1,...,5;unit=km")
