"""Two GPUs in ONE process (fb200_set_device, include/fimex_b200.h): handles on different devices used alternately from one
host thread, through the host-buffer path (per-thread, per-device streams; page-locked bounce buffers) and the device path.
Skipped on a one-GPU box.  Regression test for the advisor's round-1 finding: the host pipeline's streams were created on the
device the thread had used LAST, so the second handle launched on the wrong GPU."""
import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu

import fimex_b200 as fb  # noqa: E402
from fimex_b200 import Method  # noqa: E402


def _two_devices():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.skipif(not _two_devices(), reason="needs two CUDA devices")
@pytest.mark.parametrize("method", [Method.BILINEAR, Method.NEAREST_NEIGHBOR, Method.BICUBIC])
def test_alternating_handles_on_two_devices_in_one_thread(oracle, method):
    import torch
    rng = np.random.default_rng(2)
    inX, inY, outX, outY, nz = 90, 70, 256, 96, 20
    px = rng.uniform(-1, inX, outX * outY)
    py = rng.uniform(-1, inY, outX * outY)
    fields = [rng.normal(250, 30, (nz, inY, inX)).astype(np.float32) for _ in range(2)]
    want = [oracle.cached_interpolate(int(method), px, py, inX, inY, outX, outY, f) for f in fields]
    big = np.tile(fields[0], (6, 1, 1))  # > 8 MB of input + output: the pageable path with bounce buffers
    want_big = np.tile(want[0], (6, 1, 1))
    handles = []
    try:
        for dev in (0, 1):
            fb.set_device(dev)
            handles.append(fb.CachedInterpolation("x", "y", method, px, py, inX, inY, outX, outY))
        before = torch.cuda.current_device()
        for rounds in range(3):
            for dev in (1, 0, 1, 0):  # host arrays, alternating devices, the thread's own current device left alone
                got = handles[dev].interpolateValues(fields[dev])
                assert_bit_equal(got, want[dev], f"host path, device {dev}", nan_payload=(method == Method.NEAREST_NEIGHBOR))
            got = handles[1].interpolateValues(big)
            assert_bit_equal(got, want_big, "pageable host path on device 1", nan_payload=(method == Method.NEAREST_NEIGHBOR))
        assert torch.cuda.current_device() == before
        for dev in (0, 1):  # device-resident slices on each handle's own GPU
            d = torch.from_numpy(fields[dev]).to(f"cuda:{dev}")
            with torch.cuda.device(dev):
                got = handles[dev].interpolateValues(d)
                torch.cuda.synchronize()
            assert got.device.index == dev
            assert_bit_equal(got.cpu().numpy(), want[dev], f"device path, device {dev}", nan_payload=(method == Method.NEAREST_NEIGHBOR))
    finally:
        for h in handles:
            h.close()
        fb.set_device(0)
