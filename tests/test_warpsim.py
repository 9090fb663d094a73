"""The warp-aggregated gradient sinks (device code: segmented warp scans, run detection, last-lane reductions) executed on
the CPU by tests/hostsim/warpsim.cu - 32 host threads play the lanes of a warp, the intrinsics are emulated - and compared
with the plain per-lane scatter of the same samples.  Test infrastructure only; covers l4d_warp_runs, l4d_seg_sum8,
l4d_plane_scatter_warp{,_t}, l4d_row_scatter_warp and l4d_static_scatter_warp of lidar4d_b200/csrc/l4d_bwd.cuh."""
import ctypes as C

import numpy as np
import pytest

import hostsim_util


@pytest.fixture(scope="module")
def ws():
    SO = hostsim_util.build_warpsim()
    lib = C.CDLL(SO)
    fp, ip, up = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_uint)
    lib.ws_plane_sinks.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp]
    lib.ws_static_level.argtypes = [C.c_float, C.c_uint32, C.c_uint32, C.c_int, fp, fp, fp, fp]
    lib.ws_warp_runs.argtypes = [ip, ip, ip, ip, up]
    return lib


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def _close(a, b):
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / scale


@pytest.mark.parametrize("keys", [
    list(range(32)),                                            # every lane its own run
    [7] * 32,                                                   # one run
    [0] * 5 + [1] * 1 + [2] * 10 + [3] * 16,
    [3, 3, 4, 4, 4, 9, 9, 9, 9, 1] + [5] * 21 + [6],
    [i // 3 for i in range(32)],
    [0, 1] * 16,                                                # equal keys that are NOT neighbours form different runs
])
def test_warp_runs(ws, keys):
    k = np.asarray(keys, np.int32)
    dist, maxd, tail = np.zeros(32, np.int32), np.zeros(32, np.int32), np.zeros(32, np.int32)
    mask = np.zeros(32, np.uint32)
    assert ws.ws_warp_runs(_p(k, C.c_int), _p(dist, C.c_int), _p(maxd, C.c_int), _p(tail, C.c_int), _p(mask, C.c_uint)) == 0
    first = np.zeros(32, np.int64)
    for i in range(32):
        first[i] = i if i == 0 or k[i] != k[i - 1] else first[i - 1]
    last = np.zeros(32, np.int64)
    for i in range(31, -1, -1):
        last[i] = i if i == 31 or k[i] != k[i + 1] else last[i + 1]
    assert np.array_equal(dist, np.arange(32) - first)
    assert np.all(maxd == (np.arange(32) - first).max())
    assert np.array_equal(tail, (last == np.arange(32)).astype(np.int32))
    want = np.array([sum(1 << j for j in range(first[i], last[i] + 1)) for i in range(32)], np.uint32)
    assert np.array_equal(mask, want)


def _ray(n, rng, step, D=3):
    """n consecutive samples of rays through [0,1]^D (a new ray every 96 samples), `step` apart: runs of equal cells"""
    x = np.zeros((n, D), np.float32)
    for s in range(0, n, 96):
        o = rng.uniform(0.0, 1.0, D)
        d = rng.normal(size=D)
        d /= np.linalg.norm(d)
        t = np.arange(min(96, n - s))[:, None] * step
        x[s:s + 96] = np.clip(o + d * t, 0.0, 1.0)          # clipping exercises the border rule of grid_sample
    return x


@pytest.mark.parametrize("mode,H,W,step", [(0, 32, 32, 0.002), (0, 256, 256, 0.0011), (1, 8, 32, 0.002), (1, 8, 256, 0.0011),
                                            (2, 8, 64, 0.0015), (2, 8, 256, 0.0011), (2, 8, 32, 0.05)])
def test_plane_sinks_match_the_per_lane_scatter(ws, mode, H, W, step):
    rng = np.random.default_rng(100 * mode + W)
    n = 32 * 24
    x = _ray(n, rng, step, 2)
    cx = np.ascontiguousarray(x[:, 0])
    cy = np.ascontiguousarray(x[:, 1]) if mode == 0 else np.full(n, 0.37, np.float32)      # time planes: one tau per launch
    g = rng.normal(size=(n, 8)).astype(np.float32)
    g[rng.uniform(size=n) < 0.1] = 0.0                                                          # lanes without a contribution
    rows = 1 if mode == 2 else H
    Ga, Gr = np.zeros((rows, W, 8), np.float32), np.zeros((rows, W, 8), np.float32)
    assert ws.ws_plane_sinks(mode, H, W, n, _p(cx), _p(cy), _p(g), _p(Ga), _p(Gr)) == 0
    assert np.abs(Gr).max() > 0
    assert _close(Ga, Gr) < 2e-6, (mode, W)
    assert np.array_equal(Ga != 0, Gr != 0) or _close(Ga, Gr) < 2e-6


@pytest.mark.parametrize("res,entries,step", [(512, 1 << 19, 0.00055), (1176, 1 << 19, 0.00055), (16, 4096, 0.004), (64, 1 << 12, 0.003)])
def test_static_level_matches_the_per_lane_scatter(ws, res, entries, step):
    rng = np.random.default_rng(res)
    n = 32 * 24
    x = np.ascontiguousarray(_ray(n, rng, step, 3))
    dd = rng.normal(size=(n, 4)).astype(np.float32)
    dd[-7:] = 0.0                                                      # the shadow lanes past the end of a launch
    Ga, Gr = np.zeros((entries, 4), np.float32), np.zeros((entries, 4), np.float32)
    assert ws.ws_static_level(float(res - 1), res, entries, n, _p(x), _p(dd), _p(Ga), _p(Gr)) == 0
    assert np.abs(Gr).max() > 0
    assert _close(Ga, Gr) < 2e-6
    assert np.array_equal(np.abs(Ga).sum(1) > 0, np.abs(Gr).sum(1) > 0)      # the same entries are touched
