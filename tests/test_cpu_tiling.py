"""CPU: the trajectory tiling of the tensor-core GNN rollout (csrc/kernels.cuh: tc_tile / tc_group -- rounds, the 3/4 "pure issuer"
split, mixed rounds, warp-wise filling) evaluated on the host through the probe library: over a sweep of population sizes every
trajectory of every ensemble member is taken by exactly one (CTA pair, CTA, slot, round) group, no group exceeds the tile, and the
rounds that promise an empty MMA-issuing warp keep it empty.  (The kernel evaluates the same two functions; the GPU parity tests cover
specific sizes, this covers the arithmetic for all of them.)"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "cee-us_b200", "libceeus_b200_probe.so")


@pytest.fixture(scope="module")
def probe():
    if not os.path.exists(PROBE):
        import importlib.util

        spec = importlib.util.spec_from_file_location("ceeus_b200_build", os.path.join(ROOT, "cee-us_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = C.CDLL(PROBE)
    lib.ceeus_probe_tc_partition.restype = C.c_int
    lib.ceeus_probe_tc_partition.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    return lib


@pytest.mark.parametrize("n_obj,hidden", [(2, 128), (3, 128), (6, 128), (12, 128), (4, 64), (7, 64)])
@pytest.mark.parametrize("grow", [0, 1])
def test_groups_partition_the_population(probe, n_obj, hidden, grow):
    E, sms = 5, 148
    tpw = 32 // (n_obj + grow)
    sizes = sorted(set(list(range(1, 70)) + list(range(100, 12000, 37)) + [128, 2048, 4096, 8192, 16384]))
    for n in sizes:
        cover = np.zeros(E * n, np.int32)
        issuer = np.zeros(2, np.int32)
        gmax = probe.ceeus_probe_tc_partition(sms, E, n, n_obj, hidden, grow, cover.ctypes.data_as(C.POINTER(C.c_int32)),
                                              issuer.ctypes.data_as(C.POINTER(C.c_int32)))
        assert 0 < gmax <= 4 * tpw, (n, gmax)
        assert (cover == 1).all(), (n, int(cover.min()), int(cover.max()))
        assert issuer[0] == 0, (n, issuer)  # rounds that promise a pure issuer keep its warp empty


def test_mixed_rounds_at_the_benchmark_sizes(probe):
    """config 2 on the global-row layout and config 3 on the node-major layout: the 14-pair member gets one pure-issuer round."""
    for n, n_obj, grow in ((4096, 2, 1), (2048, 6, 0)):
        cover = np.zeros(5 * n, np.int32)
        issuer = np.zeros(2, np.int32)
        probe.ceeus_probe_tc_partition(148, 5, n, n_obj, 128, grow, cover.ctypes.data_as(C.POINTER(C.c_int32)),
                                       issuer.ctypes.data_as(C.POINTER(C.c_int32)))
        assert (cover == 1).all() and issuer[0] == 0 and issuer[1] > 0  # only the even round puts trajectories on issuing warps
