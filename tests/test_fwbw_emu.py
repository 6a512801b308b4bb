"""Host emulation of the Forward/Backward kernels' per-thread column code (csrc/nc_fwbw_core.cuh, compiled for the
host) against the oracle: every alpha and beta bit.  This pins the order logic of the shared-prefix chains, the merged
duplicate edges and the seven-instruction p7_FLogsum without a GPU; the -m gpu tests then check the same code on the
device (synchronisation, memory layout)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nanocall_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R73T = "r73.t.006.ont.model"


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "emu", "fwbw_emu.cpp")
    out = os.path.join(ROOT, "tests", "emu", "libfwbw_emu.so")
    hdr = os.path.join(ROOT, "nanocall_b200", "csrc", "nc_fwbw_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-x", "c++",
                        "-I", os.path.dirname(hdr), "-o", out, src], check=True)
    return C.CDLL(out)


def _emissions(port, table, pm, mean, stdv, start):
    """E[i][j] through the oracle's scaled model and nco_emission (Pore_Model.hpp:24-40,126-138)."""
    n = mean.size
    lib = port.lib

    class M(C.Structure):
        _fields_ = [(k, C.c_float * 4096) for k in ("level_mean", "level_stdv", "sd_mean", "sd_stdv", "sd_lambda",
                                                    "log_level_stdv", "log_sd_lambda")] + [("mean", C.c_float), ("stdv", C.c_float)]
    m0, m1 = M(), M()
    t = np.ascontiguousarray(table, np.float32)
    lib.nco_model_prepare(t.ctypes.data_as(C.c_void_p), C.byref(m0))
    p = np.ascontiguousarray(pm, np.float32)
    lib.nco_model_scale(C.byref(m0), p.ctypes.data_as(C.c_void_p), C.byref(m1))
    lib.nco_emission.restype = C.c_float
    E = np.zeros((n, 4096), np.float32)
    y = np.where(stdv == 0, np.float32(0.01), stdv).astype(np.float32)
    x = (mean - np.float32(pm[2]) * start).astype(np.float32)
    libm = C.CDLL("libm.so.6")   # Event::update_logs uses std::log(float) = glibc logf (numpy's SIMD log differs in rare ulps)
    libm.logf.restype = C.c_float
    ly = np.array([libm.logf(C.c_float(v)) for v in y], np.float32)
    for i in range(n):
        for j in range(4096):
            E[i, j] = lib.nco_emission(C.byref(m1), j, C.c_float(x[i]), C.c_float(y[i]), C.c_float(ly[i]))
    return E


@pytest.mark.parametrize("st", [(0.1, 0.3), (0.07, 0.21)])
def test_emulated_kernel_columns_match_oracle(emu, port, models, st):
    table = models[R73T]["table"]
    rng = np.random.default_rng(11)
    pm = synth.random_params(rng, 1)[0]
    rd = synth.make_read(rng, table, 12, tuple(pm))
    exp = port.fwbw(table, pm, st[0], st[1], rd["mean"], rd["stdv"], rd["start"])
    E = _emissions(port, table, pm, rd["mean"], rd["stdv"], rd["start"])
    # the emission inputs above must be the oracle's own: alpha[0] = E[0] - log(4096)
    assert np.array_equal((E[0] - np.log(np.float32(4096))).astype(np.float32).view(np.uint32), exp["alpha"][0].view(np.uint32))
    lut = api.transition_lut(st[0], st[1])
    tbl = port.flogsum_table().copy()
    tbl[15999] = 0.0
    n = E.shape[0]
    alpha = np.zeros((n, 4096), np.float32)
    beta = np.zeros((n, 4096), np.float32)
    hist = np.zeros(8, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emu.emu_forward(p(lut), p(tbl), p(E), n, C.c_float(np.log(np.float32(4096))), p(alpha)) == 0
    assert emu.emu_backward(p(lut), p(tbl), p(E), n, p(beta), p(hist)) == 0
    assert np.array_equal(alpha.view(np.uint32), exp["alpha"].view(np.uint32)), np.argwhere(alpha != exp["alpha"])[:5]
    assert np.array_equal(beta.view(np.uint32), exp["beta"].view(np.uint32)), np.argwhere(beta != exp["beta"])[:5]
    assert hist.sum() == 16 * 8 and hist[0] < 40   # most (warp, state) pairs take a shared-prefix path
