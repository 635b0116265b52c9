"""CPU check of the Kholodenko integral used by the kernels (csrc/models.cuh: kho_p0sq_series):
the device functions are plain C++ apart from their CUDA qualifiers, so they are compiled here
for the host with g++ and compared with (a) QUADPACK with the reference's settings
(kholodenko.py:31-38: epsabs 0, epsrel 1e-10) and (b) the composite-panel integrator kept in the
same header as a cross-check."""
import ctypes
import math
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM = r'''
#include <cmath>
#include <algorithm>
#define __device__
#define __noinline__
#define __forceinline__ inline
#define __constant__
#include "snippet.h"
extern "C" double p0sq(double q, double lk, double x, int series)
{
    double t = q * q * lk * lk / 9.;
    int regime = (q < 3.0 / lk) ? -1 : ((q > 3.0 / lk) ? 1 : 0);
    double ef = regime < 0 ? std::sqrt(1.0 - t) : (regime > 0 ? std::sqrt(t - 1.0) : 0.);
    return series ? kho_p0sq_series(regime, ef, x) : kho_p0sq_panels(regime, ef, x);
}
'''


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = open(os.path.join(ROOT, "mcsas_b200", "csrc", "models.cuh")).read()
    a = src.index("static __constant__ double GL16_X[8]")
    b = src.index("template <class Real> struct Ops<MCSAS_KHOLODENKO, Real> {")
    d = tmp_path_factory.mktemp("kho")
    (d / "snippet.h").write_text(src[a:b])
    (d / "host.cpp").write_text(SHIM)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", str(d / "libkho.so"), str(d / "host.cpp")])
    L = ctypes.CDLL(str(d / "libkho.so"))
    L.p0sq.restype = ctypes.c_double
    L.p0sq.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int]
    return L


def _core(z, q, lk, x):                                   # kholodenko.py:16-29
    t = (q * lk / 3.) ** 2
    if z <= 0:
        return 2. / x
    if t < 1:
        e = math.sqrt(1 - t)
        fz = math.sinh(e * z) / (e * math.sinh(z)) if z < 300 else math.exp((e - 1) * z) / e
    elif t > 1:
        f = math.sqrt(t - 1)
        fz = math.sin(f * z) / (f * math.sinh(z)) if z < 300 else 0.
    else:
        fz = z / math.sinh(z)
    return fz * (2. / x) * (1. - z / x)


def test_series_matches_quadpack_on_the_model_ranges(host_lib):
    from scipy.integrate import quad
    rng = np.random.default_rng(1)
    cases = [(0.3, 10., 1000.), (0.3000001, 10., 1000.), (0.2999999, 10., 1000.), (1e-4, 10., 100.),
             (10., 50., 1000.), (0.01, 50., 10.)]
    cases += [(10 ** rng.uniform(-2, 1), rng.uniform(10, 50), rng.uniform(100, 1000)) for _ in range(120)]
    worst = 0.
    for q, lk, L in cases:                                # q [1/nm], lK, L [nm]
        x = 3. * L / lk
        ref, _ = quad(_core, 0, x, args=(q, lk, x), limit=10000, epsabs=0, epsrel=1e-10)
        worst = max(worst, abs(host_lib.p0sq(q, lk, x, 1) / ref - 1.))
    assert worst < 1e-9, worst


def test_series_matches_panel_integration_far_outside_the_defaults(host_lib):
    rng = np.random.default_rng(7)
    worst = 0.
    for _ in range(1500):
        q, lk, L = 10 ** rng.uniform(-2.5, 1.3), 10 ** rng.uniform(0, 2), 10 ** rng.uniform(1, 3.5)
        x = 3. * L / lk
        worst = max(worst, abs(host_lib.p0sq(q, lk, x, 1) / host_lib.p0sq(q, lk, x, 0) - 1.))
    assert worst < 1e-11, worst
