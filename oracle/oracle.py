# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/liboracle.so, the
plain-C float64 restatement of McSAS's Monte-Carlo fitting core
(oracle/mcsas_oracle.c).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module; the
product package mcsas_b200 never does.

Model parameter vectors follow the reference's ``parameters`` tuples
(src/mcsas/models/*.py), SI units.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")

MAX_PARAMS, MAX_ACTIVE = 16, 4

NM, SLD = 1e-9, 1e20  # units.py: nm -> m, A^-2 -> m^-2

# name -> (id, [(param, default SI, generator kind, default active range or None, valueRange)], default active)
# generator kind: 0 RandomUniform, 1 RandomExponential (numbergenerator.py:28-31,168-175)
MODEL_SPECS = {
    "Sphere": (0, [("radius", 10 * NM, 0, (1 * NM, 1000 * NM), (0., np.inf)),
                   ("sld", 1e-6 * SLD, 0, None, (0., np.inf))], ["radius"]),
    "LMADenseSphere": (1, [("radius", 1 * NM, 0, None, (0., np.inf)),
                           ("volFrac", 0.1, 0, None, (1e-5, 1.0)),
                           ("mf", -1., 0, None, (-1., 1e6)),
                           ("sld", 1e-6 * SLD, 0, None, (0., np.inf))], ["radius"]),
    "SphericalCoreShell": (2, [("radius", 1 * NM, 1, (0.1 * NM, 1e3 * NM), (0., np.inf)),
                               ("t", 1 * NM, 1, (0.1 * NM, 1e3 * NM), (0., np.inf)),
                               ("eta_c", 3.16e-6 * SLD, 0, None, (0., np.inf)),
                               ("eta_s", 2.53e-6 * SLD, 0, None, (0., np.inf)),
                               ("eta_sol", 0., 0, None, (0., np.inf))], ["radius"]),
    "EllipsoidalCoreShell": (3, [("a", 1 * NM, 1, (0.1 * NM, 1e3 * NM), (0., np.inf)),
                                 ("b", 10 * NM, 1, (1 * NM, 1e4 * NM), (0., np.inf)),
                                 ("t", 1 * NM, 1, (0.1 * NM, 1e3 * NM), (0., np.inf)),
                                 ("eta_c", 3.15e-6 * SLD, 0, None, (0., np.inf)),
                                 ("eta_s", 2.53e-6 * SLD, 0, None, (0., np.inf)),
                                 ("eta_sol", 0., 0, None, (0., np.inf)),
                                 ("intDiv", 100, 0, None, (0, 1e4))], ["a"]),
    "EllipsoidsIsotropic": (4, [("a", 1 * NM, 1, (0.1 * NM, 1e3 * NM), (0.1 * NM, 1e10 * NM)),
                                ("useAspect", 1., 0, None, (0., 1.)),
                                ("c", 10 * NM, 1, (1 * NM, 1e4 * NM), (0.1 * NM, 1e10 * NM)),
                                ("aspect", 10., 1, None, (1e-3, 1e3)),
                                ("intDiv", 100, 0, None, (0, 1e4)),
                                ("sld", 1e-6 * SLD, 0, None, (0., 1e-2 * SLD))], ["a"]),
    "CylindersIsotropic": (5, [("radius", 1 * NM, 1, None, (0.1 * NM, np.inf)),
                               ("useAspect", 1., 0, None, (0., 1.)),
                               ("length", 10 * NM, 1, None, (0.1 * NM, 1e10 * NM)),
                               ("aspect", 10., 1, None, (1e-3, 1e3)),
                               ("intDiv", 100, 0, None, (1, 1e4)),
                               ("sld", 1e-6 * SLD, 0, None, (0., np.inf))], ["radius"]),
    "GaussianChain": (6, [("rg", 1 * NM, 1, (1 * NM, 100 * NM), (0., np.inf)),
                          ("bp", 100 * NM, 0, (0.1 * NM, 1e3 * NM), (0., np.inf)),
                          ("etas", 1e-6 * SLD, 0, (0.1 * SLD, 10 * SLD), (0., np.inf)),
                          ("k", 1.0, 0, (0.1, 10.), (0., np.inf))], ["rg"]),
    "Kholodenko": (7, [("radius", 1 * NM, 1, (1 * NM, 5 * NM), (0., np.inf)),
                       ("lenKuhn", 1 * NM, 0, (10 * NM, 50 * NM), (0., np.inf)),
                       ("lenContour", 2 * NM, 0, (100 * NM, 1000 * NM), (0., np.inf))],
                   ["radius", "lenKuhn", "lenContour"]),
}


class _Model(C.Structure):
    _fields_ = [("model_id", C.c_int), ("n_params", C.c_int),
                ("params", C.c_double * MAX_PARAMS), ("n_active", C.c_int),
                ("active_idx", C.c_int * MAX_ACTIVE),
                ("lo", C.c_double * MAX_ACTIVE), ("hi", C.c_double * MAX_ACTIVE),
                ("quad_rule", C.c_int), ("quad_nodes", C.c_int)]


class _Settings(C.Structure):
    _fields_ = [("n_contribs", C.c_int), ("conv_crit", C.c_double),
                ("max_iter", C.c_longlong), ("comp_exp", C.c_double),
                ("find_background", C.c_int), ("positive_background", C.c_int),
                ("start_from_minimum", C.c_int), ("q_limit_hi", C.c_double)]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or (
            os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "mcsas_oracle.c"))):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_philox4x32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                        C.POINTER(C.c_uint32)]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Model(object):
    """Model description: name, full parameter dict overrides, active list with ranges."""

    def __init__(self, name, active=None, ranges=None, values=None, quad_rule=0,
                 quad_nodes=None):
        self.name = name
        mid, plist, dflt_active = MODEL_SPECS[name]
        self.model_id = mid
        self.pnames = [p[0] for p in plist]
        self.values = np.array([float(p[1]) for p in plist])
        for k, v in (values or {}).items():
            self.values[self.pnames.index(k)] = float(v)
        self.active = list(active if active is not None else dflt_active)
        self.gen_kind = []
        self.lo, self.hi = [], []
        for a in self.active:
            _, _, kind, arange, vrange = plist[self.pnames.index(a)]
            if ranges and a in ranges:
                arange = ranges[a]
            if arange is None:
                arange = vrange
            lo = max(min(arange), vrange[0])
            hi = min(max(arange), vrange[1])
            self.lo.append(lo)
            self.hi.append(hi)
            self.gen_kind.append(kind)
        self.quad_rule = quad_rule
        if quad_nodes is None:
            quad_nodes = int(self.values[self.pnames.index("intDiv")]) if "intDiv" in self.pnames else 0
        self.quad_nodes = int(quad_nodes)

    @property
    def n_active(self):
        return len(self.active)

    def cstruct(self):
        m = _Model()
        m.model_id = self.model_id
        m.n_params = len(self.values)
        for i, v in enumerate(self.values):
            m.params[i] = v
        m.n_active = self.n_active
        for i, a in enumerate(self.active):
            m.active_idx[i] = self.pnames.index(a)
            m.lo[i] = self.lo[i]
            m.hi[i] = self.hi[i]
        m.quad_rule = self.quad_rule
        m.quad_nodes = self.quad_nodes
        return m


def settings(n_contribs=300, conv_crit=1.0, max_iter=100000, comp_exp=0.6666666,
             find_background=True, positive_background=False, start_from_minimum=False,
             q_limit_hi=0.0):
    return _Settings(int(n_contribs), float(conv_crit), int(max_iter), float(comp_exp),
                     int(bool(find_background)), int(bool(positive_background)),
                     int(bool(start_from_minimum)), float(q_limit_hi))


def set_smearing(q_offset=None, weights=None, two_d_coll=False):
    """SmearingConfig.prepared (qOffset, weights) for the canSmear models; None switches it off.
    Process-global, like data.config.smearing in the reference."""
    if q_offset is None:
        rc = lib().orc_set_smearing(0, None, None, 0)
    else:
        qo = np.ascontiguousarray(q_offset, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        assert qo.size == w.size
        rc = lib().orc_set_smearing(int(qo.size), _dp(qo), _dp(w), int(bool(two_d_coll)))
    assert rc == 0


def eval_contribs(model, q, pset, comp_exp=0.6666666):
    """ScatteringModel.calc per contribution (scatteringmodel.py:79-105):
    returns it[M, nq], v[M], w[M], s[M]."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    pset = np.ascontiguousarray(np.atleast_2d(pset), dtype=np.float64)
    M = pset.shape[0]
    it = np.zeros((M, q.size))
    v, w, s = np.zeros(M), np.zeros(M), np.zeros(M)
    m = model.cstruct()
    cv, cw, cs = C.c_double(), C.c_double(), C.c_double()
    for i in range(M):
        rc = lib().orc_eval(C.byref(m), int(q.size), _dp(q), _dp(pset[i]),
                            C.c_double(comp_exp), _dp(it[i]), C.byref(cv), C.byref(cw),
                            C.byref(cs))
        assert rc == 0
        v[i], w[i], s[i] = cv.value, cw.value, cs.value
    return it, v, w, s


def formfactor(model, q, pvals):
    q = np.ascontiguousarray(q, dtype=np.float64)
    pv = np.ascontiguousarray(pvals, dtype=np.float64)
    ff = np.zeros(q.size)
    m = model.cstruct()
    rc = lib().orc_formfactor(C.byref(m), int(q.size), _dp(q), _dp(pv), _dp(ff))
    assert rc == 0
    return ff


def bgscale(I, sigma, M, find_bg=True, pos_bg=False, n_active=1):
    I = np.ascontiguousarray(I, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    M = np.ascontiguousarray(M, dtype=np.float64)
    sc = np.zeros(2)
    conval, agofs = C.c_double(), C.c_double()
    lib().orc_bgscale(int(I.size), _dp(I), _dp(sigma), _dp(M), int(find_bg), int(pos_bg),
                      int(n_active), _dp(sc), C.byref(conval), C.byref(agofs))
    return sc, conval.value, agofs.value


def chain(model, st, q, I, sigma, stream, trace=False):
    """McSAS.mcFit (mcsas.py:287-439) driven by a generator-output stream."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    I = np.ascontiguousarray(I, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    stream = np.ascontiguousarray(stream, dtype=np.float64)
    N, P = st.n_contribs, model.n_active
    rset = np.zeros((N, P))
    fit = np.zeros(q.size)
    sc = np.zeros(2)
    conval, conval0 = C.c_double(), C.c_double()
    nit, nmv = C.c_longlong(), C.c_longlong()
    acc = np.zeros(st.max_iter if trace else 1, dtype=np.uint8)
    chi = np.zeros(st.max_iter if trace else 1)
    m = model.cstruct()
    rc = lib().orc_chain(C.byref(m), C.byref(st), int(q.size), _dp(q), _dp(I), _dp(sigma),
                         _dp(stream), C.c_longlong(stream.size), _dp(rset), _dp(fit), _dp(sc),
                         C.byref(conval), C.byref(conval0), C.byref(nit), C.byref(nmv),
                         acc.ctypes.data_as(C.POINTER(C.c_ubyte)) if trace else None,
                         _dp(chi) if trace else None)
    out = dict(rc=rc, rset=rset, fit=fit, scaling=sc[0], background=sc[1],
               conval=conval.value, conval0=conval0.value, num_iter=nit.value,
               num_moves=nmv.value)
    if trace:
        out["accept"] = acc[:nit.value].copy()
        out["chi"] = chi[:nit.value].copy()
    return out


def philox_stream(seed, chain_id, N, model, start_from_minimum, n_steps):
    P = model.n_active
    n0 = 0 if start_from_minimum else N * P
    out = np.zeros(n0 + n_steps * P)
    gk = (C.c_int * MAX_ACTIVE)(*model.gen_kind)
    lib().orc_philox_stream(C.c_uint64(seed), C.c_uint32(chain_id), int(N), int(P), gk,
                            int(bool(start_from_minimum)), C.c_longlong(n_steps), _dp(out))
    return out


def philox4x32(seed, idx, chain_id, domain):
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32(seed, idx, chain_id, domain, o)
    return list(o)


def fractions(model, st, q, I, sigma, rset):
    """McSAS.histogram per-repetition part (mcsas.py:548-604)."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    I = np.ascontiguousarray(I, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    rset = np.ascontiguousarray(rset, dtype=np.float64)
    N = st.n_contribs
    arrs = [np.zeros(N) for _ in range(8)]
    sc, tot = np.zeros(2), np.zeros(4)
    m = model.cstruct()
    lib().orc_fractions(C.byref(m), C.byref(st), int(q.size), _dp(q), _dp(I), _dp(sigma),
                        _dp(rset), _dp(sc), *[_dp(a) for a in arrs], _dp(tot))
    keys = ["vol", "num", "volsqr", "surf", "min_vol", "min_num", "min_volsqr", "min_surf"]
    out = dict(zip(keys, arrs))
    out["sc"] = sc
    out["totals"] = tot
    return out
