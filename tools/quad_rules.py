#!/usr/bin/env python
"""Experiment: orientation-average rules of the quadrature models.  Reference rule (100 linspace nodes:
trapezoid for the cylinder, mean for the ellipsoids) vs n-node Gauss-Legendre, both against a 512-node
Gauss-Legendre value, over the active ranges of BASELINE configs[2] and the config-3 q grid."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, FP64, _lib
NM = 1e-9
q = np.logspace(np.log10(0.03), np.log10(3.0), 500) * 1e9
rng = np.random.default_rng(0)
cases = {
    5: ("CylindersIsotropic", [10 * NM, 1, 100 * NM, 10., 100., 1e14], [0, 3],
        np.stack([10 ** rng.uniform(0, np.log10(50), 200) * NM, 10 ** rng.uniform(0, np.log10(20), 200)], 1)),
    4: ("EllipsoidsIsotropic", [10 * NM, 1, 10 * NM, 10., 100., 1e14], [0, 3],
        np.stack([10 ** rng.uniform(0, np.log10(50), 200) * NM, 10 ** rng.uniform(-1, 1, 200)], 1)),
}
for mid, (name, params, act, pset) in cases.items():
    res = {}
    for tag, rule, nodes in (("reference-100", _lib.QUAD_REFERENCE, 100), ("GL-16", 1, 16), ("GL-32", 1, 32), ("GL-64", 1, 64),
                             ("GL-128", 1, 128), ("GL-512", 1, 512)):
        eng = Engine(0)
        eng.upload_datasets([(q, np.ones_like(q), np.ones_like(q))])
        eng.set_model(mid, params, act, [0, 0], [1, 1], [0, 0], rule, nodes)
        it, v, w, s = eng.eval(pset, 0.6666666, precision=FP64)
        eng.close()
        res[tag] = it
    exact = res["GL-512"]
    env = np.maximum.accumulate(np.abs(exact[:, ::-1]), axis=1)[:, ::-1]
    print(name)
    for tag in ("reference-100", "GL-16", "GL-32", "GL-64", "GL-128"):
        err = np.abs(res[tag] - exact) / env
        big = pset[:, 0] * pset[:, 1] if mid == 5 else pset[:, 0] * np.maximum(pset[:, 1], 1.)
        print("  %-14s vs GL-512: max %.2e  median of per-contribution max %.2e   (contributions with q_max*L < 30: max %.2e)" % (
            tag, err.max(), np.median(err.max(1)), err[(big * q.max() < 30)].max() if (big * q.max() < 30).any() else np.nan))
    d = np.abs(res["GL-32"] - res["reference-100"]) / env
    print("  GL-32 vs reference-100: max %.2e, median of per-contribution max %.2e" % (d.max(), np.median(d.max(1))))
