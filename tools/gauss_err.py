import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import common, gpu_common
import oracle as orc
g = np.load(os.path.join(common.GOLDEN, "eval_models.npz"))
name = "GaussianChain"
spec = orc.MODEL_SPECS[name][1]
values = dict(zip([p[0] for p in spec], g[name + "_pvalues"]))
om = orc.Model(name, active=[str(a) for a in g[name + "_active"]], values=values)
q = g["q"]
eng = gpu_common.engine_for(om, [(q, np.ones_like(q), np.ones_like(q))])
pset = g[name + "_pset"]
it64, *_ = eng.eval(pset, 0.6666666, precision=0)
it32, *_ = eng.eval(pset, 0.6666666, precision=1)
eng.close()
rel = np.abs(it32 / it64 - 1)
print("pset", pset.ravel()[:6], "values", values)
for i in range(min(6, rel.shape[0])):
    j = rel[i].argmax()
    u = (q * pset[i, 0]) ** 2
    print("contrib %d rg %.3e: max rel %.2e at q %.3e (u %.3e); median %.1e; rel at u<0.5: %.1e, u>=0.5: %.1e" % (
        i, pset[i, 0], rel[i].max(), q[j], u[j], np.median(rel[i]), rel[i][u < 0.5].max() if (u < 0.5).any() else 0, rel[i][u >= 0.5].max() if (u >= 0.5).any() else 0))
