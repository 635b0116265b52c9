import os, sys, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, FP32, FP64
from mcsas_b200.models import CylindersIsotropic
sys.path.insert(0, os.path.join(ROOT, "tools"))
from run_configs import noisy, NM
q = np.logspace(np.log10(0.03), np.log10(3.0), 500) * 1e9
m = CylindersIsotropic()
m.aspect.setActive(True)
m.radius.setActiveRange((1 * NM, 50 * NM))
m.aspect.setActiveRange((1., 20.))
rng = np.random.default_rng(3)
radii = np.clip(rng.normal(5., 1., 400), 0.5, None) * NM
pset = np.stack([radii, np.full_like(radii, 5.0)], 1)
it = m.calc(q, pset, 1.0).cumInt
I, sig = noisy(it, 3)
for prec, name in ((FP32, "fp32"), (FP64, "fp64")):
    eng = Engine(0)
    eng.upload_datasets([(q, I, sig)])
    m.configureEngine(eng)
    cfg = Engine.make_config(n_contribs=300, n_reps=8, precision=prec, seed=99, max_iter=100000, max_retries=-1)
    t0 = time.perf_counter()
    ms = eng.run(cfg)
    r = eng.fetch()
    print(name, "kernel ms %.0f" % ms, "conval", np.round(r["conval"][0], 4), "iters", r["num_iter"][0], "conval0", np.round(r["conval0"][0][:3], 1), flush=True)
    eng.close()
