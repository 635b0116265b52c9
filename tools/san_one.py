import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcsas_b200 import Engine, synth, FP32
q, I, sig = synth.sphere_curve(nq=1000)
eng = Engine(0)
eng.upload_datasets([(q, I, sig)])
eng.set_model(0, [10e-9, 1e14], [0], [1e-9], [100e-9], [0])
cfg = Engine.make_config(n_contribs=300, n_reps=int(os.environ.get("REPS", "4")), precision=FP32, seed=7919,
                         max_iter=int(os.environ.get("MAXIT", "3000")), max_retries=-1)
ms = eng.run(cfg)
r = eng.fetch()
print("kernel ms %.2f steps %d" % (ms, r["steps_all"].sum()), r["conval"])
