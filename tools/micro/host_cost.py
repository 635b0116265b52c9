"""Experiment: where the host time of one McSAS.analyse() goes (config 2, as bench.py builds it)."""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench as B
from mcsas_b200.mcsas import McSAS
from mcsas_b200.models import Sphere
from mcsas_b200.data import SASData
q, I, sig = B.config2_data()
data = SASData(q, I, sig)
model = Sphere()
model.radius.setActiveRange(B.RADIUS_RANGE)
algo = McSAS(device=0, seed=0, first_chain=0)
algo.model, algo.data = model, data
algo.numContribs.setValue(B.N_CONTRIBS)
algo.numReps.setValue(B.N_REPS)
algo.convergenceCriterion.setValue(1.0)
for i in range(3):
    algo.seed = 7919 * (i + 1); algo.result = []; algo.analyse()
pr = cProfile.Profile()
tk = 0.
t0 = time.perf_counter()
pr.enable()
for i in range(10):
    algo.seed = 7919 * (i + 4); algo.result = []; algo.analyse(); tk += algo.lastKernelMs
pr.disable()
t1 = time.perf_counter()
print("per analyse %.3f ms, kernel %.3f ms, host %.3f ms" % ((t1 - t0) * 100, tk / 10, (t1 - t0) * 100 - tk / 10))
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
