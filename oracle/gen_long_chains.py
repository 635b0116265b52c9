# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- one long chain of the ORACLE on BASELINE configs[2] (CylindersIsotropic,
radius + aspect active, 500 q, 100-node rule) and configs[3] (Kholodenko worm, 3 active parameters,
500 q, adaptive Gauss-Kronrod integral) on the counter-based Philox stream the CUDA kernels draw
from, so that the GPU chains can be checked step by step at the full problem size (do
those configurations really not converge, or is it the production path?).  Hours of CPU for the
worm; run once in the build container, outputs under tests/golden/.

  python oracle/gen_long_chains.py cyl|kho [steps]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
NM = 1e-9
SEED, CHAIN, N = 0x5EED0C0FFEE, 3, 300


def data_cyl():
    q = np.logspace(np.log10(0.03), np.log10(3.0), 500) * 1e9
    om = orc.Model("CylindersIsotropic", active=["radius", "aspect"],
                   ranges={"radius": (1 * NM, 50 * NM), "aspect": (1., 20.)})
    rng = np.random.default_rng(3)
    radii = np.clip(rng.normal(5., 1., 400), 0.5, None) * NM
    it, _, _, _ = orc.eval_contribs(om, q, np.stack([radii, np.full_like(radii, 5.0)], 1), 1.0)
    I = it.sum(0)
    I = I / I[0] * 1e6 + 1.0
    sig = 0.01 * I
    return om, q, I + rng.normal(0., 1., I.size) * sig, sig


def data_kho():
    q = np.logspace(np.log10(0.01), np.log10(10.0), 500) * 1e9
    om = orc.Model("Kholodenko", active=["radius", "lenKuhn", "lenContour"])
    it, _, _, _ = orc.eval_contribs(om, q, np.array([[1 * NM, 10 * NM, 1000 * NM]]), 1.0)
    rng = np.random.default_rng(4)
    I = it.sum(0)
    I = I / I[0] * 1e6 + 1.0
    sig = 0.01 * I
    return om, q, I + rng.normal(0., 1., I.size) * sig, sig


def run(tag, steps):
    om, q, I, sig = data_cyl() if tag == "cyl" else data_kho()
    st = orc.settings(n_contribs=N, conv_crit=1.0, max_iter=steps)
    stream = orc.philox_stream(SEED, CHAIN, N, om, False, steps)
    t0 = time.time()
    o = orc.chain(om, st, q, I, sig, stream, trace=True)
    el = time.time() - t0
    acc, chi = np.asarray(o["accept"], dtype=np.uint8), np.asarray(o["chi"])
    # running chi-squared after every step (the accepted values)
    run_chi = np.empty(acc.size)
    cur = o["conval0"]
    for t in range(acc.size):
        if acc[t]:
            cur = chi[t]
        run_chi[t] = cur
    np.savez_compressed(os.path.join(OUT, "long_%s.npz" % tag), model=np.array(om.name), q=q, I=I, sigma=sig,
                        active=np.array(om.active), lo=np.array(om.lo), hi=np.array(om.hi), seed=SEED, chain=CHAIN,
                        N=N, max_iter=steps, num_iter=o["num_iter"], num_moves=o["num_moves"], conval0=o["conval0"],
                        conval=o["conval"], accept=np.packbits(acc), run_chi_every_100=run_chi[99::100],
                        rset=o["rset"], elapsed_s=el)
    print("%s: %d steps, %d moves, chi2 %.6g -> %.6g in %.0f s (%.1f steps/s)" % (
        tag, o["num_iter"], o["num_moves"], o["conval0"], o["conval"], el, o["num_iter"] / el))


if __name__ == "__main__":
    tag = sys.argv[1]
    run(tag, int(sys.argv[2]) if len(sys.argv) > 2 else (100000 if tag == "cyl" else 20000))
