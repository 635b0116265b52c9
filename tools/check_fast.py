#!/usr/bin/env python
"""Experiment driver (not part of the product): fp32 production kernel against the fp64
parity kernel on the same Philox stream -- accept sequence / chi2 trace of the first steps
under several layouts (cluster, QPT, proposal groups, speculation depth)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import common  # noqa: E402
import gpu_common  # noqa: E402
from mcsas_b200 import Engine, FP32, FP64  # noqa: E402


def run(g, om, prec, spec, cluster, max_iter, env=None, **kw):
    for k in list(os.environ):
        if k.startswith("MCSAS_B200_") and k != "MCSAS_B200_DEBUG":
            del os.environ[k]
    os.environ.update(env or {})
    eng = gpu_common.engine_for(om, [(g["q"], g["I"], g["sigma"])])
    cfg = Engine.make_config(n_contribs=g["N"], n_reps=2, conv_crit=float(g["conv_crit"]),
                             max_iter=max_iter, max_retries=-1, comp_exp=float(g["comp_exp"]),
                             precision=prec, seed=424242, record_trace=True, spec_max=spec,
                             cluster_size=cluster, **kw)
    eng.run(cfg)
    out = eng.fetch()
    tr = [eng.fetch_trace(r, max_iter) for r in range(2)]
    eng.close()
    return out, tr


def main():
    bad = 0
    for name in sys.argv[1:] or ["sphere_small", "sphere_cfg1", "sphere_nobg", "sphere_posbg", "gauss_small", "sphcs_small", "elliso_small", "cyl_small"]:
        g = common.load_chain(name)
        om = common.oracle_model(g)
        max_iter = min(int(g["max_iter"]), 3000)
        kw = dict(find_background=g["find_background"], positive_background=g["positive_background"])
        ref, rtr = run(g, om, FP64, 32, 1, max_iter, **kw)
        for spec, cluster, env in [(1, 1, {}), (32, 1, {}), (8, 2, {}), (32, 4, {}), (32, 8, {}),
                                   (32, 1, {"MCSAS_B200_QPT": "1"}), (32, 1, {"MCSAS_B200_QPT": "4"}),
                                   (32, 2, {"MCSAS_B200_PS": "2"}), (32, 4, {"MCSAS_B200_PS": "4", "MCSAS_B200_QPT": "1"})]:
            try:
                out, tr = run(g, om, FP32, spec, cluster, max_iter, env, **kw)
            except Exception as ex:
                print("%-14s spec %2d cluster %d %s: FAILED %s" % (name, spec, cluster, env, ex), flush=True)
                bad += 1
                continue
            msg = []
            for r in range(2):
                n = min(int(out["num_iter"][0, r]), int(ref["num_iter"][0, r]))
                acc, chi = tr[r][0][:n], tr[r][1][:n]
                racc, rchi = rtr[r][0][:n], rtr[r][1][:n]
                neq = np.nonzero(acc != racc)[0]
                first = int(neq[0]) if neq.size else n
                m = max(first, 1)
                rel = np.abs(chi[:m] / rchi[:m] - 1).max()
                msg.append("n %5d same-until %5d chi-rel %.1e conv0 %.2e chi %.6g/%.6g sc %.4g/%.4g bg %.3g/%.3g" % (
                    n, first, rel, abs(out["conval0"][0, r] / ref["conval0"][0, r] - 1),
                    out["conval"][0, r], ref["conval"][0, r], out["scaling"][0, r], ref["scaling"][0, r],
                    out["background"][0, r], ref["background"][0, r]))
                if first < min(n, 50) or rel > 1e-3 or not np.isfinite(out["conval"][0, r]):
                    bad += 1
                    msg[-1] += "  <-- CHECK"
            print("%-14s spec %2d cluster %d %-40s\n    %s\n    %s" % (name, spec, cluster, env, msg[0], msg[1]), flush=True)
    print("BAD", bad)


if __name__ == "__main__":
    main()
