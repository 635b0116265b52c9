"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SMEAR_CHAINS = ["sphere_smear", "lma_smear2d"]      # chains recorded with data.config.smearing on
CHAIN_CASES = ["sphere_small", "sphere_nobg", "sphere_posbg", "sphere_startmin", "lma_small",
               "sphcs_small", "ellcs_small", "elliso_small", "cyl_small", "gauss_small",
               "kho_small", "sphere_cfg1"]


def load_chain(name):
    g = dict(np.load(os.path.join(GOLDEN, "chain_%s.npz" % name)))
    g["model"] = str(g["model"])
    for k in ("N", "max_iter", "seed", "num_iter", "num_moves", "stream_len"):
        g[k] = int(g[k])
    for k in ("find_background", "positive_background", "start_from_minimum"):
        g[k] = bool(g[k])
    g["accept"] = np.unpackbits(g["accept"])[:g["num_iter"]]
    if "stream" not in g:   # regenerate from the legacy numpy seed (sphere_cfg1)
        g["stream"] = regen_stream(g)
    return g


def regen_stream(g):
    """RandomUniform stream of a seeded reference run: numpy's legacy MT19937
    global stream (numbergenerator.py:31) -- initial N draws, then 1 per step."""
    rs = np.random.RandomState(g["seed"])
    s = np.concatenate([rs.uniform(size=g["N"])] +
                       [rs.uniform(size=1) for _ in range(g["num_iter"])])
    assert s.size == g["stream_len"] and abs(s.sum() - float(g["stream_sum"])) < 1e-9
    return s


def oracle_model(g, quad_rule=0):
    names = [str(x) for x in g["pnames"]]
    values = dict(zip(names, g["pvalues"]))
    active = [str(a) for a in g["active"]]
    ranges = {a: (float(lo), float(hi)) for a, lo, hi in zip(active, g["lo"], g["hi"])}
    return orc.Model(g["model"], active=active, ranges=ranges, values=values,
                     quad_rule=quad_rule)


def oracle_settings(g, **over):
    kw = dict(n_contribs=g["N"], conv_crit=float(g["conv_crit"]), max_iter=g["max_iter"],
              comp_exp=float(g["comp_exp"]), find_background=g["find_background"],
              positive_background=g["positive_background"],
              start_from_minimum=g["start_from_minimum"], q_limit_hi=float(g["q_limit_hi"]))
    kw.update(over)
    return orc.settings(**kw)


def smear_of(g):
    """(qOffset, weights, twoDColl) of a golden chain recorded with smearing, else None."""
    if "smear_off" not in g:
        return None
    return np.asarray(g["smear_off"]), np.asarray(g["smear_w"]), bool(g["smear_two_d"])
