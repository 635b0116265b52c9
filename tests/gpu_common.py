"""Helpers for the -m gpu parity tests: drive libmcsas_b200.so through its C-ABI
(ctypes, mcsas_b200.engine.Engine) with the same inputs as the oracle."""
import numpy as np

import common
import oracle as orc


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def engine_for(om, datasets, device=0, smear=None):
    """Engine with datasets uploaded and the oracle-described model set."""
    from mcsas_b200 import Engine
    eng = Engine(device)
    eng.upload_datasets(datasets)
    eng.set_model(om.model_id, om.values, [om.pnames.index(a) for a in om.active],
                  om.lo, om.hi, om.gen_kind, om.quad_rule, om.quad_nodes)
    if smear is not None:
        eng.set_smearing(*smear)
    return eng


def run_golden_chain(g, precision, spec_max, pad_steps=8, quad_rule=0, cluster=0):
    """Replay a golden chain (uploaded generator-output stream) on the GPU."""
    from mcsas_b200 import Engine
    om = common.oracle_model(g, quad_rule=quad_rule)
    eng = engine_for(om, [(g["q"], g["I"], g["sigma"])], smear=common.smear_of(g))
    P = om.n_active
    max_iter = min(g["max_iter"], g["num_iter"] + pad_steps)
    n0 = 0 if g["start_from_minimum"] else g["N"] * P
    stream = np.full(n0 + max_iter * P, 0.5)
    stream[:g["stream"].size] = g["stream"][:stream.size]
    cfg = Engine.make_config(n_contribs=g["N"], n_reps=1, conv_crit=float(g["conv_crit"]),
                             max_iter=max_iter, comp_exp=float(g["comp_exp"]),
                             find_background=g["find_background"],
                             positive_background=g["positive_background"],
                             start_from_minimum=g["start_from_minimum"], precision=precision,
                             spec_max=spec_max, record_trace=True, cluster_size=cluster)
    eng.run(cfg, stream=stream)
    out = eng.fetch()
    acc, chi = eng.fetch_trace(0, max_iter)
    nit = int(out["num_iter"][0, 0])
    out["accept"], out["chi"] = acc[:nit], chi[:nit]
    eng.close()
    return out
