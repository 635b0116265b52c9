"""CPU tests of the host logic: parameter / histogram mirrors against vectors from the
reference, the C-ABI library (loads, exports every declared symbol, fails loudly without
a GPU), partitioning and the gloo world_size=2 gather."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NM = 1e-9


def test_cabi_exports_every_declared_symbol():
    from mcsas_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mcsas_b200.h")).read()
    declared = set(re.findall(r"\b(mcsas_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_model_ids_match_oracle_and_header():
    import oracle as orc
    from mcsas_b200.models import MODELS
    hdr = open(os.path.join(ROOT, "include", "mcsas_b200.h")).read()
    ids = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"MCSAS_([A-Z_]+) = (\d+),", hdr))
    for name, cls in MODELS.items():
        oid, plist, active = orc.MODEL_SPECS[name]
        assert cls.modelId == oid
        m = cls()
        assert [p.name() for p in m.params()] == [p[0] for p in plist]
        np.testing.assert_allclose([float(p()) for p in m.params()], [float(p[1]) for p in plist],
                                   rtol=1e-14)
        assert [p.name() for p in m.activeParams()] == active
        for p, spec in zip(m.params(), plist):
            if hasattr(p, "generator"):
                assert p.generator().kind == spec[2], (name, p.name())
            if spec[3] is not None:
                np.testing.assert_allclose(p.activeRange(), spec[3], rtol=1e-14)
    assert ids["SPHERE"] == 0 and ids["KHOLODENKO"] == 7


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    from mcsas_b200 import Engine, McSASB200Error
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere
    from mcsas_b200.data import SASData
    with pytest.raises(McSASB200Error, match="no CPU fallback"):
        Engine(0)
    algo = McSAS()
    algo.model = Sphere()
    algo.data = SASData(np.linspace(1e8, 1e9, 10), np.ones(10), np.ones(10) * 0.01)
    with pytest.raises(McSASB200Error):
        algo.calc()


def test_user_defined_model_rejected():
    from mcsas_b200.models import ScatteringModel

    class Mine(ScatteringModel):
        pass
    with pytest.raises(NotImplementedError):
        Mine()


def test_parameter_surface():
    from mcsas_b200.models import Sphere, CylindersIsotropic
    from mcsas_b200.parameter import ValueRangeError, isActiveFitParam
    m = Sphere()
    assert m.radius() == pytest.approx(10e-9) and m.radius.isActive()
    assert m.radius.activeRange() == pytest.approx((1e-9, 1e-6))
    m.radius.setActiveRange((100e-9, 1e-9))                 # order-insensitive (:633)
    assert m.radius.activeRange() == pytest.approx((1e-9, 100e-9))
    m.radius.setActiveRange((-5., 3e-9))                    # clipped to valueRange (:631)
    assert m.radius.activeRange() == pytest.approx((0., 3e-9))
    with pytest.raises(ValueRangeError):
        m.radius.setActiveRange((1, 2, 3))
    m.radius.setValue(-1.)                                  # setValue clips (parameter.py:405)
    assert m.radius() == 0.
    m.radius.setActiveVal(np.arange(3.), index=2)           # list grows (utils/parameter.py:685)
    assert len(m.radius.activeValues()) == 3 and m.radius.activeValues()[0] is None
    assert not isActiveFitParam(m.sld)
    c = CylindersIsotropic()
    c.aspect.setActive(True)
    assert c.activeParamNames() == ["radius", "aspect"]
    with pytest.raises(ValueError, match="finite"):
        c.kernelDescription()                               # radius has no activeRange preset
    c.radius.setActiveRange((1e-9, 50e-9)); c.aspect.setActiveRange((1., 20.))
    d = c.kernelDescription()
    assert d["active_idx"] == [0, 3] and d["gen_kind"] == [1, 1] and d["quad_nodes"] == 100
    np.random.seed(3)
    g = c.generateParameters(5)
    assert g.shape == (5, 2) and (g[:, 0] >= 1e-9).all() and (g[:, 0] < 50e-9).all()


def test_histogram_and_moments_match_reference():
    from mcsas_b200.models import Sphere
    from mcsas_b200.parameter import Histogram
    g = np.load(os.path.join(common.GOLDEN, "histogram_sphere.npz"))
    m = Sphere()
    m.radius.setActiveRange((1 * NM, 100 * NM))
    fractions = dict((k, (g["frac_" + k], g["minreq_" + k])) for k in ("vol", "num", "int", "surf"))
    hs = [Histogram(m.radius, 1 * NM, 100 * NM, 20, 'log', 'vol'),
          Histogram(m.radius, 2 * NM, 60 * NM, 12, 'lin', 'num', autoFollow=False)]
    for i, h in enumerate(hs):
        m.radius.histograms().append(h)
    m.radius.histograms().calc(g["contribs"], 0, fractions)
    for i, h in enumerate(hs):
        np.testing.assert_allclose(h.xLowerEdge, g["h%d_edges" % i], rtol=1e-13)
        np.testing.assert_allclose(h.bins.mean, g["h%d_mean" % i], rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose(h.bins.std, g["h%d_std" % i], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(h.cdf.mean, g["h%d_cdf" % i], rtol=1e-11)
        np.testing.assert_allclose(h.observability, g["h%d_obs" % i], rtol=1e-11)
        np.testing.assert_allclose(h.moments.fields, g["h%d_moments" % i], rtol=1e-10)


def test_oracle_fractions_match_reference():
    import oracle as orc
    g = np.load(os.path.join(common.GOLDEN, "histogram_sphere.npz"))
    om = orc.Model("Sphere", ranges={"radius": (1 * NM, 100 * NM)})
    st = orc.settings(n_contribs=g["contribs"].shape[0], comp_exp=float(g["comp_exp"]))
    for r in range(g["contribs"].shape[2]):
        f = orc.fractions(om, st, g["q"], g["I"], g["sigma"], g["contribs"][:, :, r])
        for mine, ref in (("vol", "vol"), ("num", "num"), ("volsqr", "int"), ("surf", "surf")):
            np.testing.assert_allclose(f[mine], g["frac_" + ref][:, r], rtol=2e-5)
            np.testing.assert_allclose(f["min_" + mine], g["minreq_" + ref][:, r], rtol=2e-5)


def test_partition():
    from mcsas_b200.dist import partition
    for n, w in ((100, 8), (7, 2), (3, 4), (4096 * 100, 8)):
        parts = [partition(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (f0, c0), (f1, c1) in zip(parts, parts[1:]):
            assert f1 == f0 + c0 and abs(c0 - c1) <= 1


_GLOO_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as dist
from mcsas_b200 import dist as mdist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
N, P, R = 6, 2, 3
first, cnt = mdist.partition(world * R, world, rank)
rng = np.random.default_rng(rank)
res = dict(contribs=rng.random((N, P, R)) + rank, scalingAll=np.full(R, 1.0 + rank),
           backgroundAll=np.full(R, -1.0 * rank), conval=np.full(R, 0.5),
           numIterAll=np.arange(R) + 10 * rank, numMoves=np.arange(R))
g = mdist.gather_results(res)
assert g["contribs"].shape == (N, P, world * R), g["contribs"].shape
for r in range(world):
    exp = np.random.default_rng(r).random((N, P, R)) + r
    assert np.array_equal(g["contribs"][:, :, r * R:(r + 1) * R], exp)
    assert (g["scalingAll"][r * R:(r + 1) * R] == 1.0 + r).all()
    assert (g["numIterAll"][r * R:(r + 1) * R] == np.arange(R) + 10 * r).all()
# ragged blocks (partition sizes differ by one), fixed-capacity asynchronous form: start() does not
# wait for the other rank, two gathers back to back reuse the pinned buffers
R2 = 2 + rank
def block(seed):
    r = np.random.default_rng(seed)
    return dict(contribs=r.random((N, P, R2)) + rank, scalingAll=np.full(R2, 2.0 + rank),
                backgroundAll=np.zeros(R2), conval=np.full(R2, 0.7), numIterAll=np.arange(R2) + 100 * rank,
                numMoves=np.arange(R2))
G = mdist.Gatherer(0, max_rows=3, n_contribs=N, n_active=P)
for seed in (5, 6):
    G.start(block(seed + 10 * rank))
    g = G.wait()
    assert g["contribs"].shape == (N, P, 2 + 3), g["contribs"].shape
    off = 0
    for r in range(world):
        exp = np.random.default_rng(seed + 10 * r).random((N, P, 2 + r)) + r
        assert np.array_equal(g["contribs"][:, :, off:off + 2 + r], exp)
        assert (g["numIterAll"][off:off + 2 + r] == np.arange(2 + r) + 100 * r).all()
        off += 2 + r
# a rank without a result (fit not converged, showIncomplete off): EVERY rank raises, none hangs
G.start(block(1) if rank == 0 else None)
try:
    G.wait()
    raise SystemExit("expected an error on rank %%d" %% rank)
except RuntimeError as ex:
    assert "hold no result" in str(ex)
try:
    mdist.gather_results(None if rank == 0 else block(2))
    raise SystemExit("expected an error on rank %%d" %% rank)
except RuntimeError as ex:
    assert "hold no result" in str(ex)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world2_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
                          "29731", str(script)], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


@pytest.mark.parametrize("tag", ["sphere10_default", "quick_default", "quick_nbin50_lim", "quick_nobin"])
def test_dataprep_matches_reference(tag):
    """mcsas_b200.dataprep against the vectors the unmodified reference hands
    to mcFit after loaddatafile -> updateFuMin -> setConfig (golden: oracle/gen_golden.py dataprep)."""
    from mcsas_b200 import dataprep
    g = np.load(os.path.join(common.GOLDEN, "dataprep.npz"))
    raw, cfg = g[tag + "_raw"], g[tag + "_cfg"]
    d = dataprep.prepare(raw[:, 0], raw[:, 1], raw[:, 2], n_bin=cfg[0], fu_min=cfg[1], x0_low=cfg[2],
                         x0_high=cfg[3], mask_zero=bool(cfg[4]), mask_neg=bool(cfg[5]))
    assert d.q.size == g[tag + "_q"].size
    np.testing.assert_allclose(d.q, g[tag + "_q"], rtol=1e-14)
    np.testing.assert_allclose(d.f.binnedData, g[tag + "_I"], rtol=1e-14)
    np.testing.assert_allclose(d.f.binnedDataU, g[tag + "_sigma"], rtol=1e-13)


def test_dataprep_edge_cases():
    from mcsas_b200 import dataprep
    q = np.array([0.1, 0.2, 0.3, 0.4])
    I = np.array([10., np.nan, 0., -1.])
    d = dataprep.prepare(q, I, None, n_bin=0, mask_zero=True, mask_neg=True)
    assert d.q.size == 1 and d.f.binnedDataU[0] == pytest.approx(0.1)       # fuMin * I, no sigma column
    d = dataprep.prepare(q, np.ones(4), np.zeros(4), n_bin=2)
    assert d.q.size == 2 and (d.f.binnedDataU >= 0.01).all()                # floor at 1 % of I
    with pytest.raises(ValueError):
        dataprep.prepare(q, np.ones(3))


def test_output_tables_match_reference_layout(tmp_path):
    """mcsas_b200.output: byte-identical to the reference's AsciiFile writer
    (golden text recorded by oracle/gen_golden.py output)."""
    from mcsas_b200 import output
    g = np.load(os.path.join(common.GOLDEN, "output_formats.npz"))
    res = dict((c, g["col_" + c]) for c in output.FIT_COLUMNS)
    fn = output.write_fit(str(tmp_path / "x_fit.dat"), res)
    assert open(fn).read() == str(g["fit_txt"])
    cells = [[output.format_value(1e-9), output.format_value(5e-9)], [output.format_value(1e-7)] * 2,
             [output.format_value("vol"), output.format_value("num")],
             [output.format_value(9.7455e-9), output.format_value(-1.5)]]
    fn = output.write_table(str(tmp_path / "x_stats.dat"), ("lower", "upper", "weighting", "mean"), cells)
    assert open(fn).read() == str(g["stats_txt"])


def test_output_files_from_a_result(tmp_path):
    """Distribution / statistics / contributions / settings files from host objects (no GPU:
    the histogram is fed with a synthetic contribution set)."""
    import pickle
    from mcsas_b200 import output
    from mcsas_b200.mcsas import McSAS
    from mcsas_b200.models import Sphere
    from mcsas_b200.parameter import Histogram
    rng = np.random.default_rng(0)
    N, R = 50, 4
    contribs = rng.uniform(2e-9, 9e-8, (N, 1, R))
    frac = rng.uniform(0.1, 1., (N, R))
    algo = McSAS.factory()()
    algo.model = Sphere()
    p = algo.model.radius
    p.setActiveRange((1e-9, 1e-7))
    p.histograms().append(Histogram(p, 1e-9, 1e-7, 10, 'log', 'vol'))
    fractions = dict(vol=(frac, frac * 0.1), num=(frac, frac * 0.1), int=(frac, frac * 0.1), surf=(frac, frac * 0.1))
    p.histograms().calc(contribs, 0, fractions)
    h = p.histograms()[0]
    lines = open(output.write_distribution(str(tmp_path / "h.dat"), h)).read().split("\n")
    assert lines[0] == "xMean xWidth yMean yStd Obs cdfMean cdfStd" and len(lines) == 11
    assert all(len(l) == 7 * 14 + 6 for l in lines[1:])
    lines = open(output.write_statistics(str(tmp_path / "s.dat"), p)).read().split("\n")
    assert lines[0].split()[:3] == ["lower", "upper", "weighting"] and len(lines) == 2
    assert " vol " in lines[1]
    fn = output.write_contributions(str(tmp_path / "c.pickle"), dict(contribs=contribs))
    assert np.array_equal(pickle.load(open(fn, "rb")), contribs)
    txt = open(output.write_settings(str(tmp_path / "s.cfg"), algo, "in.dat", "out", (1e7, 1e9))).read()
    assert "[MCSAS Settings]" in txt and "numcontribs = 300" in txt.lower() and "radius_min" in txt


def test_precision_follows_the_data_uncertainty():
    """Production (fp32) arithmetic only where the relative uncertainties are well above the
    fp32 form-factor accuracy; an explicit choice is kept."""
    from mcsas_b200 import _lib
    from mcsas_b200.data import SASData
    from mcsas_b200.mcsas import McSAS
    q = np.logspace(-1, 0, 20) * 1e9
    I = np.linspace(10., 1., 20)
    algo = McSAS(seed=1)
    algo.data = SASData(q, I, 0.01 * I)
    assert algo._resolvePrecision() == _lib.FP32
    algo.data = SASData(q, I, 1e-4 * I)
    assert algo._resolvePrecision() == _lib.FP64
    algo.precision = _lib.FP32
    assert algo._resolvePrecision() == _lib.FP32


def _integration_stub():
    """The ctypes stub a McSAS maintainer would add (way B), executed as written."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    part = text[text.index("## B."):]
    code = part[part.index("```python") + len("```python"):]
    code = code[:code.index("```")]
    ns = {}
    exec(compile(code, "INTEGRATION.md#B", "exec"), ns)
    return ns


def test_integration_stub_matches_the_header_and_the_binding():
    """INTEGRATION.md way B is documentation that must stay executable: its RunConfig is the
    header's struct mcsas_run_config (field order, C types, size) and the package's own binding;
    its model / generator tables are the header's enums."""
    import ctypes as C
    import re
    from mcsas_b200 import _lib
    ns = _integration_stub()
    stub = ns["RunConfig"]
    assert [(n, t) for n, t in stub._fields_] == [(n, t) for n, t in _lib.RunConfig._fields_]
    assert C.sizeof(stub) == C.sizeof(_lib.RunConfig)
    hdr = open(os.path.join(ROOT, "include", "mcsas_b200.h")).read()
    body = hdr[hdr.index("typedef struct mcsas_run_config {"):hdr.index("} mcsas_run_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    ctype = {"int": C.c_int, "double": C.c_double, "long long": C.c_longlong,
             "unsigned long long": C.c_ulonglong}
    fields = [(m.group(2), ctype[m.group(1).strip()])
              for m in re.finditer(r"((?:unsigned )?(?:long long|int|double))\s+(\w+)\s*;", body)]
    assert fields == list(stub._fields_)
    enum = hdr[hdr.index("enum mcsas_model_id {"):hdr.index("MCSAS_N_MODELS")]
    enum = re.sub(r"/\*.*?\*/", "", enum, flags=re.S)
    ids = {m.group(1): int(m.group(2)) for m in re.finditer(r"(MCSAS_\w+)\s*=\s*(\d+)", enum)}
    names = {"Sphere": "MCSAS_SPHERE", "LMADenseSphere": "MCSAS_LMA_DENSE_SPHERE",
             "SphericalCoreShell": "MCSAS_SPHERICAL_CORE_SHELL", "EllipsoidalCoreShell": "MCSAS_ELLIPSOIDAL_CORE_SHELL",
             "EllipsoidsIsotropic": "MCSAS_ELLIPSOIDS_ISOTROPIC", "CylindersIsotropic": "MCSAS_CYLINDERS_ISOTROPIC",
             "GaussianChain": "MCSAS_GAUSSIAN_CHAIN", "Kholodenko": "MCSAS_KHOLODENKO"}
    assert {k: ids[v] for k, v in names.items()} == ns["MODEL_ID"]
    assert ns["GEN_KIND"]["RandomUniform"] == _lib.GEN_UNIFORM and ns["GEN_KIND"]["RandomExponential"] == _lib.GEN_EXP1
    assert ns["GEN_KIND"]["RandomExponential2"] == _lib.GEN_EXP2 and ns["GEN_KIND"]["RandomExponential3"] == _lib.GEN_EXP3
    # every entry point the stub calls is exported
    L = _lib.lib()
    for name in set(re.findall(r"L\.(mcsas_\w+)", open(os.path.join(ROOT, "INTEGRATION.md")).read())):
        assert hasattr(L, name), name
