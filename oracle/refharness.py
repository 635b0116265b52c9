# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Headless import of the *unmodified* reference (BAMresearch/McSAS v1.3.1,
mounted read-only at /root/reference) so that golden vectors can be
generated in the build container.  /root/reference does not exist on the
GPU box: only ``oracle/gen_golden.py`` (run here, output committed under
``tests/golden/``) and container-only tests call into this module.

The reference needs four ``sys.modules`` stubs:
``future`` (scatteringmodel.py:7), ``QtWidgets`` (gui/utils/__init__.py:4-7,
imported by mcsas.py:23), ``matplotlib`` (mcsas/plotting.py:11) and ``h5py``
(utils/hdf.py:11).  ``numpy.linspace`` is wrapped for CylindersIsotropic,
whose ``intDiv`` is the float ``100.`` (cylindersisotropic.py:56).
"""
import logging
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    """/root/reference (build container), or the archive __graft_entry__.build() leaves in the
    git-ignored oracle/_ref/ so that the reference arm of bench.py can time the UNMODIFIED reference
    on the GPU box's host cores (it travels with gpurun like the built .so files; never committed,
    never unpacked inside the repository: the archive is extracted to a temporary directory, once
    per process tree)."""
    if os.path.isdir("/root/reference/src/mcsas"):
        return "/root/reference/src", "/root/reference/testdata"
    root = os.environ.get("MCSAS_B200_REFDIR")
    if not (root and os.path.isdir(os.path.join(root, "src", "mcsas"))):
        arc = os.path.join(_HERE, "_ref", "reference.tar.gz")
        if not os.path.exists(arc):
            return "/root/reference/src", "/root/reference/testdata"
        import atexit
        import shutil
        import tarfile
        import tempfile
        root = tempfile.mkdtemp(prefix="mcsas_reference_")
        with tarfile.open(arc) as tf:
            tf.extractall(root)
        os.environ["MCSAS_B200_REFDIR"] = root            # child processes reuse it
        atexit.register(shutil.rmtree, root, True)
    return os.path.join(root, "src"), os.path.join(root, "testdata")


REFERENCE_SRC, REFERENCE_TESTDATA = _find_reference()


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "mcsas"))


_installed = False


def install():
    """Install the stubs and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not mounted at " + REFERENCE_SRC)
    fut = types.ModuleType("future")
    futu = types.ModuleType("future.utils")
    sl = types.ModuleType("future.standard_library")

    def with_metaclass(meta, *bases):
        class metaclass(type):
            def __new__(cls, name, this_bases, d):
                return meta(name, bases, d)

            @classmethod
            def __prepare__(cls, name, this_bases):
                return meta.__prepare__(name, bases)
        return type.__new__(metaclass, "temporary_class", (), {})

    futu.with_metaclass = with_metaclass
    sl.install_aliases = lambda: None
    fut.utils, fut.standard_library = futu, sl
    sys.modules.update({"future": fut, "future.utils": futu,
                        "future.standard_library": sl})
    qtw = types.ModuleType("QtWidgets")
    qtw.QApplication = type("QApplication", (),
                            {"processEvents": staticmethod(lambda: None)})
    sys.modules["QtWidgets"] = qtw
    for n in ("matplotlib", "matplotlib.font_manager", "matplotlib.gridspec",
              "matplotlib.pyplot", "h5py"):
        sys.modules[n] = MagicMock(name=n)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    # numpy >= 2: linspace(num=100.) raises TypeError -> coerce, reference untouched
    _linspace = np.linspace

    def linspace(start, stop, num=50, *a, **kw):
        return _linspace(start, stop, int(num), *a, **kw)
    np.linspace = linspace
    # same numpy >= 2 incompatibility in sasconfig.py:133,220 (logspace(num=np.ceil(n / 2.)))
    _logspace = np.logspace

    def logspace(start, stop, num=50, *a, **kw):
        return _logspace(start, stop, int(num), *a, **kw)
    np.logspace = logspace
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    import mcsas.mcsas  # noqa: F401  (calls logging.basicConfig(INFO), mcsas.py:15)
    logging.getLogger().setLevel(logging.WARNING)
    _installed = True


def model_class(name):
    """Return the reference model class by our canonical name."""
    install()
    import importlib
    modname = {
        "Sphere": "sphere", "LMADenseSphere": "lmadensesphere",
        "SphericalCoreShell": "sphericalcoreshell",
        "EllipsoidalCoreShell": "ellipsoidalcoreshell",
        "EllipsoidsIsotropic": "ellipsoidsisotropic",
        "CylindersIsotropic": "cylindersisotropic",
        "GaussianChain": "gaussianchain", "Kholodenko": "kholodenko",
    }[name]
    return getattr(importlib.import_module("mcsas.models." + modname), name)


def make_data(q_nm, intensity, sigma, nbin=0):
    """SASData from arrays (q in nm^-1), optional re-binning disabled with
    nbin=0 (dataconfig.py:99); headless equivalent of FileList.loadData
    (gui/filelist.py:34-69)."""
    install()
    from mcsas.dataobj import SASData
    d = SASData(title="synthetic",
                rawArray=np.stack([q_nm, intensity, sigma], 1))
    d.config.nBin.setValue(nbin)
    d.config.updateFuMin()
    d.setConfig(d.config)
    return d


def load_data(fn, nbin=None):
    install()
    from mcsas.datafile import loaddatafile
    d = loaddatafile(fn).getDataObj()
    if nbin is not None:
        d.config.nBin.setValue(nbin)
    d.config.updateFuMin()
    d.setConfig(d.config)
    return d


def make_algo(model, data, **settings):
    """McSAS instance wired like Calculator.__call__ (gui/calc.py:300-311)."""
    install()
    from mcsas.mcsas import McSAS
    algo = McSAS.factory()()
    algo.model = model
    algo.data = data
    for k, v in settings.items():
        getattr(algo, k).setValue(v)
    algo.stop = False
    algo.result = []
    return algo


class RecordingRNG(object):
    """Context manager that records every numpy.random.uniform draw made by
    the reference (numbergenerator.py:31,173) as flat raw values."""

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self._orig = np.random.uniform
        rec = self

        def uniform(low=0.0, high=1.0, size=None):
            r = rec._orig(low, high, size)
            rec.draws.append(np.atleast_1d(np.asarray(r, dtype=np.float64)).copy())
            return r
        np.random.uniform = uniform
        return self

    def __exit__(self, *a):
        np.random.uniform = self._orig

    def flat(self):
        return np.concatenate(self.draws) if self.draws else np.zeros(0)


class RecordingGenerators(object):
    """Records the OUTPUT of every number-generator call the reference makes
    (RandomUniform.get numbergenerator.py:28-31, RandomExponential.get
    :168-175) in consumption order -- the "proposal stream" that the oracle
    and the kernel's deterministic mode replay."""

    def __init__(self):
        self.chunks = []

    def __enter__(self):
        install()
        from mcsas.bases.algorithm import numbergenerator as ng
        self._ng = ng
        self._orig = {}
        rec = self
        for cls in (ng.RandomUniform, ng.RandomExponential):
            orig = cls.__dict__["get"]
            self._orig[cls] = orig

            def make(orig_cm):
                def get(c, count=1):
                    r = orig_cm.__func__(c, count)
                    rec.chunks.append(np.atleast_1d(np.asarray(r, dtype=np.float64)).copy())
                    return r
                return classmethod(get)
            setattr(cls, "get", make(orig))
        return self

    def __exit__(self, *a):
        for cls, orig in self._orig.items():
            setattr(cls, "get", orig)

    def flat(self):
        return np.concatenate(self.chunks) if self.chunks else np.zeros(0)


class RecordingFit(object):
    """Records (conval, sc) of every BackgroundScalingFit.calc call
    (backgroundscalingfit.py:112-139): 2 initial calls, 1 per MC step, 1 final."""

    def __init__(self):
        self.conval, self.sc, self.agofs = [], [], []

    def __enter__(self):
        install()
        from mcsas.mcsas.backgroundscalingfit import BackgroundScalingFit as B
        self._B = B
        self._orig = B.calc
        rec = self

        def calc(bself, data, modelData, sc, ver=2):
            out = rec._orig(bself, data, modelData, sc, ver=ver)
            rec.sc.append(np.array(out[0], dtype=np.float64).copy())
            rec.conval.append(float(out[1]))
            rec.agofs.append(float(out[3]))
            return out
        B.calc = calc
        return self

    def __exit__(self, *a):
        self._B.calc = self._orig
