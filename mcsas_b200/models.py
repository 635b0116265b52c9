# -*- coding: utf-8 -*-
"""Host mirror of McSAS's ScatteringModel plug-in surface.

  ScatteringModel / SASModel   src/mcsas/bases/model/scatteringmodel.py:14-243, sasmodel.py:11-81
  ModelData                    src/mcsas/bases/model/modeldata.py:4-66
  model plug-ins               src/mcsas/models/{sphere,lmadensesphere,sphericalcoreshell,
                               ellipsoidalcoreshell,ellipsoidsisotropic,cylindersisotropic,
                               gaussianchain,kholodenko}.py
Parameter tuples (names, defaults, value/active ranges, generators) repeat the
reference's declarations; the arithmetic (formfactor, calcIntensity, calc) runs on the
GPU through the C-ABI (mcsas_eval) -- there is no numpy implementation of the form
factors in this package.  User-defined Python models cannot run inside a fixed kernel
and raise (north_star: no CPU fallback).
"""
import numpy as np

from . import _lib
from .data import hot_path_vectors
from .engine import Engine
from .parameter import (FitParameter, Parameter, RandomUniform, RandomExponential, NM, SLD,
                        Percent, NoUnit, isActiveFitParam, isFitParam)


class ModelData(object):
    """modeldata.py:4-66"""

    def __init__(self, cumInt, vset, wset, sset, numParams):
        self._cumInt = np.asarray(cumInt).flatten()
        self._vset = np.asarray(vset).flatten()
        self._wset = np.asarray(wset).flatten()
        self._sset = np.asarray(sset).flatten()
        self._numParams = abs(numParams)

    cumInt = property(lambda self: self._cumInt)
    chisqrInt = property(lambda self: self._cumInt)
    vset = property(lambda self: self._vset)
    wset = property(lambda self: self._wset)
    sset = property(lambda self: self._sset)
    numParams = property(lambda self: self._numParams)

    def volumeFraction(self, scaling):
        return (self.wset * scaling / self.vset).flatten()


SASModelData = ModelData


class ScatteringModel(object):
    """Base of the model plug-ins.  Subclasses declare `modelId` (kernel selector),
    `shortName` and `parameterSpec` (a callable returning fresh parameter objects)."""
    modelId = None
    shortName = None
    canSmear = False
    compensationExponent = None
    quadRule = _lib.QUAD_REFERENCE      # orientation average: the reference's own rule
    quadNodes = None                    # None -> intDiv

    def __init__(self):
        if self.modelId is None:
            raise NotImplementedError(
                "user-defined ScatteringModel subclasses cannot run in the fixed CUDA kernel; "
                "mcsas_b200 has no CPU fallback")
        self._params = list(self.parameterSpec())
        for p in self._params:
            setattr(self, p.name(), p)
        self._engine = None

    @classmethod
    def factory(cls):
        return cls

    # -- parameter bookkeeping (algorithmbase.py, scatteringmodel.py:117-185) --------
    def params(self):
        return self._params

    def paramCount(self):
        return len(self._params)

    def param(self, index):
        return self._params[index]

    def activeParams(self):
        return [p for p in self._params if isActiveFitParam(p)]

    def activeParamCount(self):
        return len(self.activeParams())

    def activeParamNames(self):
        return [p.name() for p in self.activeParams()]

    def generateParameters(self, count=1):
        lst = np.zeros((count, self.activeParamCount()))
        for idx, param in enumerate(self.activeParams()):
            lst[:, idx] = param.generate(count=count)
        return lst

    def name(self):
        return self.shortName

    # -- description handed to the kernel ---------------------------------------------
    def kernelDescription(self):
        """(model_id, params[SI], active_idx, lo, hi, generator kinds, quad rule, nodes)."""
        values = np.array([float(p()) for p in self._params])
        idx, lo, hi, kinds = [], [], [], []
        for i, p in enumerate(self._params):
            if not isActiveFitParam(p):
                continue
            a = p.activeRange()
            v = p.valueRange()
            idx.append(i)
            lo.append(max(min(a), v[0]))           # generateValues, parameter.py:74-75
            hi.append(min(max(a), v[1]))
            kinds.append(p.generator().kind)
        if any(not np.isfinite(x) for x in lo + hi):
            raise ValueError("active range of %s must be finite: set it with setActiveRange()"
                             % self.shortName)
        nodes = self.quadNodes
        if nodes is None:
            nodes = int(getattr(self, "intDiv")()) if hasattr(self, "intDiv") else 0
        return dict(model_id=self.modelId, params=values, active_idx=idx, lo=lo, hi=hi,
                    gen_kind=kinds, quad_rule=self.quadRule, quad_nodes=int(nodes))

    def configureEngine(self, engine):
        d = self.kernelDescription()
        engine.set_model(d["model_id"], d["params"], d["active_idx"], d["lo"], d["hi"],
                         d["gen_kind"], d["quad_rule"], d["quad_nodes"])

    # -- arithmetic: all on the GPU ----------------------------------------------------
    def _evalEngine(self, data):
        if self._engine is None:
            self._engine = Engine(0)
        q, I, sigma = hot_path_vectors(data) if not isinstance(data, np.ndarray) else (
            np.asarray(data, dtype=np.float64).ravel(), None, None)
        if I is None:
            I = np.ones_like(q)
            sigma = np.ones_like(q)
        self._engine.upload_datasets([(q, I, sigma)])
        self.configureEngine(self._engine)
        from . import smearing
        prof = None if isinstance(data, np.ndarray) else smearing.profile_for(data, self)
        if prof is None:
            self._engine.set_smearing(None)
        else:
            self._engine.set_smearing(*prof)
        return self._engine

    def calc(self, data, pset, compensationExponent=None, precision=_lib.FP64):
        """ScatteringModel.calc (scatteringmodel.py:79-105): summed intensity of the
        contributions in pset[M, nActive] plus their v/w/s sets."""
        if compensationExponent is None:
            compensationExponent = self.compensationExponent
        pset = np.atleast_2d(np.asarray(pset, dtype=np.float64))
        eng = self._evalEngine(data)
        it, v, w, s = eng.eval(pset, compensationExponent, precision=precision)
        cumInt = np.zeros(it.shape[1])
        for row in it:                       # cumInt += it, contribution order (:100)
            cumInt += row
        return self.getModelData(cumInt, v, w, s)

    def getModelData(self, cumInt, vset, wset, sset):
        return ModelData(cumInt, vset, wset, sset, self.activeParamCount())

    def calcIntensity(self, data, compensationExponent=None):
        """SASModel.calcIntensity (sasmodel.py:46-79) at the current parameter values."""
        if compensationExponent is None:
            compensationExponent = self.compensationExponent
        pv = np.array([[float(p()) for p in self.activeParams()]])
        eng = self._evalEngine(data)
        it, v, w, s = eng.eval(pv, compensationExponent)
        return it[0], v[0], w[0], s[0]

    def formfactor(self, dataset):
        """|F(q)| at the current parameter values (the reference returns the signed
        amplitude; only F^2 enters the fit, sasmodel.py:77)."""
        pv = np.array([[float(p()) for p in self.activeParams()]])
        eng = self._evalEngine(dataset)
        it, v, w, s = eng.eval(pv, 0.0)      # w = volume^0 = 1
        return np.sqrt(it[0])

    def weight(self):
        return self.volume() ** (2 * self.compensationExponent)

    def absVolume(self):
        return self.volume()

    def surface(self):
        return 0


class SASModel(ScatteringModel):
    pass


def _p(name, value, **kw):
    return Parameter(name, value, **kw)


def _f(name, value, **kw):
    return FitParameter(name, value, **kw)


class Sphere(SASModel):
    """sphere.py:12-63"""
    modelId = 0
    shortName = "Sphere"
    canSmear = True

    @staticmethod
    def parameterSpec():
        return (_f("radius", NM.toSi(10.), unit=NM, displayName="Sphere radius",
                   valueRange=(0., np.inf), activeRange=NM.toSi((1., 1000.)),
                   generator=RandomUniform, decimals=9),
                _p("sld", SLD.toSi(1e-6), unit=SLD,
                   displayName="scattering length density difference",
                   valueRange=(0., np.inf), decimals=9))

    def __init__(self):
        super(Sphere, self).__init__()
        self.radius.setActive(True)

    def surface(self):
        return 4. * np.pi * self.radius() * self.radius()

    def volume(self):
        return (np.pi * 4. / 3.) * self.radius() ** 3

    def absVolume(self):
        return self.volume() * self.sld() ** 2


class LMADenseSphere(SASModel):
    """lmadensesphere.py:13-102"""
    modelId = 1
    shortName = "LMADenseSphere"
    canSmear = True

    @staticmethod
    def parameterSpec():
        return (_f("radius", NM.toSi(1.), unit=NM, displayName="Sphere radius",
                   valueRange=(0., np.inf), generator=RandomUniform, decimals=9),
                _f("volFrac", Percent.toSi(10), unit=Percent,
                   displayName="Volume fraction of spheres",
                   valueRange=(Percent.toSi(0.001), Percent.toSi(100.)),
                   generator=RandomUniform, decimals=9),
                _p("mf", -1., displayName="standoff multiplier (-1 = auto)",
                   valueRange=(-1., 1.e6), unit=NoUnit, decimals=9),
                _p("sld", SLD.toSi(1e-6), unit=SLD,
                   displayName="Scattering length density difference",
                   valueRange=(0., np.inf), decimals=9))

    def __init__(self):
        super(LMADenseSphere, self).__init__()
        self.radius.setActive(True)

    def volume(self):
        return (np.pi * 4. / 3.) * self.radius() ** 3

    def absVolume(self):
        return self.volume() * self.sld() ** 2


class SphericalCoreShell(SASModel):
    """sphericalcoreshell.py:12-76"""
    modelId = 2
    shortName = "Core-Shell Sphere"

    @staticmethod
    def parameterSpec():
        return (_f("radius", NM.toSi(1.), unit=NM, displayName="Core Radius",
                   generator=RandomExponential, valueRange=(0., np.inf),
                   activeRange=NM.toSi((0.1, 1e3))),
                _f("t", NM.toSi(1.), unit=NM, displayName="Thickness of Shell",
                   generator=RandomExponential, valueRange=(0., np.inf),
                   activeRange=NM.toSi((0.1, 1e3))),
                _p("eta_c", SLD.toSi(3.16e-6), unit=SLD, displayName="Core SLD",
                   generator=RandomUniform, valueRange=(0, np.inf)),
                _p("eta_s", SLD.toSi(2.53e-6), unit=SLD, displayName="Shell SLD",
                   generator=RandomUniform, valueRange=(0, np.inf)),
                _p("eta_sol", 0., unit=SLD, displayName="Solvent SLD",
                   generator=RandomUniform, valueRange=(0, np.inf)))

    def __init__(self):
        super(SphericalCoreShell, self).__init__()
        self.radius.setActive(True)

    def volume(self):
        return 4. / 3 * np.pi * (self.radius() + self.t()) ** 3


class EllipsoidalCoreShell(SASModel):
    """ellipsoidalcoreshell.py:14-97"""
    modelId = 3
    shortName = "Core-Shell Ellipsoid"

    @staticmethod
    def parameterSpec():
        return (_f("a", NM.toSi(1.), unit=NM, displayName="Principal Core Radius",
                   generator=RandomExponential, valueRange=(0., np.inf),
                   activeRange=NM.toSi((0.1, 1e3))),
                _f("b", NM.toSi(10.), unit=NM, displayName="Equatorial Core Radius",
                   generator=RandomExponential, valueRange=(0., np.inf),
                   activeRange=NM.toSi((1.0, 1e4))),
                _f("t", NM.toSi(1.), unit=NM, displayName="Thickness of Shell",
                   generator=RandomExponential, valueRange=(0., np.inf),
                   activeRange=NM.toSi((0.1, 1e3))),
                _p("eta_c", SLD.toSi(3.15e-6), unit=SLD, displayName="Core SLD",
                   generator=RandomUniform, valueRange=(0, np.inf)),
                _p("eta_s", SLD.toSi(2.53e-6), unit=SLD, displayName="Shell SLD",
                   generator=RandomUniform, valueRange=(0, np.inf)),
                _p("eta_sol", 0., unit=SLD, displayName="Solvent SLD",
                   generator=RandomUniform, valueRange=(0, np.inf)),
                _p("intDiv", 100, displayName="Orientation Integration Divisions",
                   generator=RandomUniform, valueRange=(0, 1e4)))

    def __init__(self):
        super(EllipsoidalCoreShell, self).__init__()
        self.a.setActive(True)

    def volume(self):
        return 4. / 3 * np.pi * (self.a() + self.t()) * (self.b() + self.t()) ** 2


class EllipsoidsIsotropic(SASModel):
    """ellipsoidsisotropic.py:15-84"""
    modelId = 4
    shortName = "Isotropic Ellipsoids"

    @staticmethod
    def parameterSpec():
        return (_f("a", NM.toSi(1.), unit=NM, displayName="Radius of semi-axes a, b",
                   generator=RandomExponential, valueRange=NM.toSi((0.1, 1e10)),
                   activeRange=NM.toSi((0.1, 1e3))),
                _p("useAspect", True,
                   displayName="Use aspect ratio (checked) or length to define c-axis"),
                _f("c", NM.toSi(10.), unit=NM, displayName="Radius of semi-axes c",
                   generator=RandomExponential, valueRange=NM.toSi((0.1, 1e10)),
                   activeRange=NM.toSi((1., 1e4))),
                _f("aspect", 10.0, displayName="aspect ratio of semi-axes c to a, b",
                   generator=RandomExponential, valueRange=(1e-3, 1e3)),
                _p("intDiv", 100, displayName="Orientation Integration Divisions",
                   valueRange=(0, 1e4)),
                _p("sld", SLD.toSi(1e-6), unit=SLD,
                   displayName="Scattering length density difference",
                   valueRange=(0, SLD.toSi(1e-2))))

    def __init__(self):
        super(EllipsoidsIsotropic, self).__init__()
        self.a.setActive(True)

    def volume(self):
        Rc = self.a() * self.aspect() if self.useAspect() else self.c()
        return 4. / 3. * np.pi * self.a() ** 2. * Rc

    def absVolume(self):
        return self.volume() * self.sld() ** 2


class CylindersIsotropic(SASModel):
    """cylindersisotropic.py:16-101"""
    modelId = 5
    shortName = "SASfit Isotropic Cylinders"

    @staticmethod
    def parameterSpec():
        return (_f("radius", NM.toSi(1.), unit=NM, displayName="Cylinder Radius",
                   generator=RandomExponential, valueRange=(NM.toSi(0.1), np.inf)),
                _p("useAspect", True, displayName="Use aspect ratio (checked) or length "),
                _f("length", NM.toSi(10.), unit=NM, displayName="Length L of the Cylinder",
                   generator=RandomExponential, valueRange=(NM.toSi(0.1), NM.toSi(1e10))),
                _f("aspect", 10.0, displayName="Aspect ratio of the Cylinder",
                   generator=RandomExponential, valueRange=(1e-3, 1e3)),
                _p("intDiv", 100., displayName="Orientation Integration Divisions",
                   valueRange=(1, 1e4)),
                _p("sld", SLD.toSi(1e-6), unit=SLD,
                   displayName="Scattering length density difference",
                   valueRange=(0, np.inf)))

    def __init__(self):
        super(CylindersIsotropic, self).__init__()
        self.radius.setActive(True)

    def volume(self):
        hl = self.radius() * self.aspect() if self.useAspect() else self.length() / 2.
        return np.pi * self.radius() ** 2 * (hl * 2.)

    def absVolume(self):
        return self.volume() * self.sld() ** 2


class GaussianChain(SASModel):
    """gaussianchain.py:12-65"""
    modelId = 6
    shortName = "Gaussian Chain"

    @staticmethod
    def parameterSpec():
        return (_f("rg", NM.toSi(1.), unit=NM, displayName="radius of gyration, Rg",
                   generator=RandomExponential, valueRange=(0., np.inf),
                   activeRange=NM.toSi((1.0, 1e2))),
                _f("bp", NM.toSi(100.), unit=NM, displayName="scattering length of the polymer",
                   generator=RandomUniform, valueRange=(0., np.inf),
                   activeRange=NM.toSi((0.1, 1e3))),
                _f("etas", SLD.toSi(1e-6), unit=SLD,
                   displayName="scattering length density of the solvent",
                   generator=RandomUniform, valueRange=(0., np.inf),
                   activeRange=SLD.toSi((0.1, 10.))),
                _f("k", 1.0, displayName="volumetric scaling factor of Rg",
                   generator=RandomUniform, valueRange=(0., np.inf), activeRange=(0.1, 10.)))

    def __init__(self):
        super(GaussianChain, self).__init__()
        self.rg.setActive(True)

    def volume(self):
        return self.k() * self.rg() ** 2


class Kholodenko(SASModel):
    """kholodenko.py:16-94"""
    modelId = 7
    shortName = "Kholodenko Worm"

    @staticmethod
    def parameterSpec():
        return (_f("radius", NM.toSi(1.), unit=NM, displayName="Radius",
                   generator=RandomExponential, valueRange=(0., np.inf),
                   activeRange=NM.toSi((1, 5))),
                _f("lenKuhn", NM.toSi(1.), unit=NM, displayName="kuhn length",
                   generator=RandomUniform, valueRange=(0., np.inf),
                   activeRange=NM.toSi((10, 50))),
                _f("lenContour", NM.toSi(2.), unit=NM, displayName="contour length",
                   generator=RandomUniform, valueRange=(0., np.inf),
                   activeRange=NM.toSi((100, 1000))))

    def __init__(self):
        super(Kholodenko, self).__init__()
        self.radius.setActive(True)
        self.lenKuhn.setActive(True)
        self.lenContour.setActive(True)

    def volume(self):
        return np.pi * self.lenContour() * self.radius() ** 2


MODELS = dict((c.__name__, c) for c in (
    Sphere, LMADenseSphere, SphericalCoreShell, EllipsoidalCoreShell, EllipsoidsIsotropic,
    CylindersIsotropic, GaussianChain, Kholodenko))
