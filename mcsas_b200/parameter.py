# -*- coding: utf-8 -*-
"""Host mirror of McSAS's parameter surface for the fitting path.

Mirrors (names, argument meaning, error behaviour) of
  Parameter / FitParameter factories    src/mcsas/bases/algorithm/parameter.py:597-651,
                                        src/mcsas/utils/parameter.py:735-743
  ParameterNumerical.setValue / clip    src/mcsas/bases/algorithm/parameter.py:405-414,489-495
  FitParameterBase (isActive, activeRange, activeValues, setActiveVal, histograms)
                                        src/mcsas/utils/parameter.py:578-698
  RandomUniform / RandomExponential*    src/mcsas/bases/algorithm/numbergenerator.py:28-31,168-189
  Histogram / Histograms / Moments / VectorResult   src/mcsas/utils/parameter.py:20-568
The reference builds one dynamic *type* per parameter; here a parameter is a plain
instance with the same accessor methods (p(), p.value(), p.setValue(), ...).
"""
import logging

import numpy as np


class ParameterError(Exception):
    pass


class ValueRangeError(ParameterError):
    pass


class DefaultValueError(ParameterError):
    pass


# ---------------------------------------------------------------- generators --
class NumberGenerator(object):
    """Number generators are *descriptors* of the proposal distribution here: the draws
    themselves happen in the kernel (Philox).  get() exists for API compatibility
    (ScatteringModel.generateParameters on the host)."""
    kind = 0

    @classmethod
    def get(cls, count=1):
        raise NotImplementedError


class RandomUniform(NumberGenerator):
    kind = 0

    @classmethod
    def get(cls, count=1):
        return np.random.uniform(size=count)


class RandomExponential(NumberGenerator):
    lower, upper = 0., 1.
    kind = 1

    @classmethod
    def get(cls, count=1):
        rs = 10 ** (np.random.uniform(cls.lower, cls.upper, count))
        return (rs - 1) / (10 ** (cls.upper - cls.lower))


class RandomExponential1(RandomExponential):
    pass


class RandomExponential2(RandomExponential):
    upper = 2.
    kind = 2


class RandomExponential3(RandomExponential):
    upper = 3.
    kind = 3


# --------------------------------------------------------------------- units --
class Unit(object):
    """Minimal stand-in for utils/units.py: SI <-> display magnitude."""

    def __init__(self, name="", factor=1.0):
        self.displayMagnitudeName, self.factor = name, factor

    def toSi(self, v):
        if isinstance(v, (tuple, list)):
            return type(v)(self.toSi(x) for x in v)
        return v * self.factor

    def toDisplay(self, v):
        if isinstance(v, (tuple, list)):
            return type(v)(self.toDisplay(x) for x in v)
        return v / self.factor


NoUnit = Unit("", 1.0)
NM = Unit("nm", 1e-9)            # units.py Length('nm')
SLD = Unit(u"Å⁻²", 1e20)         # units.py SLD('Å⁻²')
Percent = Unit("%", 1e-2)        # units.py Fraction('%')


# ---------------------------------------------------------------- parameters --
class ParameterBase(object):
    isFit = False

    def __init__(self, name, value, displayName=None, unit=None, **kw):
        self._name = name
        self._value = value
        self._displayName = displayName or name
        self._unit = unit or NoUnit

    def name(self):
        return self._name

    def displayName(self):
        return self._displayName

    def unit(self):
        return self._unit

    def value(self):
        return self._value

    def __call__(self):
        return self.value()

    def setValue(self, newValue, clip=True):
        if newValue is None:
            return
        self._value = newValue

    def toSi(self, v):
        return self._unit.toSi(v)

    def toDisplay(self, v):
        return self._unit.toDisplay(v)

    def displayMagnitudeName(self):
        return self._unit.displayMagnitudeName

    def isActive(self):
        return False

    def __str__(self):
        return u"{0}: {1}".format(self._name, self._value)


class ParameterBoolean(ParameterBase):
    def setValue(self, newValue, clip=True):
        self._value = bool(newValue)


class ParameterNumerical(ParameterBase):
    def __init__(self, name, value, valueRange=None, generator=None, decimals=None, **kw):
        super(ParameterNumerical, self).__init__(name, value, **kw)
        if valueRange is None:
            valueRange = (-np.inf, np.inf)
        self.setValueRange(valueRange)
        self._generator = generator or RandomUniform
        self._decimals = decimals

    def setValueRange(self, newRange):
        if newRange is None or len(newRange) != 2:
            raise ValueRangeError("A value range is mandatory for a numerical parameter!")
        self._valueRange = (min(newRange), max(newRange))

    def valueRange(self):
        return self._valueRange

    def min(self):
        return self._valueRange[0]

    def max(self):
        return self._valueRange[1]

    def generator(self):
        return self._generator

    def setGenerator(self, gen):
        if not (isinstance(gen, type) and issubclass(gen, NumberGenerator)):
            raise ParameterError("NumberGenerator type expected!")
        self._generator = gen

    def clip(self, value=None):
        if value is None:
            value = self.value()
        if value is None:
            return None
        if isinstance(value, (tuple, list)):
            return type(value)(self.clip(v) for v in value)
        return min(max(value, self.min()), self.max())

    def setValue(self, newValue, clip=True):
        if newValue is None:
            return
        if not isinstance(newValue, (int, float, np.integer, np.floating)):
            raise DefaultValueError(u"A value has to be numerical! ({})".format(newValue))
        if clip:
            newValue = self.clip(newValue)
        self._value = newValue

    def generate(self, lower=None, upper=None, count=1):
        """generateValues (parameter.py:66-84): values * (hi - lo) + lo."""
        lo, hi = self.valueRange()
        if lower is not None:
            lo = max(lo, lower)
        if upper is not None:
            hi = min(hi, upper)
        return np.asarray(self.generator().get(count) * (hi - lo) + lo, dtype=float)


class ParameterFloat(ParameterNumerical):
    pass


class FitParameterMixin(object):
    """FitParameterBase, utils/parameter.py:578-698."""
    isFit = True

    def _initFit(self, activeRange=None):
        self._isActive = False
        self._activeRange = None
        self._histograms = Histograms()
        self._activeValues = []
        if activeRange is not None:
            self.setActiveRange(activeRange)

    def isActive(self):
        return self._isActive

    def setIsActive(self, isActive):
        self._isActive = bool(isActive)

    setActive = setIsActive

    def setActiveRange(self, newRange):
        if len(newRange) != 2:
            raise ValueRangeError("Active ranges have to consist of two values!")
        newRange = self.clip(tuple(newRange))
        self._activeRange = (min(newRange), max(newRange))
        self._histograms.updateRanges()

    def activeRange(self):
        if self._activeRange is None:
            return self.valueRange()
        return self._activeRange

    def displayActiveRange(self):
        lo, hi = self.activeRange()
        return (self.toDisplay(lo), self.toDisplay(hi))

    def setDisplayActiveRange(self, newRange):
        self.setActiveRange((self.toSi(min(newRange)), self.toSi(max(newRange))))

    def histograms(self):
        return self._histograms

    def setHistograms(self, h):
        self._histograms = h

    def activeValues(self):
        return self._activeValues

    def setActiveValues(self, v):
        self._activeValues = v

    def activeVal(self, val=None, index=None):
        if index is None:
            return self.activeValues()
        return self.activeValues()[index % len(self.activeValues())]

    def setActiveVal(self, val, index=None):
        if not self.isActive():
            logging.error("Value of parameter cannot be set, parameter not active!")
            return
        tempVal = self.activeValues()
        if index is None:
            index = len(tempVal)
        elif index < 0:
            index = len(tempVal) + index + 1
        while len(tempVal) <= index:
            tempVal.append(None)
        tempVal[index] = val
        self.setActiveValues(tempVal)

    def generate(self, lower=None, upper=None, count=1):
        if lower is None:
            lower = min(self.activeRange())
        if upper is None:
            upper = max(self.activeRange())
        return ParameterNumerical.generate(self, lower=lower, upper=upper, count=count)


class FitParameterFloat(FitParameterMixin, ParameterFloat):
    def __init__(self, name, value, activeRange=None, **kw):
        ParameterFloat.__init__(self, name, value, **kw)
        self._initFit(activeRange)


def Parameter(name, value, **kw):
    """Parameter factory (bases/algorithm/parameter.py:597-651): type chosen by value."""
    kw.pop("paramTypes", None)
    if isinstance(value, bool):
        return ParameterBoolean(name, value, **kw)
    if isinstance(value, float):
        return ParameterFloat(name, value, **kw)
    if isinstance(value, (int, np.integer)):
        return ParameterNumerical(name, value, **kw)
    return ParameterBase(name, value, **kw)


def FitParameter(name, value, **kw):
    """FitParameter factory (utils/parameter.py:735-743)."""
    if isinstance(value, bool) or not isinstance(value, (int, float)):
        raise ParameterError("FitParameter '%s': numerical value expected" % name)
    return FitParameterFloat(name, float(value), **kw)


def isFitParam(param):
    return getattr(param, "isFit", False)


def isActiveFitParam(param):
    return isFitParam(param) and param.isActive()


# ---------------------------------------------------- histograms and moments --
class VectorResult(object):
    """utils/parameter.py:156-184"""

    def __init__(self, vecResult):
        assert vecResult.ndim == 2
        self._full = vecResult
        self._mean = vecResult.mean(axis=1)
        ddof = 1 if len(vecResult) > 1 else 0
        self._std = vecResult.std(axis=1, ddof=ddof)

    mean = property(lambda self: self._mean)
    std = property(lambda self: self._std)
    full = property(lambda self: self._full)


class Moments(object):
    """Distribution moments per repetition and their spread, utils/parameter.py:20-122."""

    @staticmethod
    def fieldNames():
        return ("totalValue", "totalValueStd", "mean", "meanStd", "variance", "varianceStd",
                "skew", "skewStd", "kurtosis", "kurtosisStd")

    @property
    def fields(self):
        return self._total + self._mean + self._variance + self._skew + self._kurtosis

    total = property(lambda self: self._total)
    mean = property(lambda self: self._mean)
    variance = property(lambda self: self._variance)
    skew = property(lambda self: self._skew)
    kurtosis = property(lambda self: self._kurtosis)

    def __str__(self):
        return str(self.fields)

    def __init__(self, contribs, paramIndex, valueRange, fraction):
        par = contribs[:, paramIndex, :]
        numContribs, numReps = par.shape
        lo, hi = min(valueRange), max(valueRange)
        valid = (par > lo) & (par < hi)                  # :70-74 (open interval)
        val, mu, var, skw, krt = (np.zeros(numReps) for _ in range(5))
        for ri in range(numReps):
            m = valid[:, ri]
            if not m.any():
                continue
            rset, frac = par[m, ri], fraction[m, ri]
            sf = frac.sum()
            val[ri] = sf
            mu[ri] = (rset * frac).sum()
            if 0 != sf:
                mu[ri] /= sf
            var[ri] = ((rset - mu[ri]) ** 2 * frac).sum() / sf
            sigma = np.sqrt(abs(var[ri]))
            if (sf * sigma) == 0.0:
                continue
            skw[ri] = ((rset - mu[ri]) ** 3 * frac).sum() / (sf * sigma ** 3)
            krt[ri] = ((rset - mu[ri]) ** 4 * frac).sum() / (sf * sigma ** 4)
        ddof = 1 if numReps > 1 else 0
        self._total = (val.mean(), val.std(ddof=ddof))
        self._mean = (mu.mean(), mu.std(ddof=ddof))
        self._variance = (var.mean(), var.std(ddof=ddof))
        self._skew = (skw.mean(), skw.std(ddof=ddof))
        self._kurtosis = (krt.mean(), krt.std(ddof=ddof))
        self.perRepetition = dict(total=val, mean=mu, variance=var, skew=skw, kurtosis=krt)


class Histogram(object):
    """Histogram settings + results of one fit parameter, utils/parameter.py:187-541."""

    @staticmethod
    def xscaling(index=None):
        avail = ('lin', 'log')
        try:
            return avail[index]
        except Exception:
            return avail

    @staticmethod
    def yweighting(index=None):
        avail = ('vol', 'num', 'int', 'surf')
        try:
            return avail[index]
        except Exception:
            return avail

    def __init__(self, param, lower, upper, binCount=50, xscale=None, yweight=None,
                 autoFollow=True):
        self.param = param
        if not (isinstance(binCount, (int, np.integer)) and binCount > 0):
            raise ValueError("Histogram bin count has to be an integer larger 0!")
        self.binCount = int(binCount)
        self.xrange = (float(lower), float(upper))
        self.xscale = xscale
        self.yweight = yweight
        if not isinstance(autoFollow, bool):
            autoFollow = (str(autoFollow).title() == "True")
        self.autoFollow = autoFollow
        self._xLowerEdge = self._xMean = self._xWidth = None
        self._bins = self._cdf = self._observability = self._moments = None

    @property
    def xscale(self):
        return self._xscale

    @xscale.setter
    def xscale(self, kind):
        self._xscale = str(kind).strip()
        if self._xscale not in self.xscaling():
            self._xscale = self.xscaling(1)

    @property
    def yweight(self):
        return self._yweight

    @yweight.setter
    def yweight(self, kind):
        self._yweight = str(kind).strip()
        if self._yweight not in self.yweighting():
            self._yweight = self.yweighting(0)

    @property
    def xrange(self):
        return self._xrange

    @xrange.setter
    def xrange(self, valueRange):
        lo, hi = min(valueRange), max(valueRange)
        lo, hi = max(self.param.min(), lo), min(self.param.max(), hi)
        self._xrange = lo, hi

    lower = property(lambda self: self._xrange[0])
    upper = property(lambda self: self._xrange[1])
    xLowerEdge = property(lambda self: self._xLowerEdge)
    xMean = property(lambda self: self._xMean)
    xWidth = property(lambda self: self._xWidth)
    bins = property(lambda self: self._bins)
    cdf = property(lambda self: self._cdf)
    observability = property(lambda self: self._observability)
    moments = property(lambda self: self._moments)

    def updateRange(self):
        if self.autoFollow:
            self.xrange = self.param.activeRange()
        self.xrange = self.xrange

    def _setXLowerEdge(self):
        if 'lin' in self.xscale:
            self._xLowerEdge = np.linspace(self.lower, self.upper, self.binCount + 1)
        else:
            self._xLowerEdge = np.logspace(np.log10(self.lower), np.log10(self.upper),
                                           self.binCount + 1)
        self._xWidth = np.diff(self._xLowerEdge)
        self._xMean = 0.5 * (self._xLowerEdge[:-1] + self._xLowerEdge[1:])

    def calc(self, contribs, paramIndex, fractions):
        """:420-469.  Bin sums, observability (mean minimum-required fraction per bin),
        normalised CDF per repetition; then statistics over repetitions and Moments."""
        self._setXLowerEdge()
        numContribs, dummy, numReps = contribs.shape
        fraction, minReq = fractions[self.yweight]
        edges = self._xLowerEdge
        bins = np.zeros((self.binCount, numReps))
        obs = np.zeros((self.binCount, numReps))
        cdf = np.zeros((self.binCount, numReps))
        for ri in range(numReps):
            par = contribs[:, paramIndex, ri]
            for bi in range(self.binCount):
                mask = (par >= edges[bi]) & (par < edges[bi + 1])
                v = fraction[mask, ri].sum()
                bins[bi, ri] = 0. if np.isnan(v) else v
                obs[bi, ri] = minReq[mask, ri].mean() if mask.any() else 0.
            c = np.cumsum(bins[:, ri])
            cdf[:, ri] = c / c.max() if c.max() != 0.0 else 0.
        self._bins = VectorResult(bins)
        self._cdf = VectorResult(cdf)
        self._observability = np.zeros(self.binCount)
        for bi in range(self.binCount):
            o = obs[bi, :]
            o = o[o < np.inf]
            if o.size:
                self._observability[bi] = o.max()
        self._moments = Moments(contribs, paramIndex, self.xrange, fraction)

    def __str__(self):
        return "-".join(["hist", self.param.name(), str(self.autoFollow),
                         "%g" % self.lower, "%g" % self.upper, str(self.binCount),
                         self.xscale, self.yweight])

    __repr__ = __str__

    def __hash__(self):
        value = hash(id(self.param))
        for attr in "lower", "upper", "binCount", "xscale", "yweight":
            value ^= hash(getattr(self, attr))
        return value

    def __eq__(self, other):
        return hash(self) == hash(other)


class Histograms(list):
    """utils/parameter.py:543-568"""

    def append(self, value):
        if value in self:
            return
        super(Histograms, self).append(value)

    def clear(self):
        self[:] = []

    def updateRanges(self):
        for h in self:
            h.updateRange()

    def calc(self, *args):
        for h in self:
            h.calc(*args)
