# -*- coding: utf-8 -*-
"""Smearing profiles: the integration points and weights the reference's
SmearingConfig classes prepare for the smeared branch of SASModel.calcIntensity
(sasmodel.py:56-73).  Mirrors dataobj/sasconfig.py:

  TrapezoidSmearing.setIntPoints / halfTrapzPDF   :110-149   slit / 2-D averaged trapezoidal beam
  GaussianSmearing.setIntPoints                   :209-233   Gaussian beam profile
  SmearingConfig.prepared / doSmear / twoDColl / nSteps      :17-75

A profile is attached to the data (`SASData.smearing = TrapezoidSmearing(...)`); McSAS hands
`prepared` to the kernels through mcsas_set_smearing for the canSmear models (Sphere,
LMADenseSphere).  Quantities are in the units of q (1/m), like DataConfig.x0Low."""
import numpy as np


class SmearingConfig(object):
    def __init__(self, doSmear=True, nSteps=25, twoDColl=False):
        self._doSmear, self._nSteps, self._twoDColl = bool(doSmear), int(nSteps), bool(twoDColl)
        self._qOffset = self._weights = None

    def doSmear(self):
        return self._doSmear

    def nSteps(self):
        return self._nSteps

    def twoDColl(self):
        return self._twoDColl

    @property
    def qOffset(self):
        return self._qOffset

    @property
    def weights(self):
        return self._weights

    @property
    def prepared(self):
        return self._qOffset, self._weights

    def inputValid(self):
        raise NotImplementedError

    def _offsets(self, q, low_div, upper):
        """The logarithmic offset grid shared by both profiles (sasconfig.py:131-141, 218-228)."""
        q = np.asarray(q, dtype=np.float64)
        n = self._nSteps
        if self._twoDColl:
            qo = np.logspace(np.log10(q.min() / low_div), np.log10(upper), num=int(np.ceil(n / 2.)))
            return np.concatenate((-qo[::-1], [0, ], qo))
        qo = np.logspace(np.log10(q.min() / low_div), np.log10(upper), num=n)
        return np.concatenate(([0, ], qo))

    def active(self):
        return self._doSmear and self.inputValid()

    def locs(self, q):
        """SASConfig.prepareSmearing (sasconfig.py:308-339): [nq, len(qOffset)]."""
        q = np.asarray(q, dtype=np.float64)
        if self._twoDColl:
            return np.add.outer(q, self._qOffset)
        return np.sqrt(np.add.outer(q ** 2, self._qOffset ** 2))


class TrapezoidSmearing(SmearingConfig):
    def __init__(self, umbra=0., penumbra=0., **kw):
        super(TrapezoidSmearing, self).__init__(**kw)
        self._umbra, self._penumbra = float(umbra), float(penumbra)

    def umbra(self):
        return self._umbra

    def penumbra(self):
        return self._penumbra

    def inputValid(self):                                    # sasconfig.py:95-97
        return self._umbra > 0. and self._penumbra > self._umbra

    @staticmethod
    def halfTrapzPDF(x, c, d):                               # sasconfig.py:110-126
        x = np.abs(np.asarray(x, dtype=np.float64))
        pdf = x * 0.
        pdf[x < c] = 1.
        if d > c:
            sel = (c <= x) & (x < d)
            pdf[sel] = (1. / (d - c)) * (d - x[sel])
        norm = 1. / (d + c)
        return pdf * norm, norm

    def setIntPoints(self, q):                               # sasconfig.py:128-149
        self._qOffset = self._offsets(q, 5., self._penumbra / 2.)
        self._weights, _ = self.halfTrapzPDF(self._qOffset, self._umbra, self._penumbra)
        return self.prepared


class GaussianSmearing(SmearingConfig):
    def __init__(self, variance=0., **kw):
        super(GaussianSmearing, self).__init__(**kw)
        self._variance = float(variance)

    def variance(self):
        return self._variance

    def inputValid(self):                                    # sasconfig.py:201-203
        return self._variance > 0.

    def setIntPoints(self, q):                               # sasconfig.py:209-233 (scipy.stats.norm.pdf)
        self._qOffset = self._offsets(q, 3., 2.5 * self._variance)
        s = self._variance
        self._weights = np.exp(-0.5 * (self._qOffset / s) ** 2) / (s * np.sqrt(2. * np.pi))
        return self.prepared


def profile_for(data, model):
    """(qOffset, weights, twoDColl) McSAS must hand to the kernels for this data/model pair, or
    None -- the conditions of sasmodel.py:56-60."""
    sm = getattr(data, "smearing", None)
    if sm is None:
        cfg = getattr(data, "config", None)               # a reference SASData object
        sm = getattr(cfg, "smearing", None)
    if sm is None or not getattr(model, "canSmear", False):
        return None
    if not (sm.doSmear() and sm.inputValid()):
        return None
    qo, w = sm.prepared
    if qo is None:
        qo, w = sm.setIntPoints(np.asarray(data.q))
    return np.asarray(qo, dtype=np.float64), np.asarray(w, dtype=np.float64), bool(sm.twoDColl())
