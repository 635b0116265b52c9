# -*- coding: utf-8 -*-
"""Minimal data container with the attributes McSAS.mcFit reads from a SASData object
(SURVEY 8(a) row a10): data.q (sasdata.py:51-54), data.f.binnedData / binnedDataU
(datavector.py:96-117), data.x0.binnedData, data.f.limit / data.x0.limit (:147-148),
data.count (sasdata.py:74-75).  A reference SASData instance can be passed to
mcsas_b200.McSAS directly -- it is read through the same attributes."""
import numpy as np


class DataVector(object):
    def __init__(self, values, unc=None):
        self.binnedData = np.asarray(values, dtype=np.float64).ravel()
        self.binnedDataU = None if unc is None else np.asarray(unc, dtype=np.float64).ravel()

    @property
    def limit(self):
        return [self.binnedData.min(), self.binnedData.max()]


class SASData(object):
    """q in 1/m, intensity and uncertainty in (m sr)^-1 -- already clipped / rebinned /
    fuMin-floored (dataobj.py:204-345 is outside the accelerated path)."""

    def __init__(self, q, intensity, sigma, title="data"):
        q = np.asarray(q, dtype=np.float64).ravel()
        if not (q.size == np.size(intensity) == np.size(sigma)):
            raise ValueError("q, intensity and sigma must have the same length")
        self.title = title
        self.x0 = DataVector(q)
        self.f = DataVector(intensity, sigma)
        self.smearing = None        # a mcsas_b200.smearing.SmearingConfig (data.config.smearing)

    @property
    def q(self):
        return self.x0.binnedData

    @property
    def count(self):
        return self.x0.binnedData.size

    @classmethod
    def fromNanometers(cls, q_nm, intensity, sigma, **kw):
        """q given in 1/nm as in the reference's input files (sasdata.py:133-159)."""
        return cls(np.asarray(q_nm, dtype=np.float64) * 1e9, intensity, sigma, **kw)


def hot_path_vectors(data):
    """(q, I, sigma) float64 arrays exactly as mcFit / BackgroundScalingFit read them."""
    q = np.asarray(data.q, dtype=np.float64).ravel()
    I = np.asarray(data.f.binnedData, dtype=np.float64).ravel()
    u = data.f.binnedDataU
    sigma = np.ones_like(I) if u is None else np.asarray(u, dtype=np.float64).ravel()
    return q, I, sigma
