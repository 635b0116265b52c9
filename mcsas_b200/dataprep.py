# -*- coding: utf-8 -*-
"""Data preparation that feeds the hot path: what DataObj/SASData do
between a raw (q, I, sigma) table and the vectors McSAS.mcFit reads.

Mirrors, in numpy, for 1-D data:
  SASData.__init__            dataobj/sasdata.py:133-159   q [1/nm] -> [1/m], I in (m sr)^-1
  DataObj._prepareUncertainty dataobj/dataobj.py:204-226   sigma <- max(sigma, fuMin * I)
  DataObj._applyFMasks/_applyLimits  :253-269              zero / negative masks, q window
  DataObj._reBin              :288-345                     log-spaced rebinning to <= nBin bins
with the DataConfig defaults (dataobj/dataconfig.py:77-103): nBin 100, fuMin 1 %.
This is O(nq) host work that runs once per dataset; it is here so that batches of datasets
(BASELINE config 5) can be prepared without the reference's object machinery."""
import numpy as np

from .data import SASData

NM_TO_SI = 1e9      # ScatteringVector("nm^-1").toSi


def prepare_uncertainty(intensity, sigma, fu_min=0.01):
    """dataobj.py:204-226."""
    intensity = np.asarray(intensity, dtype=np.float64)
    floor = fu_min * intensity
    if sigma is None:
        out = floor.copy()
    else:
        out = np.maximum(np.asarray(sigma, dtype=np.float64), floor)
    out[~np.isfinite(out)] = np.inf
    return out


def valid_mask(q, intensity, x0_low=0., x0_high=np.inf, mask_zero=False, mask_neg=False):
    """dataobj.py:239-269 (_initMask, _applyFMasks, _applyLimits)."""
    q = np.asarray(q, dtype=np.float64)
    intensity = np.asarray(intensity, dtype=np.float64)
    m = np.isfinite(intensity)
    if mask_zero:
        m &= intensity != 0.0
    if mask_neg:
        m &= intensity > 0.0
    m &= q >= x0_low
    m &= q <= x0_high
    return m


def rebin(q, intensity, sigma, n_bin=100):
    """dataobj.py:288-345: log-spaced bins over [min q, max q]; a bin with one point keeps
    it, a bin with several takes the means and max(SEM, propagated uncertainty); empty bins
    are dropped.  n_bin <= 0: no rebinning."""
    q = np.asarray(q, dtype=np.float64)
    intensity = np.asarray(intensity, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    if not (n_bin > 0) or q.size == 0:
        return q.copy(), intensity.copy(), sigma.copy()
    n_bin = int(n_bin)
    last = np.diff(q)[-1] / 100. if q.size > 1 else 0.
    edges = np.logspace(np.log10(q.min()), np.log10(q.max() + last), n_bin + 1)
    # (q >= edges[i]) & (q < edges[i+1])
    idx = np.searchsorted(edges, q, side="right") - 1
    qb, ib, sb = [], [], []
    for b in range(n_bin):
        sel = idx == b
        n = int(sel.sum())
        if n == 0:
            continue
        if n == 1:
            qb.append(q[sel][0]); ib.append(intensity[sel][0]); sb.append(sigma[sel][0])
            continue
        f, fu = intensity[sel], sigma[sel]
        qb.append(q[sel].mean()); ib.append(f.mean())
        sb.append(max(f.std(ddof=1) / np.sqrt(1. * n), np.sqrt((fu ** 2).sum() / n)))
    qb, ib, sb = np.array(qb), np.array(ib), np.array(sb)
    keep = ~np.isnan(ib)
    return qb[keep], ib[keep], sb[keep]


def prepare(q_nm, intensity, sigma=None, n_bin=100, fu_min=0.01, x0_low=0., x0_high=np.inf,
            mask_zero=False, mask_neg=False, title="data"):
    """Raw table (q in 1/nm as in the reference's input files) -> SASData for McSAS.
    x0_low / x0_high are in 1/m like DataConfig.x0Low/x0High."""
    q = np.asarray(q_nm, dtype=np.float64).ravel() * NM_TO_SI
    intensity = np.asarray(intensity, dtype=np.float64).ravel()
    sig = prepare_uncertainty(intensity, None if sigma is None else np.asarray(sigma).ravel(), fu_min)
    m = valid_mask(q, intensity, x0_low, x0_high, mask_zero, mask_neg)
    qb, ib, sb = rebin(q[m], intensity[m], sig[m], n_bin)
    return SASData(qb, ib, sb, title=title)
