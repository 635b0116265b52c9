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


def batch_edges(n_raw, q, intensity, n_bin, x0_low=0., x0_high=np.inf, mask_zero=False, mask_neg=False):
    """Bin edges of every table of a batch exactly as DataObj._reBin computes them
    (dataobj.py:308-311: numpy.logspace(log10(min), log10(max + diff[-1] / 100), nBin + 1) over the
    masked rows), vectorised over the concatenated tables -- numpy's own log10 / power, so the bits
    are the reference's.  Returns [n_sets, n_bin + 1]; a table with fewer than two valid rows gets
    the degenerate edges of its only row (or NaN when it has none)."""
    n_raw = np.asarray(n_raw, dtype=np.int64)
    q = np.asarray(q, dtype=np.float64)
    m = valid_mask(q, intensity, x0_low, x0_high, mask_zero, mask_neg)
    start = np.concatenate([[0], np.cumsum(n_raw)[:-1]])
    nonempty = n_raw > 0
    idx = np.arange(q.size)
    seg = np.repeat(np.arange(n_raw.size), n_raw)

    def seg_reduce(ufunc, values, fill):
        out = np.full(n_raw.size, fill, dtype=values.dtype)
        if values.size:
            red = ufunc.reduceat(values, start[nonempty])
            out[nonempty] = red
        return out
    qmin = seg_reduce(np.minimum, np.where(m, q, np.inf), np.inf)
    qmax = seg_reduce(np.maximum, np.where(m, q, -np.inf), -np.inf)
    last = seg_reduce(np.maximum, np.where(m, idx, -1), -1)              # last valid row of each table
    m2 = m.copy()
    m2[last[last >= 0]] = False
    prev = seg_reduce(np.maximum, np.where(m2, idx, -1), -1)             # the valid row before it
    have2 = prev >= 0
    step = np.zeros(n_raw.size)
    step[have2] = (q[last[have2]] - q[prev[have2]]) / 100.               # numpy.diff(sanX)[-1] / 100.
    with np.errstate(invalid="ignore", divide="ignore"):
        edges = np.logspace(np.log10(qmin), np.log10(qmax + step), int(n_bin) + 1, axis=1)
    return np.ascontiguousarray(edges)


def prepare_batch(tables, engine, n_bin=100, fu_min=0.01, x0_low=0., x0_high=np.inf,
                  mask_zero=False, mask_neg=False):
    """Data preparation of MANY raw tables in one device call (mcsas_prepare_datasets):
    no per-dataset objects, bit-identical to `prepare` / the reference.  `tables`: sequence of
    (q [1/nm], I, sigma or None) -- all with or all without an uncertainty column.
    Returns a list of (q [1/m], I, sigma) array triples (views of one buffer), ready for
    Engine.upload_datasets."""
    n_raw = np.array([np.size(t[0]) for t in tables], dtype=np.int32)
    q = np.concatenate([np.asarray(t[0], dtype=np.float64).ravel() for t in tables]) * NM_TO_SI
    intensity = np.concatenate([np.asarray(t[1], dtype=np.float64).ravel() for t in tables])
    has_sigma = [t[2] is not None for t in tables]
    if any(has_sigma) and not all(has_sigma):
        raise ValueError("prepare_batch: either every table has an uncertainty column or none")
    sigma = np.concatenate([np.asarray(t[2], dtype=np.float64).ravel() for t in tables]) if all(has_sigma) else None
    n_bin = int(n_bin) if n_bin and n_bin > 0 else 0
    edges = batch_edges(n_raw, q, intensity, n_bin, x0_low, x0_high, mask_zero, mask_neg) if n_bin else None
    cap = n_bin if n_bin else int(n_raw.max())
    nq, qo, io, so = engine.prepare_datasets(n_raw, q, intensity, sigma, fu_min, x0_low, x0_high, mask_zero,
                                             mask_neg, n_bin, edges, cap)
    return [(qo[s, :nq[s]], io[s, :nq[s]], so[s, :nq[s]]) for s in range(n_raw.size)]
