# -*- coding: utf-8 -*-
"""Synthetic SAS curves of the shapes BASELINE.json names (SURVEY 8(d) "Concrete inputs").
Data synthesis only -- not part of the fitting path."""
import numpy as np

NM = 1e-9


def _sphere_intensity(q_nm, radii_nm):
    qr = np.outer(q_nm, radii_nm)
    F = 3. * (np.sin(qr) - qr * np.cos(qr)) / qr ** 3
    V = 4. / 3. * np.pi * radii_nm ** 3
    return ((V * F) ** 2).mean(axis=1)


def sphere_curve(nq=1000, mean_nm=10., sd_nm=2., seed=0, n_radii=2000, qmin=0.03, qmax=3.0):
    """Config 2: Gaussian sphere-size distribution, I(q0)=1e6, +1.0 background, 1 % noise.
    Returns q [1/m], I, sigma (what SASData hands to McSAS.mcFit with nBin=0)."""
    rng = np.random.default_rng(seed)
    q_nm = np.logspace(np.log10(qmin), np.log10(qmax), nq)
    radii = np.clip(rng.normal(mean_nm, sd_nm, n_radii), 0.5, None)
    I = _sphere_intensity(q_nm, radii)
    I = I / I[0] * 1e6 + 1.0
    sigma = 0.01 * I
    I = I + rng.normal(0., 1., nq) * sigma
    return q_nm * 1e9, I, sigma


def sphere_batch(n_sets, nq=1000, first=0):
    """Config 5: datasets with (mu, sd) ~ U(5,50) nm, U(0.05,0.3)*mu; dataset seed = index."""
    out = []
    for i in range(first, first + n_sets):
        rng = np.random.default_rng(100000 + i)
        mu = rng.uniform(5., 50.)
        sd = rng.uniform(0.05, 0.3) * mu
        out.append(sphere_curve(nq=nq, mean_nm=mu, sd_nm=sd, seed=i, n_radii=500))
    return out
