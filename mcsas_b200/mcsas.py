# -*- coding: utf-8 -*-
"""McSAS -- drop-in host for the Monte-Carlo fitting path of the reference's McSAS class
(src/mcsas/mcsas/mcsas.py:27-615): same attributes (`data`, `model`, `result`, `stop`),
same settings as callables (src/mcsas/mcsas/mcsasparameters.json), same methods
(`calc`, `analyse`, `mcFit`, `histogram`) and the same `result[0]` dictionary
(mcsas.py:268-285).  The arithmetic runs in the CUDA chain kernel through the C-ABI.

Differences that are deliberate and documented in DESIGN.md:
  * `analyse` issues ONE kernel launch for all repetitions (the reference loops serially,
    mcsas.py:214); retries (mcsas.py:220-246) happen inside the kernel per repetition;
  * proposals come from in-kernel Philox (seeded by `seed`), not numpy's global stream;
    the deterministic parity mode uploads a recorded stream (`proposalStream`);
  * scale/background is the closed-form weighted LSQ that scipy's leastsq converges to.
"""
import ctypes
import logging
import time

import numpy as np

from . import _lib, smearing
from .data import hot_path_vectors
from .engine import Engine
from .models import ScatteringModel, Sphere
from .parameter import Parameter, isActiveFitParam


def _settings():
    """The 12 algorithm settings with the reference's defaults and ranges
    (mcsasparameters.json:1-104)."""
    return [
        Parameter("numContribs", 300, displayName="number of contributions", valueRange=(1, 1e6)),
        Parameter("numReps", 10, displayName="number of repetitions", valueRange=(1, 1e6)),
        Parameter("maxIterations", 100000, displayName="maximum iterations", valueRange=(1, 1e100)),
        Parameter("compensationExponent", 0.6666666, displayName="compensation exponent",
                  valueRange=(-np.inf, np.inf)),
        Parameter("convergenceCriterion", 1.0, displayName="convergence criterion",
                  valueRange=(0., np.inf)),
        Parameter("findBackground", True, displayName="find background level?"),
        Parameter("positiveBackground", False, displayName="background positive only?"),
        Parameter("startFromMinimum", False, displayName="start from minimum"),
        Parameter("maxRetries", 5, displayName="maximum retries", valueRange=(1, 100)),
        Parameter("autoClose", False, displayName="close after completion?"),
        Parameter("seriesStats", False, displayName="series statistics?"),
        Parameter("showIncomplete", False, displayName="show incomplete fits?"),
    ]


class McSAS(object):
    data = None
    model = None
    result = None

    @classmethod
    def factory(cls):
        """`McSAS.factory()()` as in the reference (mcsas.py:143-147)."""
        return cls

    def __init__(self, device=0, seed=None, precision=None, first_chain=0, specMax=0):
        self._params = _settings()
        for p in self._params:
            setattr(self, p.name(), p)
        self.device = device
        self.seed = seed if seed is not None else int(time.time_ns() & 0x7FFFFFFFFFFFFFFF)
        self.precision = precision      # None: chosen per data set (_resolvePrecision)
        self.firstChain = first_chain
        self.specMax = specMax
        self.proposalStream = None      # deterministic parity mode: [numReps][stream] in [0,1)
        self.result = []
        # McSAS.stop (mcsas.py:357, 238-243): a host int the kernel scheduler polls at every
        # time-slice boundary; it may be set from another thread / a callback while a fit runs
        self._stopFlag = ctypes.c_int(0)
        self.lastKernelMs = 0.0
        self.details = None
        self._engine = None

    def params(self):
        return self._params

    @property
    def stop(self):
        return bool(self._stopFlag.value)

    @stop.setter
    def stop(self, value):
        self._stopFlag.value = 1 if value else 0

    @property
    def engine(self):
        if self._engine is None:
            self._engine = Engine(self.device)
        return self._engine

    # ------------------------------------------------------------------ calc --
    def calc(self, **kwargs):
        """mcsas.py:149-179"""
        self.result = []
        self.stop = False
        assert self.data is not None
        if self.model is None:
            self.model = Sphere()
            logging.info("Default model not provided, setting to: {0}".format(self.model.name()))
        if not isinstance(self.model, ScatteringModel):
            raise TypeError("model must be one of the mcsas_b200.models plug-ins (user-defined "
                            "Python models cannot run in the CUDA kernel; no CPU fallback)")
        logging.info("Using model: {0}".format(self.model.name()))
        if not self.model.paramCount():
            logging.warning("No parameters to analyse given! Breaking up.")
            return
        self.analyse()
        if not len(self.result):
            return
        self.histogram()

    # --------------------------------------------------------------- analyse --
    def _config(self, numContribs, numReps, record_trace=False):
        return Engine.make_config(
            n_contribs=numContribs, n_reps=numReps,
            conv_crit=float(self.convergenceCriterion()), max_iter=int(self.maxIterations()),
            max_retries=int(self.maxRetries()), comp_exp=float(self.compensationExponent()),
            find_background=bool(self.findBackground()),
            positive_background=bool(self.positiveBackground()),
            start_from_minimum=bool(self.startFromMinimum()), precision=self._resolvePrecision(),
            seed=self.seed, spec_max=self.specMax, record_trace=record_trace,
            first_chain=self.firstChain)

    # fp32 form factors are ~3e-7 accurate: enough while the relative uncertainties are well
    # above that; data with sigma/I below this run the fp64 kernel
    FP32_MIN_REL_SIGMA = 1e-3

    def _resolvePrecision(self):
        if self.precision is not None:
            return self.precision
        _, I, sigma = hot_path_vectors(self.data)
        nz = np.abs(I) > 0
        if not nz.any():
            return _lib.FP32
        rel = np.median(np.abs(sigma[nz] / I[nz]))
        return _lib.FP32 if rel >= self.FP32_MIN_REL_SIGMA else _lib.FP64

    def _launch(self, numContribs, numReps, stream=None, record_trace=False):
        eng = self.engine
        q, I, sigma = hot_path_vectors(self.data)
        eng.upload_datasets([(q, I, sigma)])
        self.model.compensationExponent = float(self.compensationExponent())
        self.model.configureEngine(eng)
        self._configureSmearing(eng)
        cfg = self._config(numContribs, numReps, record_trace)
        eng.run_async(cfg, stream, stop_flag=self._stopFlag)
        self.lastKernelMs = eng.wait()
        return eng.fetch()

    def _configureSmearing(self, eng):
        """sasmodel.py:56-60: smear when the data carries a valid, enabled profile and the model
        can be smeared."""
        prof = smearing.profile_for(self.data, self.model)
        if prof is None:
            eng.set_smearing(None)
        else:
            eng.set_smearing(*prof)

    def analyse(self):
        """mcsas.py:191-285: numReps independent chains, retried up to maxRetries+1 attempts
        each; on failure the reference returns without results unless showIncomplete."""
        if not any(isActiveFitParam(p) for p in self.model.params()):
            raise ValueError("no active fit parameter: nothing to fit (mcsas.py:198-199)")
        numContribs, numReps = int(self.numContribs()), int(self.numReps())
        out = self._launch(numContribs, numReps, self.proposalStream)
        conval = out["conval"][0]
        failed = conval > float(self.convergenceCriterion())
        self.details = out
        if self.stop:                                         # mcsas.py:238-243
            logging.warning("Stop button pressed, exiting...")
            if not self.showIncomplete():
                return
        elif failed.any():
            logging.warning("Could not reach optimization criterion within {0} attempts "
                            "({1} of {2} repetitions), exiting..."
                            .format(int(self.maxRetries()) + 2, int(failed.sum()), numReps))   # (the reference's text)
            if not self.showIncomplete():
                return
        contributions = out["contribs"][0]                    # [N, P, R]
        contribMeasVal = out["fit_meas"][0][np.newaxis, :, :]  # [1, nq, R]
        scalings, backgrounds = out["scaling"][0], out["background"][0]
        numIter = out["num_iter"][0].astype(np.float64)
        times = out["elapsed"][0]
        for idx, param in enumerate(self.model.activeParams()):   # mcsas.py:434-437
            for nr in range(numReps):
                param.setActiveVal(contributions[:, idx, nr], index=nr)
        ddof = 1 if numReps > 1 else 0
        q = np.asarray(self.data.q)
        self.result.append(dict(
            contribs=contributions,
            fitMeasValMean=contribMeasVal.mean(axis=2),
            fitMeasValStd=contribMeasVal.std(axis=2),
            fitX0=q, dataX0=q,
            dataMean=np.asarray(self.data.f.binnedData),
            dataStd=np.asarray(self.data.f.binnedDataU),
            scaling=(scalings.mean(), scalings.std(ddof=ddof)),
            background=(backgrounds.mean(), backgrounds.std(ddof=ddof)),
            times=times,
            numIter=numIter.mean(),
            # extras (not in the reference's dict)
            numIterAll=out["num_iter"][0], numMoves=out["num_moves"][0],
            conval=conval, attempts=out["attempts"][0], evals=out["evals"][0],
            stepsAll=out["steps_all"][0],
            scalingAll=scalings, backgroundAll=backgrounds))

    # ----------------------------------------------------------------- mcFit --
    def mcFit(self, numContribs, outputMeasVal=False, outputDetails=False, nRun=None,
              proposalStream=None):
        """mcsas.py:287-439: one chain.  Returns [rset, (measVal), conval, (details)]."""
        saved = self.firstChain
        if nRun is not None:
            self.firstChain = saved + int(nRun)
        try:
            out = self._launch(int(numContribs), 1, proposalStream, record_trace=False)
        finally:
            self.firstChain = saved
        rset = out["contribs"][0][:, :, 0]
        result = [rset]
        if outputMeasVal:
            result.append(out["fit_meas"][0][:, 0])
        result.append(float(out["conval"][0, 0]))
        if outputDetails:
            result.append({"numIterations": int(out["num_iter"][0, 0]),
                           "numMoves": int(out["num_moves"][0, 0]),
                           "elapsed": float(out["elapsed"][0, 0]),
                           "scaling": float(out["scaling"][0, 0]),
                           "background": float(out["background"][0, 0])})
        if nRun is not None:
            for idx, param in enumerate(self.model.activeParams()):
                param.setActiveVal(rset[:, idx], index=nRun)
        return result

    # ------------------------------------------------------------- histogram --
    def histogram(self, contribs=None):
        """mcsas.py:445-615: fractions and observability on the GPU (mcsas_fractions),
        then each active parameter's histograms (Histogram.calc)."""
        if not isinstance(self.result, list) or not len(self.result):
            logging.info("There are no results to histogram, breaking up.")
            return
        if contribs is None:
            contribs = self.result[0]['contribs']
        if not all(np.array(contribs.shape, dtype=bool)):
            logging.info("Nothing to histogram, giving up.")
            return
        numContribs, dummy, numReps = contribs.shape
        eng = self.engine
        q, I, sigma = hot_path_vectors(self.data)
        eng.upload_datasets([(q, I, sigma)])
        self.model.configureEngine(eng)
        self._configureSmearing(eng)
        cfg = self._config(numContribs, numReps)
        cfg.precision = _lib.FP64
        f = eng.fractions(cfg, contribs[np.newaxis])
        fractions = dict(vol=(f["vol"][0], f["min_vol"][0]), num=(f["num"][0], f["min_num"][0]),
                         int=(f["volsqr"][0], f["min_volsqr"][0]),
                         surf=(f["surf"][0], f["min_surf"][0]))
        self.result[0].update(dict(
            volumeFraction=f["vol"][0], numberFraction=f["num"][0],
            totalVolumeFraction=f["totals"][0, 0], totalNumberFraction=f["totals"][0, 1],
            minimumRequiredVolume=f["min_vol"][0], minimumRequiredNumber=f["min_num"][0],
            scalingFactors=f["sc"][0]))
        for paramIndex, param in enumerate(self.model.activeParams()):
            param.histograms().calc(contribs, paramIndex, fractions)

    # ------------------------------------------------------------- multi-GPU --
    def gatherResults(self):
        """All ranks' contribution sets / scalings over NCCL (torch.distributed), SURVEY 8(e):
        the only collective of the path.  Returns the gathered dict on every rank."""
        from . import dist
        self.gathered = dist.gather_results(self.result[0] if self.result else None, self.device)
        return self.gathered
