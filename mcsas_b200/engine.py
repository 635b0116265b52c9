# -*- coding: utf-8 -*-
"""Engine: numpy-facing wrapper of one C-ABI context (one GPU).

This is plumbing between the McSAS-compatible host classes (mcsas.py, models.py) and
libmcsas_b200.so; all arrays are float64 host buffers in SI units, as the reference
hands them to McSAS.mcFit."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import McSASB200Error, RunConfig, Tuning


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def _llp(a):
    return a.ctypes.data_as(C.POINTER(C.c_longlong)) if a is not None else None


class Engine(object):
    def __init__(self, device=0):
        self._L = _lib.lib()
        self._ctx = C.c_void_p()
        rc = self._L.mcsas_ctx_create(int(device), C.byref(self._ctx))
        if rc != 0:
            msg = self._L.mcsas_last_error(None)
            raise McSASB200Error("mcsas_ctx_create(%d) failed (%d): %s" % (
                device, rc, msg.decode() if msg else "?"))
        self.device = device
        self.nq = None
        self.n_sets = 0
        self.n_active = 0
        self._last = None

    def close(self):
        if self._ctx:
            self._L.mcsas_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._L.mcsas_last_error(self._ctx)
            raise McSASB200Error("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    def device_info(self):
        sm, khz = C.c_int(), C.c_int()
        name = C.create_string_buffer(128)
        self._check(self._L.mcsas_device_info(self._ctx, C.byref(sm), C.byref(khz), name, 128),
                    "device_info")
        return dict(sm_count=sm.value, clock_khz=khz.value, name=name.value.decode())

    def upload_datasets(self, datasets):
        """datasets: list of (q, I, sigma) float64 arrays (SI units)."""
        nq = np.array([len(d[0]) for d in datasets], dtype=np.int32)
        q = np.ascontiguousarray(np.concatenate([np.asarray(d[0], dtype=np.float64).ravel() for d in datasets]))
        I = np.ascontiguousarray(np.concatenate([np.asarray(d[1], dtype=np.float64).ravel() for d in datasets]))
        s = np.ascontiguousarray(np.concatenate([np.asarray(d[2], dtype=np.float64).ravel() for d in datasets]))
        if not (q.size == I.size == s.size):
            raise McSASB200Error("q, I, sigma lengths differ")
        self._check(self._L.mcsas_upload_datasets(self._ctx, len(datasets), _ip(nq), _dp(q), _dp(I), _dp(s)),
                    "upload_datasets")
        self.nq = nq.copy()
        self.n_sets = len(datasets)

    def set_model(self, model_id, params, active_idx, lo, hi, gen_kind, quad_rule=0, quad_nodes=0):
        params = np.ascontiguousarray(params, dtype=np.float64)
        active_idx = np.ascontiguousarray(active_idx, dtype=np.int32)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        gen_kind = np.ascontiguousarray(gen_kind, dtype=np.int32)
        self._check(self._L.mcsas_set_model(self._ctx, int(model_id), _dp(params), params.size,
                                            _ip(active_idx), active_idx.size, _dp(lo), _dp(hi),
                                            _ip(gen_kind), int(quad_rule), int(quad_nodes)),
                    "set_model")
        self.n_active = int(active_idx.size)

    def set_smearing(self, q_offset=None, weights=None, two_d_coll=False):
        """SmearingConfig.prepared of the data set (None: no smearing); canSmear models only."""
        if q_offset is None:
            self._check(self._L.mcsas_set_smearing(self._ctx, 0, None, None, 0), "set_smearing")
            return
        qo = np.ascontiguousarray(q_offset, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if qo.size != w.size:
            raise McSASB200Error("smearing profile: qOffset and weights differ in length")
        self._check(self._L.mcsas_set_smearing(self._ctx, int(qo.size), _dp(qo), _dp(w), int(bool(two_d_coll))),
                    "set_smearing")

    def eval(self, pset, comp_exp, set_index=0, precision=_lib.FP64):
        """ScatteringModel.calc on single contributions -> (it[M,nq], v, w, s)."""
        pset = np.ascontiguousarray(np.atleast_2d(pset), dtype=np.float64)
        M = pset.shape[0]
        nq = int(self.nq[set_index])
        it = np.zeros((M, nq))
        v, w, s = np.zeros(M), np.zeros(M), np.zeros(M)
        self._check(self._L.mcsas_eval(self._ctx, int(set_index), _dp(pset), M, float(comp_exp),
                                       int(precision), _dp(it), _dp(v), _dp(w), _dp(s)), "eval")
        return it, v, w, s

    @staticmethod
    def make_config(n_contribs=300, n_reps=10, conv_crit=1.0, max_iter=100000, max_retries=5,
                    comp_exp=0.6666666, find_background=True, positive_background=False,
                    start_from_minimum=False, precision=_lib.FP32, seed=0, spec_max=0,
                    record_trace=False, first_chain=0, cluster_size=0, slice_steps=0):
        return RunConfig(int(n_contribs), int(n_reps), float(conv_crit), int(max_iter),
                         int(max_retries), float(comp_exp), int(bool(find_background)),
                         int(bool(positive_background)), int(bool(start_from_minimum)),
                         int(precision), int(seed) & 0xFFFFFFFFFFFFFFFF, int(spec_max),
                         int(bool(record_trace)), int(first_chain), int(cluster_size),
                         int(slice_steps))

    def set_tuning(self, **kw):
        """Experiment knobs of the scheduler / kernel (struct mcsas_tuning); no arguments: defaults."""
        t = Tuning()
        for k, v in kw.items():
            if not hasattr(t, k):
                raise McSASB200Error("unknown tuning knob %r" % k)
            setattr(t, k, int(v))
        self._check(self._L.mcsas_set_tuning(self._ctx, C.byref(t)), "set_tuning")

    def run_async(self, cfg, stream=None, stop_flag=None):
        """stop_flag: a ctypes.c_int the caller owns (kept alive until wait()); non-zero stops the
        fit at its next time-slice boundary (McSAS.stop)."""
        stride = 0
        sp = None
        if stream is not None:
            stream = np.ascontiguousarray(stream, dtype=np.float64)
            n_chains = self.n_sets * cfg.n_reps
            stream = stream.reshape(n_chains, -1)
            stride = stream.shape[1]
            sp = _dp(stream)
            self._stream_keepalive = stream
        self._stop_keepalive = stop_flag
        fp = C.cast(C.pointer(stop_flag), C.POINTER(C.c_int)) if stop_flag is not None else None
        self._check(self._L.mcsas_run_async(self._ctx, C.byref(cfg), sp, C.c_longlong(stride), fp),
                    "run_async")
        self._last = cfg

    def wait(self):
        ms = C.c_float()
        self._check(self._L.mcsas_wait(self._ctx, C.byref(ms)), "wait")
        return ms.value

    def run_info(self):
        n, c, ms = C.c_int(), C.c_int(), C.c_float()
        self._check(self._L.mcsas_run_info(self._ctx, C.byref(n), C.byref(c), C.byref(ms)), "run_info")
        return dict(launches=n.value, last_cluster=c.value, kernel_ms=ms.value)

    def scheduler_info(self):
        s, u = C.c_int(), C.c_int()
        self._check(self._L.mcsas_scheduler_info(self._ctx, C.byref(s), C.byref(u)), "scheduler_info")
        return dict(backups_started=s.value, backups_used=u.value)

    def request_stop(self):
        self._L.mcsas_request_stop(self._ctx)

    def run(self, cfg, stream=None):
        self.run_async(cfg, stream)
        return self.wait()

    def fetch(self):
        cfg = self._last
        S, R, N, P = self.n_sets, cfg.n_reps, cfg.n_contribs, self.n_active
        C_ = S * R
        contribs = np.zeros((S, N, P, R))
        fit = np.zeros(int(self.nq.sum()) * R)
        scaling, background, conval, conval0, elapsed = (np.zeros(C_) for _ in range(5))
        num_iter = np.zeros(C_, dtype=np.int64)
        num_moves = np.zeros(C_, dtype=np.int64)
        evals = np.zeros(C_, dtype=np.int64)
        steps_all = np.zeros(C_, dtype=np.int64)
        attempts = np.zeros(C_, dtype=np.int32)
        self._check(self._L.mcsas_fetch(self._ctx, _dp(contribs), _dp(fit), _dp(scaling), _dp(background),
                                        _dp(conval), _dp(conval0), _llp(num_iter), _llp(num_moves),
                                        _ip(attempts), _dp(elapsed), _llp(evals), _llp(steps_all)), "fetch")
        fits, off = [], 0
        for s in range(S):
            n = int(self.nq[s]) * R
            fits.append(fit[off:off + n].reshape(int(self.nq[s]), R))
            off += n
        shp = (S, R)
        return dict(contribs=contribs, fit_meas=fits, scaling=scaling.reshape(shp),
                    background=background.reshape(shp), conval=conval.reshape(shp),
                    conval0=conval0.reshape(shp), num_iter=num_iter.reshape(shp),
                    num_moves=num_moves.reshape(shp), attempts=attempts.reshape(shp),
                    elapsed=elapsed.reshape(shp), evals=evals.reshape(shp),
                    steps_all=steps_all.reshape(shp))

    def fetch_trace(self, chain, capacity):
        acc = np.zeros(capacity, dtype=np.uint8)
        chi = np.zeros(capacity)
        self._check(self._L.mcsas_fetch_trace(self._ctx, int(chain),
                                              acc.ctypes.data_as(C.POINTER(C.c_ubyte)), _dp(chi),
                                              C.c_longlong(capacity)), "fetch_trace")
        return acc, chi

    def fractions(self, cfg, contribs):
        """McSAS.histogram per-repetition fractions of contribs[S,N,P,R]."""
        contribs = np.ascontiguousarray(contribs, dtype=np.float64)
        S, N, P, R = contribs.shape
        outs = [np.zeros((S, N, R)) for _ in range(8)]
        totals = np.zeros((S, 4, R))
        sc = np.zeros((S, 2, R))
        self._check(self._L.mcsas_fractions(self._ctx, C.byref(cfg), _dp(contribs),
                                            *[_dp(o) for o in outs], _dp(totals), _dp(sc)), "fractions")
        keys = ["vol", "num", "volsqr", "surf", "min_vol", "min_num", "min_volsqr", "min_surf"]
        d = dict(zip(keys, outs))
        d["totals"] = totals
        d["sc"] = sc
        return d

    def prepare_datasets(self, n_raw, q, intensity, sigma, fu_min, x0_low, x0_high, mask_zero, mask_neg,
                         n_bin, edges, cap):
        """mcsas_prepare_datasets: batched uncertainty floor / masks / log rebinning on the device.
        Returns (nq[S], q[S,cap], I[S,cap], sigma[S,cap])."""
        n_raw = np.ascontiguousarray(n_raw, dtype=np.int32)
        S = n_raw.size
        q = np.ascontiguousarray(q, dtype=np.float64)
        intensity = np.ascontiguousarray(intensity, dtype=np.float64)
        sigma = None if sigma is None else np.ascontiguousarray(sigma, dtype=np.float64)
        edges = None if edges is None else np.ascontiguousarray(edges, dtype=np.float64)
        nq = np.zeros(S, dtype=np.int32)
        qo, io, so = (np.zeros((S, int(cap))) for _ in range(3))
        self._check(self._L.mcsas_prepare_datasets(
            self._ctx, S, _ip(n_raw), _dp(q), _dp(intensity), _dp(sigma), float(fu_min), float(x0_low),
            float(x0_high), int(bool(mask_zero)), int(bool(mask_neg)), int(n_bin), _dp(edges), int(cap),
            _ip(nq), _dp(qo), _dp(io), _dp(so)), "prepare_datasets")
        return nq, qo, io, so

    def measure_peaks(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._L.mcsas_measure_peaks(self._ctx, C.byref(a), C.byref(b)), "measure_peaks")
        return dict(mufu_gops=a.value, ffma_tflops=b.value)
