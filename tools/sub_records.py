#!/usr/bin/env python
"""Experiment (not product): bench.py's sub-records alone (configs[2] cylinder with both rules, configs[3],
a slice of configs[4]) -- the same bounded runs, without the headline fit."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcsas_b200 import Engine  # noqa: E402

eng = Engine(0)
peaks = eng.measure_peaks()
sm = eng.device_info()["sm_count"]
eng.close()
for k, v in bench.sub_records(peaks, sm).items():
    print(k, json.dumps(v), flush=True)
