# -*- coding: utf-8 -*-
"""Result files in the reference's formats: what gui/calc.py's
Calculator writes after McSAS.calc() -- so scripts that consume McSAS output keep working.

  *_fit.dat            gui/calc.py:398-402   fitX0 dataMean dataStd fitMeasValMean fitMeasValStd
  *_hist-<...>.dat     :404-417              xMean xWidth yMean yStd Obs cdfMean cdfStd
  *_stats_<param>.dat  :381-396              lower upper weighting + Moments.fieldNames()
  *_contributions.pickle  :419-426           pickle of result['contribs']  [N, P, R]
  *_settings.cfg       :428-462              configparser: I/O, MCSAS and model settings
Text layout = AsciiFile (datafile/asciifile.py:15-68): one header line of column names joined
by a blank, then rows of "{0: 14.6E}" values joined by a blank (non-numeric cells verbatim),
newline separated, no trailing newline."""
import configparser
import pickle

import numpy as np

from .parameter import Moments, isActiveFitParam

VALUE_FORMAT = "{0: 14.6E}"
SEPARATOR = " "
NEWLINE = "\n"

FIT_COLUMNS = ("fitX0", "dataMean", "dataStd", "fitMeasValMean", "fitMeasValStd")
HIST_COLUMNS = ("xMean", "xWidth", "yMean", "yStd", "Obs", "cdfMean", "cdfStd")


def format_value(value):
    """AsciiFile.formatValue (asciifile.py:24-29)."""
    try:
        return VALUE_FORMAT.format(value)
    except (ValueError, TypeError):
        return "{0}".format(value)


def format_table(column_names, columns):
    """Calculator.writeResults (gui/calc.py:128-155): header + vstack(columns).T rows."""
    data = np.vstack([np.ravel(c) for c in columns]).T
    lines = [SEPARATOR.join(column_names)]
    body = NEWLINE.join(SEPARATOR.join(format_value(v) for v in row) for row in data)
    return lines[0] + NEWLINE + body


def write_table(filename, column_names, columns):
    with open(filename, "w") as fd:
        fd.write(format_table(column_names, columns))
    return filename


def write_fit(filename, result):
    """*_fit.dat from McSAS.result[0]."""
    return write_table(filename, FIT_COLUMNS, [result[c] for c in FIT_COLUMNS])


def write_distribution(filename, hist):
    """*_hist-*.dat from one Histogram after McSAS.histogram()."""
    cols = [hist.xMean, hist.xWidth, hist.bins.mean, hist.bins.std, hist.observability,
            hist.cdf.mean, hist.cdf.std]
    return write_table(filename, HIST_COLUMNS, cols)


def write_statistics(filename, param):
    """*_stats_<param>.dat: one row per histogram range of the parameter."""
    names = ("lower", "upper", "weighting") + tuple(Moments.fieldNames())
    stats = dict((n, []) for n in names)
    for h in param.histograms():
        values = tuple(h.xrange) + (h.yweight,) + tuple(h.moments.fields)
        for n, v in zip(names, values):
            stats[n].append(format_value(v))
    return write_table(filename, names, [stats[n] for n in names])


def write_contributions(filename, result):
    with open(filename, "wb") as fh:
        pickle.dump(result["contribs"], fh)
    return filename


def write_settings(filename, algo, data_filename="", output_basename="", x0_limits=None):
    """*_settings.cfg (gui/calc.py:428-462)."""
    config = configparser.RawConfigParser()
    config.add_section("I/O Settings")
    config.set("I/O Settings", "fileName", data_filename)
    config.set("I/O Settings", "outputBaseName", output_basename)
    sec = "MCSAS Settings"
    config.add_section(sec)
    for p in algo.params():
        config.set(sec, p.name(), str(p.value()))
    config.set(sec, "model", algo.model.name())
    if x0_limits is not None:
        config.set(sec, "X0 limits", str(list(x0_limits)))
    sec = "Model Settings"
    config.add_section(sec)
    for p in algo.model.params():
        if isActiveFitParam(p):
            config.set(sec, p.name() + "_min", str(p.min()))
            config.set(sec, p.name() + "_max", str(p.max()))
        else:
            config.set(sec, p.name(), str(p.value()))
    with open(filename, "w") as fd:
        config.write(fd)
    return filename
