"""Posterior samples and summary statistics of a finished run.

Same definitions as the reference: chain flattened after `burn` iterations
(spamm/Samples.py:59-88), per-parameter histogram maximum / median / mean / mode
(:39-55), and `median_values` with index-based 68 % bounds (spamm/analysis.py:514-540).
Accepts a finished Model, a result dict, or the gzip-pickle written by
spamm_b200.run_spamm.spamm()."""
import gzip
import os
import pickle
import statistics

import numpy as np


def read_pickle(pname):
    with gzip.open(pname, "rb") as fh:
        p_data = pickle.loads(fh.read())
    return p_data["model"], p_data["comp_params"]


def _as_model(src):
    if isinstance(src, str):
        model, params = read_pickle(src)
        return model, params, os.path.basename(src).split(".")[0]
    if isinstance(src, dict):
        return src["model"], src.get("comp_params"), "model"
    return src, None, "model"


def get_samples(pname, burn=50):
    """(models, samples[n, D], params, name); a list of sources is concatenated."""
    if isinstance(pname, (list, tuple)):
        allmodels, allsamples, params = [], [], None
        for src in pname:
            model, params, _ = _as_model(src)
            allsamples.append(model.sampler.chain[:, burn:, :].reshape((-1, model.total_parameter_count)))
            allmodels.append(model)
        samples = np.concatenate(tuple(allsamples))
        return allmodels, samples, params, "concat_{}".format(len(samples))
    model, params, name = _as_model(pname)
    n_iter = np.shape(model.sampler.chain)[1]
    assert burn < n_iter, \
        "Chain burn value, {}, must be smaller than number of iterations, {}.\nRerun with lower burn " \
        "value or more iterations".format(burn, n_iter)
    samples = model.sampler.chain[:, burn:, :].reshape((-1, model.total_parameter_count))
    return model, samples, params, name


class Samples(object):
    def __init__(self, pickle_files, outdir=None, gif=False, last=False, step=100, burn=50, histbins=100):
        self.pname = pickle_files
        self.gif, self.last, self.step, self.burn, self.histbins = gif, last, step, burn, histbins
        self.models, self.samples, self.params, self.model_name = get_samples(pickle_files, burn)
        self.model = self.models[0] if isinstance(self.models, list) else self.models
        self.total_parameter_count = self.model.total_parameter_count
        self.model_parameter_names = self.model.model_parameter_names()
        self._get_stats()
        self.outdir = outdir if outdir is not None else "."

    def _get_stats(self):
        self.means, self.medians, self.modes, self.maxs = [], [], [], []
        for i in range(self.total_parameter_count):
            chain = self.samples[:, i]
            hist, bins = np.histogram(chain, self.histbins)
            binsize = bins[1] - bins[0]
            self.maxs.append(bins[np.argmax(hist)] + binsize / 2.)
            self.medians.append(np.median(chain))
            self.means.append(np.average(chain))
            self.modes.append(statistics.mode(chain))


def median_values(samples, frac=0.68):
    """[D, 3]: median, distance to the lower and to the upper bound of the central
    `frac` interval, with the reference's index arithmetic (analysis.py:514-540)."""
    num_params = np.size(samples[0, :])
    result = np.zeros((num_params, 3))
    for i in range(num_params):
        chain = np.sort(samples[:, i])
        n = np.size(chain)
        median = chain[int(n / 2)] if n % 2 == 1 else (chain[int(n / 2) - 1] + chain[int(n / 2)]) / 2.
        num_spread = n / 2.0 * frac
        max_value = chain[int(n / 2.0 + num_spread)]
        min_value = chain[int(n / 2.0 - num_spread)]
        result[i] = (median, abs(median - min_value), abs(median - max_value))
    return result


def mean_values(samples):
    return np.stack([np.mean(samples, axis=0), np.std(samples, axis=0)], axis=1)
