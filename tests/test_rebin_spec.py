"""`rebin_spec: True` (utils/rebin_spec.py of the reference -> pysynphot).  PARITY UNPINNED: pysynphot is not
in this image and the reference ships no vector for the branch, so the product's vectorised / sparse-matrix
restatement (spamm_b200/utils/rebin_spec.py) is checked against an independent literal restatement of the
published algorithm in extended precision (oracle/rebin_longdouble.py) and through the properties the
algorithm has (flux conservation, linearity, exactness on a linear spectrum)."""
import numpy as np
import pytest

from oracle.rebin_longdouble import rebin_spec_longdouble
from spamm_b200.utils.rebin_spec import rebin_spec, rebin_matrix, compose_csr, calculate_bin_edges


def _grids():
    rng = np.random.RandomState(3)
    wave = np.sort(rng.uniform(1000.0, 2000.0, 400))
    return wave, [
        np.linspace(900.0, 2100.0, 257),              # wider than the source: zero outside, partial edge bins
        np.linspace(1200.0, 1700.0, 1000),            # finer than the source
        np.linspace(1100.0, 1900.0, 37),              # much coarser
        np.sort(rng.uniform(1001.0, 1999.0, 300)),    # irregular bins
        np.log(np.linspace(3000.0, 6000.0, 64)) * 250.0,   # a logarithmic grid
    ]


def test_matches_extended_precision_restatement():
    wave, news = _grids()
    f = np.random.RandomState(4).rand(len(wave)) + 0.1
    for new in news:
        got = rebin_spec(wave, f, new)
        want = np.asarray(rebin_spec_longdouble(wave, f, new), dtype=np.float64)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want) + 1e-300), len(new)


def test_matrix_form_and_composition():
    from scipy.sparse import csr_matrix
    wave, news = _grids()
    f = np.random.RandomState(5).rand(len(wave))
    for new in news:
        ip, ix, w = rebin_matrix(wave, new)
        assert ip.dtype == np.int32 and ip[0] == 0 and ip[-1] == len(ix) == len(w)
        assert all(np.all(np.diff(ix[ip[i]:ip[i + 1]]) > 0) for i in range(len(new)))      # ascending columns
        m = csr_matrix((w, ix, ip), shape=(len(new), len(wave)))
        assert np.allclose(m @ f, rebin_spec(wave, f, new), rtol=0, atol=1e-13)
    a = rebin_matrix(wave, news[1])
    b = rebin_matrix(news[1], news[2])
    c = compose_csr(b, a, len(news[1]))
    two_step = rebin_spec(news[1], rebin_spec(wave, f, news[1]), news[2])
    m = csr_matrix((c[2], c[1], c[0]), shape=(len(news[2]), len(wave)))
    assert np.allclose(m @ f, two_step, rtol=0, atol=1e-13)


def test_properties():
    wave, news = _grids()
    new = np.linspace(1100.0, 1900.0, 123)            # inside the source range
    # a linear spectrum is reproduced exactly at the bin centres of a uniform grid (trapezoid is exact)
    f = 2.0 + 0.003 * wave
    assert np.allclose(rebin_spec(wave, f, new), 2.0 + 0.003 * new, rtol=1e-13)
    # flux conservation: sum(binflux * bin width) == integral of the interpolant over the covered range
    g = np.random.RandomState(6).rand(len(wave))
    e = calculate_bin_edges(new)
    nodes = np.unique(np.concatenate([wave[(wave > e[0]) & (wave < e[-1])], e]))
    integral = np.trapezoid(np.interp(nodes, wave, g), nodes)
    assert abs(np.sum(rebin_spec(wave, g, new) * np.diff(e)) - integral) < 1e-10 * integral
    # linear in the input
    assert np.allclose(rebin_spec(wave, 3 * f + g, new), 3 * rebin_spec(wave, f, new) + rebin_spec(wave, g, new),
                       rtol=1e-13)
    # bins entirely outside the source are zero
    out = rebin_spec(wave, f, np.linspace(2100.0, 2500.0, 9))
    assert np.all(out == 0.0)
    with pytest.raises(ValueError):
        rebin_spec(wave[::-1], f, new)


def test_downsample_branch_of_the_data_setter():
    """Model.py:205-240: a component coarser than the data + downsample_data_if_needed moves data, model grid
    and components to the coarse grid; with upsample_components_if_needed (or neither flag) the data grid stays."""
    from spamm_b200.Model import Model
    from spamm_b200.Spectrum import Spectrum
    from spamm_b200.components.NuclearContinuumComponent import NuclearContinuumComponent

    class CoarseNC(NuclearContinuumComponent):
        seen = None

        def grid_spacing(self):
            return 5.0

        def initialize(self, data_spectrum=None):
            self.seen = len(data_spectrum.spectral_axis)

    wl = np.arange(1000.0, 2000.0, 1.0)
    flux = 1e-15 * (wl / 1500.0) ** -1.5
    for flags, n_expected in (((False, False), 1000), ((True, False), 1000), ((False, True), 200)):
        m = Model()
        m.upsample_components_if_needed, m.downsample_data_if_needed = flags
        c = CoarseNC(broken=False)
        m.components.append(c)
        m.data_spectrum = Spectrum(wl, flux, 0.05 * flux)
        assert len(m.model_spectrum.spectral_axis) == n_expected == c.seen
        assert len(m.data_spectrum.spectral_axis) == n_expected
        assert m.worst_component is c
    assert np.allclose(m.data_spectrum.spectral_axis, np.arange(1000.0, 1999.0, 5.0))
    assert np.allclose(m.data_spectrum.flux, np.interp(m.data_spectrum.spectral_axis, wl, flux))
