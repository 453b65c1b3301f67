"""CPU tests: the C-ABI library loads and exports every symbol the header
declares, and the host-side mirror of the reference's plugin interface (priors,
parameter layout, walker draws, config) behaves like the reference.  No compute
calls -- there is no GPU here and no CPU fallback to call."""
import os
import re

import numpy as np
import pytest

from helpers import Case, lnpost_cases, grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _full_model(case):
    from spamm_b200.Model import Model
    from spamm_b200.Spectrum import Spectrum
    from spamm_b200.components.NuclearContinuumComponent import NuclearContinuumComponent
    from spamm_b200.components.FeComponent import FeComponent
    from spamm_b200.components.HostGalaxyComponent import HostGalaxyComponent
    from spamm_b200.components.BalmerContinuumCombined import BalmerCombined
    m = Model()
    m.components += [NuclearContinuumComponent(), FeComponent(), HostGalaxyComponent(),
                     BalmerCombined(BalmerContinuum=True, BalmerPseudocContinuum=True)]
    m.data_spectrum = Spectrum(case.wl, case.flux, case.err)
    return m


def test_library_exports_every_declared_symbol():
    from spamm_b200 import _build, _lib
    _build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "spamm_b200.h")).read()
    declared = set(re.findall(r"\b(spamm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.spamm_abi_version() == _lib.ABI_VERSION


def test_ctypes_struct_matches_header_layout():
    """Compile a tiny C program against the header and compare sizeof/offsetof."""
    import ctypes
    import subprocess
    import tempfile
    from spamm_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "spamm_b200.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %d %d\n", sizeof(SpammTemplateSet), sizeof(SpammPlanDesc),
    offsetof(SpammPlanDesc, fe), offsetof(SpammPlanDesc, host), offsetof(SpammPlanDesc, template_mode),
    offsetof(SpammPlanDesc, sh95_z), offsetof(SpammPlanDesc, max_walkers), offsetof(SpammPlanDesc, ext_k),
    SPAMM_ABI_VERSION, (int)SPAMM_COMP_EXTINCTION);
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    got = [ctypes.sizeof(_lib.SpammTemplateSet), ctypes.sizeof(_lib.SpammPlanDesc),
           _lib.SpammPlanDesc.fe.offset, _lib.SpammPlanDesc.host.offset,
           _lib.SpammPlanDesc.template_mode.offset, _lib.SpammPlanDesc.sh95_z.offset,
           _lib.SpammPlanDesc.max_walkers.offset, _lib.SpammPlanDesc.ext_k.offset, _lib.ABI_VERSION,
           _lib.COMP_EXTINCTION]
    assert [int(v) for v in out] == got


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from spamm_b200 import _lib
    z, _ = lnpost_cases()
    m = _full_model(Case(z, "full_fp6"))
    with pytest.raises(_lib.SpammError):
        m.lnposterior_batch(np.zeros((1, 20)))
    with pytest.raises(_lib.SpammError):
        m.components[0].flux(m.data_spectrum, [1e-15, 1.0])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spamm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), (dp, f)
                assert "spamm_numpy" not in txt and "refshim" not in txt, (dp, f)


def test_parameter_layout_and_names():
    z, _ = lnpost_cases()
    m = _full_model(Case(z, "full_fp6"))
    assert m.total_parameter_count == 20
    assert m.model_parameter_names() == (
        ["norm_PL", "slope1"] + ["fe_norm_%d" % i for i in (1, 2, 3)] + ["fe_width"]
        + ["hg_norm_%d" % i for i in range(1, 8)] + ["hg_stellar_disp"]
        + ["bc_norm", "bc_Te", "bc_tauBE", "bc_loffset", "bc_lwidth", "bc_logNe"])
    fe = m.components[1]
    assert fe.parameter_index("fe_width") == 3 and fe.parameter_index("nope") is None
    assert [len(t.spectral_axis) for t in fe.fe_templ] == [4181, 584, 4000]
    assert len(m.components[2].host_gal) == 7
    with pytest.raises(Exception):
        from spamm_b200.Model import Model
        from spamm_b200.Spectrum import Spectrum
        Model().data_spectrum = Spectrum([1., 2., 3.], [1., 1., 1.], [1., 1., 1.])


def test_walker_draws_and_priors_match_reference():
    """Same RNG call order as Model.run_mcmc:254-259 => seed 42 reproduces the
    reference's walkers; prior boxes resolve to the reference's bounds."""
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    m = _full_model(c)
    np.random.seed(42)
    w = np.array(m.initial_walkers(40))
    assert np.array_equal(w, c.params[:40])
    lo, hi = [], []
    for comp in m.components:
        l, h = comp.prior_bounds()
        lo += l
        hi += h
    assert np.array_equal(np.array(lo, dtype=float), c.lo)
    assert np.allclose(np.array(hi, dtype=float), c.hi, rtol=1e-15, atol=0)
    # Model.prior: 0 inside, -inf outside, strict at the bound
    for p, want in zip(c.params, c.lnpost):
        assert (m.prior(p) == 0) == np.isfinite(want)
    edge = c.params[0].copy()
    edge[1] = hi[1]
    assert m.prior(edge) == -np.inf


def test_unresolved_bounds_raise_like_reference():
    """ln_priors before initial_values compares floats with strings (App. B Q11)."""
    from spamm_b200.components.NuclearContinuumComponent import NuclearContinuumComponent
    nc = NuclearContinuumComponent()
    with pytest.raises(TypeError):
        nc.ln_priors([1e-15, 1.0])


def test_parse_pars_defaults_and_yaml_override(tmp_path):
    from spamm_b200.utils.parse_pars import parse_pars
    p = parse_pars()
    assert p["fe_forest"]["fe_kernel_size_sigma"] == 3 and p["host_galaxy"]["hg_kernel_size_sigma"] == 10
    assert p["balmer_continuum"]["bc_lines_max"] == 400. and p["rebin_spec"] is False
    f = tmp_path / "pars.yaml"
    f.write_text("fe_forest:\n    fe_width_max: 8000.\nbalmer_continuum:\n    bc_line_type: gaussian\n")
    q = parse_pars(str(f))
    assert q["fe_forest"]["fe_width_max"] == 8000. and q["fe_forest"]["fe_width_min"] == 901.
    assert q["balmer_continuum"]["bc_line_type"] == "gaussian"
    p["fe_forest"]["fe_width_max"] = 1.0          # callers get private copies
    assert parse_pars()["fe_forest"]["fe_width_max"] == 10000.


def test_spectrum_container():
    from spamm_b200.Spectrum import Spectrum
    wl = grid(101)
    s = Spectrum(wl, 2 * wl, 0.1 * wl)
    assert s.norm_wavelength == np.median(wl)
    assert abs(s.norm_wavelength_flux - 2 * np.median(wl)) < 1e-9
    assert s.grid_spacing() == wl[1] - wl[0]
    s.flux = 3 * wl
    assert abs(s.norm_wavelength_flux - 3 * np.median(wl)) < 1e-9
    with pytest.raises(AssertionError):
        Spectrum(wl, wl[:-1], wl)


def test_template_log_rebin_matches_oracle_initialise():
    from oracle.spamm_numpy import OracleModel
    z, _ = lnpost_cases()
    c = Case(z, "full_fp6")
    m = _full_model(c)
    o = OracleModel(c.wl, c.flux, c.err)
    for lt, (x, g, med) in zip(m.components[1].log_fe, o.log_fe):
        assert np.array_equal(lt.spectral_axis, x) and np.array_equal(lt.flux, g)
    for lt, (x, g, med) in zip(m.components[2].log_host, o.log_host):
        assert np.array_equal(lt.spectral_axis, x) and np.array_equal(lt.flux, g)


def test_sampler_argument_checks():
    from spamm_b200.sampler import EnsembleSampler
    with pytest.raises(AssertionError):
        EnsembleSampler(nwalkers=31, dim=2, plan=object())
    with pytest.raises(AssertionError):
        EnsembleSampler(nwalkers=30, dim=20, plan=object())


def test_balmer_continuum_pair_recurrence_accuracy():
    """The specialised kernel stages the Balmer continuum on adjacent pixel pairs: the second sample's
    expm1(h nu / kT) and exp(-tau t3) follow from the first through expm1 of the argument differences, a
    degree-6 / degree-5 Taylor series (walker_kernel.cuh, bc_f0_pair).  The same algebra in numpy against
    80-bit evaluation of the same arguments, over the part of the prior box where the kernel takes that
    path (max |d h nu| / kT <= 0.01) on the BASELINE config-5 grid: the recurrence is as accurate as the
    direct double-precision evaluation (both carry the x * 2^-53 conditioning of exp at x = h nu / kT)."""
    from spamm_b200 import constants as C
    L = np.longdouble
    wl = grid(8192)
    wl = wl[wl < 3660.0]
    h, kb, c_aa = C.H_CGS, C.KB_CGS, C.C_AA
    hnu = h * (c_aa / wl)
    t3 = (wl / 3646.0) ** 3
    h0, h1, t0, t1 = hnu[0:-1:2], hnu[1::2], t3[0:-1:2], t3[1::2]
    dh_max = np.max(np.abs(h1 - h0))
    for Te in (dh_max / (0.01 * kb) * 1.0001, 2.5e4, 5.025e4, 8.0e4, 9.99e4):
        inv_kT = 1.0 / (kb * Te)
        assert dh_max * inv_kT <= 0.01
        truth = np.expm1(L(h1) * L(inv_kT))
        direct = np.expm1(h1 * inv_kT)
        bm0 = np.expm1(h0 * inv_kT)
        dx = (h1 - h0) * inv_kT
        p = ((((dx / 720.0 + 1.0 / 120.0) * dx + 1.0 / 24.0) * dx + 1.0 / 6.0) * dx + 0.5) * dx + 1.0
        p = p * dx
        bm1 = p * (bm0 + 1.0) + bm0
        err_direct = np.max(np.abs(L(direct) - truth) / truth)
        err_rec = np.max(np.abs(L(bm1) - truth) / truth)
        assert err_rec <= 2.0 * err_direct + 5e-16, (Te, float(err_rec), float(err_direct))
        for tau in (1e-3, 0.5, 1.0, 2.0):
            truth = np.exp(-(L(tau) * L(t1)))
            e0 = np.exp(-tau * t0)
            dt = tau * (t0 - t1)
            assert np.max(np.abs(dt)) <= 0.004
            q = (((dt / 120.0 + 1.0 / 24.0) * dt + 1.0 / 6.0) * dt + 0.5) * dt + 1.0
            q = q * dt
            e1 = e0 * q + e0
            assert np.max(np.abs(L(e1) - truth) / truth) <= 6e-16, (Te, tau)
