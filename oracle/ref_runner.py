"""TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference (``spamm.Model.ln_posterior``,
spamm/Model.py:42-79) for the CPU legs of bench.py and for live cross-checks: the reference tree is
taken from ``oracle/_ref`` (made by oracle/make_ref.py; travels to the GPU box) or, in the build
container, from ``/root/reference``, and imported through oracle/refshim.py.

One OS-dependent behaviour is pinned: HostGalaxyComponent lists its templates with a bare
``glob.glob`` (HostGalaxyComponent.py:93), i.e. in directory order, which differs between file
systems; the order observed when the product's data bundle was packed (``host_names`` in
spamm_b200/data/spamm_data.npz) is imposed so that a parameter vector means the same model here, in
the fixtures and in the CUDA path.
"""
import contextlib
import io
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_REF = None


def reference_root():
    local = os.path.join(HERE, "_ref")
    if os.path.isdir(os.path.join(local, "spamm")) and os.path.isdir(os.path.join(local, "examples")):
        return local
    if os.path.isdir("/root/reference/spamm"):
        return "/root/reference"
    return None


def available():
    return reference_root() is not None


def _host_order():
    path = os.path.join(os.path.dirname(HERE), "spamm_b200", "data", "spamm_data.npz")
    if not os.path.exists(path):
        return None
    with np.load(path, allow_pickle=True) as z:
        return [str(n) for n in z["host_names"]]


class _OrderedGlob(object):
    """Stand-in for the `glob` module inside HostGalaxyComponent: same files, pinned order."""

    def __init__(self, real, order):
        self._real, self._order = real, order

    def glob(self, pattern, *a, **k):
        files = self._real.glob(pattern, *a, **k)
        names = {os.path.basename(f): f for f in files}
        if self._order and set(names) == set(self._order):
            return [names[n] for n in self._order]
        return files

    def __getattr__(self, name):
        return getattr(self._real, name)


def load():
    """The reference's classes (refshim namespace), imported once."""
    global _REF
    if _REF is None:
        root = reference_root()
        if root is None:
            raise RuntimeError("no reference tree: neither oracle/_ref nor /root/reference exists")
        os.environ["SPAMM_REFERENCE_ROOT"] = root
        from oracle import refshim
        refshim.REFERENCE_ROOT = root
        _REF = refshim.load()
        import glob as real_glob
        import sys
        hg = sys.modules["spamm.components.HostGalaxyComponent"]
        hg.glob = _OrderedGlob(real_glob, _host_order())
        _REF.refshim = refshim
    return _REF


def build_model(wl, flux, err, comps=("PL", "FE", "HOST", "BC")):
    """Reference Model with run_spamm.py:100-122's component list on a data spectrum."""
    ref = load()
    with ref.refshim.chdir_examples():
        m = ref.Model()
        for c in comps:
            if c == "PL":
                m.components.append(ref.NuclearContinuumComponent(broken=False))
            elif c == "FE":
                m.components.append(ref.FeComponent())
            elif c == "HOST":
                m.components.append(ref.HostGalaxyComponent())
            elif c == "BC":
                m.components.append(ref.BalmerCombined(BalmerContinuum=True, BalmerPseudocContinuum=True))
            else:
                raise KeyError(c)
        m.data_spectrum = ref.Spectrum(np.array(wl), np.array(flux), np.array(err))
        # the string prior sentinels ("max_flux", "fnw", ...) are only resolved inside initial_values
        # (Model.run_mcmc:254-259 always calls it before the first ln_posterior)
        state = np.random.get_state()
        for comp in m.components:
            comp.initial_values(m.data_spectrum)
        np.random.set_state(state)
    return m


def ln_posterior(model, params):
    """The reference's own ln_posterior(new_params, model); its progress prints are swallowed and the cwd is
    the one it expects (the SH95 pickle is re-opened by relative path on every call,
    BalmerContinuumCombined.py:288)."""
    ref = load()
    with ref.refshim.chdir_examples(), contextlib.redirect_stdout(io.StringIO()):
        return float(ref.ln_posterior(np.array(params, dtype=np.float64), model))
