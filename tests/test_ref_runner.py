"""The CPU arm of bench.py drives the UNMODIFIED reference from ``oracle/_ref`` (oracle/make_ref.py copies it there
in the build container; the GPU box has no /root/reference) through oracle/ref_runner.py.  Here: that path
reproduces the committed reference-made fixture of BASELINE config 5 bit for bit -- i.e. the copy is complete, the
shims are in place and the host-template order is the one the fixtures (and the CUDA path) use."""
import numpy as np
import pytest

from helpers import Case, lnpost_cases


def test_ref_runner_reproduces_the_config5_fixture():
    from oracle import ref_runner
    if not ref_runner.available():
        pytest.skip("no reference tree (oracle/_ref is made by __graft_entry__.build() where /root/reference exists)")
    z, _ = lnpost_cases()
    c = Case(z, "cfg5_full")
    m = ref_runner.build_model(c.wl, c.flux, c.err, tuple(c.comps))
    rows = [0, 1, 2, int(np.flatnonzero(np.isneginf(c.lnpost))[0]) if np.isneginf(c.lnpost).any() else 3]
    got = np.array([ref_runner.ln_posterior(m, c.params[i]) for i in rows])
    assert np.array_equal(got, c.lnpost[rows])


def test_make_ref_recipe_lists_only_what_the_path_needs():
    from oracle import make_ref
    assert make_ref.PARTS == ["spamm", "utils", "Data/FeModels", "Data/HostModels", "Data/SH95recombcoeff"]
    if make_ref.available():
        import os
        assert os.path.isdir(os.path.join(make_ref.DST, "examples"))
        assert not os.path.exists(os.path.join(make_ref.DST, "Data", "FakeData"))      # 186 MB of spectra stay behind
