#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Recipe for ``oracle/_ref/``: a byte-for-byte copy of the files of the
reference that its ln-posterior path needs (``spamm/``, ``utils/``, and the three data directories the
components read: ``Data/FeModels``, ``Data/HostModels``, ``Data/SH95recombcoeff``) taken from where they
lie under ``/root/reference``.

The reference is pure Python: there is nothing to compile, and its third-party imports (astropy 3.x,
specutils, emcee, ...) are absent from this image, so it is run under ``oracle/refshim.py`` with
``SPAMM_REFERENCE_ROOT=oracle/_ref``.  ``oracle/_ref/`` is git-ignored (no reference source enters the
history) but not gpurun-ignored, so the UNMODIFIED reference travels to the GPU box and
``bench.py --impl reference`` / ``cpu_baseline`` can time the reference's own ``ln_posterior``
(spamm/Model.py:42-79) on that box's host cores.

    python oracle/make_ref.py            (run by __graft_entry__.build() when /root/reference exists)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("SPAMM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PARTS = ["spamm", "utils", os.path.join("Data", "FeModels"), os.path.join("Data", "HostModels"),
         os.path.join("Data", "SH95recombcoeff")]


def available():
    """True if oracle/_ref holds a usable copy."""
    return all(os.path.isdir(os.path.join(DST, p)) for p in PARTS) and \
        os.path.isdir(os.path.join(DST, "examples"))


def make(force=False):
    if not os.path.isdir(os.path.join(SRC, "spamm")):
        return available()
    if available() and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for p in PARTS:
        shutil.copytree(os.path.join(SRC, p), os.path.join(DST, p),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    # the reference resolves its data directories relative to examples/ (utils/parameters.yaml:20,66;
    # BalmerContinuumCombined.py:288): the directory only has to exist
    os.makedirs(os.path.join(DST, "examples"), exist_ok=True)
    return True


if __name__ == "__main__":
    ok = make(force="--force" in sys.argv)
    print("oracle/_ref %s" % ("ready" if ok else "unavailable (no reference tree here)"))
