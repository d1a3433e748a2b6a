"""Loaders of the CHECKERS (test infrastructure, never the product): the scalar C oracle build and the unmodified
reference compiled from /root/reference.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; the product package (pixelforge_b200/) knows nothing about oracle/.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pixelforge_b200.binding import SceneLib, PfcuLib  # noqa: E402


def load_oracle_scenes():
    """Scene runner linked against the C99 front end + oracle/pfcu_oracle.c (scalar restatement of the fragment path)."""
    return SceneLib(os.path.join(ROOT, "oracle", "_build", "libpfscenes_oracle.so"))


def load_reference_scenes(bilinear_fix=False):
    """Scene runner linked against the unmodified reference (oracle/_ref, built by oracle/build_ref.sh);
    bilinear_fix selects the build with the one-token Q7 fix (src/internal/color.h:141)."""
    name = "libpfscenes_ref_bfix.so" if bilinear_fix else "libpfscenes_ref.so"
    return SceneLib(os.path.join(ROOT, "oracle", "_ref", name))


def load_oracle_pfcu():
    """The pfcu C-ABI implemented by the scalar C oracle."""
    return PfcuLib(os.path.join(ROOT, "oracle", "_build", "libpixelforge_oracle.so"))
