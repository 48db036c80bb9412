"""Loads the TEST-ONLY host simulation of the device ops (tests/hostsim/) behind the same C ABI, so the host
orchestration can be checked against the oracle without a GPU. Never imported by peps_b200."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from peps_b200 import _lib, build  # noqa: E402

_h = None


def load():
    global _h
    if _h is None:
        _h = _lib.bind(build.build_hostsim())
    return _h
