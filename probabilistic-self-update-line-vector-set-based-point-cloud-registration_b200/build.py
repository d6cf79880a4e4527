"""Builds libpsulvsb_b200.so in-tree with nvcc for sm_100a (csrc/Makefile)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsulvsb_b200.so")


def build(force: bool = False, verbose: bool = False) -> str:
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("build did not produce " + LIB_PATH)
    return LIB_PATH
