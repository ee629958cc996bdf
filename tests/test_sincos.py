"""CPU: the sinf/cosf restatement used by the position-integration kernel (fx_math.cuh, fx_sincosf)
equals this machine's glibc bit for bit on EVERY float the integrator can produce -- angles are
normalised to [pi, 3*pi) (reference src/rigid_body.c:228, 589-591) -- and on a wider sweep; and
glibc's sinf is odd / cosf even on that range, which lets the narrow phase reuse the cached sin/cos
for frVector2Rotate(v, -angle) (reference src/collision.c:328-330, 646, 715)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_checker(tmp):
    so = os.path.join(tmp, "libsincos_check.so")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "c", "sincos_check.cpp")])
    L = C.CDLL(so)
    L.check_range.restype = C.c_long
    L.check_range.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    return L


def test_sincos_restatement_exhaustive(tmp_path):
    L = build_checker(str(tmp_path))
    sym, nrm = C.c_long(), C.c_long()
    lo, hi = float(np.float32(math.pi) - np.float32(1e-3)), float(np.float32(3 * math.pi) + np.float32(1e-3))
    assert L.check_range(lo, hi, C.byref(sym), C.byref(nrm)) == 0          # ~13.5 M floats
    assert sym.value == 0
    assert L.check_range(0.00025, 0.8, C.byref(sym), C.byref(nrm)) == 0     # small-angle and pi/4 branches
    assert sym.value == 0
    assert L.check_range(9.5, 119.9, C.byref(sym), C.byref(nrm)) == 0       # rest of the fast-reduction range
    assert sym.value == 0
