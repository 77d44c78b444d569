"""oracle/oracle_py.py -- TEST INFRASTRUCTURE: ctypes binding of liboracle.so.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference
legs of bench.py may import this module.  The product never does.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_int64)
_LIB = None
G = 6.673E-11
REF_BINARY = os.path.join(_HERE, "_ref", "ref_gravity")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
        l = C.CDLL(path)
        common = [C.c_int64, _dp, _dp, _dp, _dp, _dp, C.c_int32, _lp, C.c_int64, C.c_int, _dp, _dp]
        l.oracle_accel.argtypes = common
        l.oracle_accel.restype = None
        l.oracle_accel_ld.argtypes = common
        l.oracle_accel_ld.restype = None
        l.oracle_step.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, _dp, C.c_int32, C.c_int64, C.c_int]
        l.oracle_step.restype = None
        l.oracle_energy.argtypes = [C.c_int64, _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp]
        l.oracle_energy.restype = None
        _LIB = l
    return _LIB


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def _accel(fn, mass, x, y, vx, vy, dt, rows, nthreads):
    mass, x, y, vx, vy = map(_c, (mass, x, y, vx, vy))
    n = len(mass)
    if rows is None:
        nrows, rp = n, None
    else:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        nrows, rp = len(rows), rows.ctypes.data_as(_lp)
    ax = np.empty(nrows)
    ay = np.empty(nrows)
    fn(n, _p(mass), _p(x), _p(y), _p(vx), _p(vy), int(dt), rp, nrows, int(nthreads), _p(ax), _p(ay))
    return ax, ay


def accel(mass, x, y, vx, vy, dt, rows=None, nthreads=1):
    """AX, AY as at sim_gravity.c:1207-1208 (fp64, reference arithmetic)."""
    return _accel(lib().oracle_accel, mass, x, y, vx, vy, dt, rows, nthreads)


def accel_ld(mass, x, y, vx, vy, dt, rows=None, nthreads=1):
    """Same sums in x87 long double: a higher-precision yardstick."""
    return _accel(lib().oracle_accel_ld, mass, x, y, vx, vy, dt, rows, nthreads)


def step(mass, x, y, vx, vy, dt, nsteps=1, nthreads=1):
    """nsteps reference steps; returns new (x, y, vx, vy), inputs untouched."""
    mass = _c(mass)
    x, y, vx, vy = (_c(a).copy() for a in (x, y, vx, vy))
    lib().oracle_step(len(mass), _p(mass), _p(x), _p(y), _p(vx), _p(vy), int(dt), int(nsteps), int(nthreads))
    return x, y, vx, vy


def energy(mass, x, y, vx, vy, nthreads=1):
    mass, x, y, vx, vy = map(_c, (mass, x, y, vx, vy))
    ke, pe = C.c_double(), C.c_double()
    lib().oracle_energy(len(mass), _p(mass), _p(x), _p(y), _p(vx), _p(vy), int(nthreads), C.byref(ke), C.byref(pe))
    return ke.value, pe.value


def ref_available():
    return os.path.exists(REF_BINARY) and os.access(REF_BINARY, os.X_OK)


def ref_run(scenario_path, steps):
    """Run the UNMODIFIED reference TU (oracle/_ref/ref_gravity); list of dicts."""
    p = subprocess.run([REF_BINARY, scenario_path] + [str(s) for s in steps], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    return p.returncode, [json.loads(l) for l in p.stdout.splitlines() if l.strip()]
