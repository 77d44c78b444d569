"""ctypes mirror of include/gravity_host.h (libgravity_host: scenario reader,
Plummer model, day sampling).  Pure host C underneath; no GPU needed."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

COLORS = ["PURPLE", "BLUE", "LIGHT_BLUE", "GREEN", "YELLOW", "ORANGE", "RED", "GRAY", "PINK", "WHITE"]


class ScenarioError(RuntimeError):
    """Carries the reference's three error_str rows (sim_gravity.c:136)."""

    def __init__(self, rows):
        super().__init__(" | ".join(r for r in rows if r))
        self.rows = list(rows)


class _CScenario(C.Structure):
    _fields_ = [("sim_name", C.c_char * 32), ("sim_time", C.c_int64), ("dt", C.c_int32), ("n", C.c_int32),
                ("sim_width", C.c_double), ("name", C.POINTER(C.c_char * 32)),
                ("mass", _dp), ("radius", _dp), ("x", _dp), ("y", _dp), ("vx", _dp), ("vy", _dp),
                ("disp_color", _ip), ("max_path_disp_dflt", _ip), ("error_str", (C.c_char * 200) * 3)]


def load_host_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libgravity_host.so")
    if not os.path.exists(path):
        raise RuntimeError("%s is not built (run `make` or __graft_entry__.build())" % path)
    lib = C.CDLL(path)
    SP = C.POINTER(_CScenario)
    sigs = {
        "gravity_scenario_load_ex": (C.c_int, [C.c_char_p, C.c_int32, C.c_double, C.c_int32, C.POINTER(SP)]),
        "gravity_scenario_load": (C.c_int, [C.c_char_p, C.POINTER(SP)]),
        "gravity_scenario_free": (None, [SP]),
        "gravity_scenario_save": (C.c_int, [SP, C.c_char_p]),
        "gravity_scenario_save_error": (C.c_char_p, [C.c_int]),
        "gravity_plummer_generate": (C.c_int, [C.c_int64, C.c_uint64, _dp, _dp, _dp, _dp, _dp]),
        "gravity_steps_until_day_change": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
        "gravity_snapshot_write": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                             _dp, _dp, _dp, _dp, _dp]),
        "gravity_snapshot_write_path": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                                  _dp, _dp, _dp, _dp, _dp, C.c_int64, C.c_int64, _dp]),
        "gravity_snapshot_read_path": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                 C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp),
                                                 C.POINTER(_dp), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                                 C.POINTER(_dp)]),
        "gravity_snapshot_read": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                            C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp),
                                            C.POINTER(_dp)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._declared = sorted(sigs)
    _LIB = lib
    return lib


class Scenario:
    """Parsed scenario: numpy copies of the SoA state plus DT, SIM_WIDTH, names."""

    def __init__(self, sim_name, dt, sim_width, names, mass, radius, x, y, vx, vy, colors, path_disp):
        self.sim_name, self.dt, self.sim_width, self.names = sim_name, dt, sim_width, names
        self.mass, self.radius, self.x, self.y, self.vx, self.vy = mass, radius, x, y, vx, vy
        self.colors, self.max_path_disp_dflt = colors, path_disp
        self.sim_time = 0
        self.n = len(names)

    def save(self, pathname):
        lib = load_host_library()
        n = self.n
        names = (C.c_char * 32 * n)()
        for i, nm in enumerate(self.names):
            names[i].value = nm.encode()[:31]
        col = np.ascontiguousarray(self.colors, dtype=np.int32)
        pd = np.ascontiguousarray(self.max_path_disp_dflt, dtype=np.int32)
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (self.mass, self.radius, self.x, self.y, self.vx, self.vy)]
        cs = _CScenario()
        cs.sim_name = self.sim_name.encode()[:31]
        cs.sim_time, cs.dt, cs.n, cs.sim_width = int(self.sim_time), int(self.dt), n, float(self.sim_width)
        cs.name = C.cast(names, C.POINTER(C.c_char * 32))
        cs.mass, cs.radius, cs.x, cs.y, cs.vx, cs.vy = [a.ctypes.data_as(_dp) for a in arrs]
        cs.disp_color, cs.max_path_disp_dflt = col.ctypes.data_as(_ip), pd.ctypes.data_as(_ip)
        rc = lib.gravity_scenario_save(C.byref(cs), os.fsencode(pathname))
        if rc == -1:
            raise OSError("cannot write %s" % pathname)
        if rc != 0:
            raise ValueError("%s: %s" % (pathname, lib.gravity_scenario_save_error(rc).decode()))


def load_scenario(pathname, inherit_dt=0, inherit_sim_width=0.0, max_objects=0):
    """gravity_scenario_load_ex; raises ScenarioError with the reference's rows."""
    lib = load_host_library()
    sp = C.POINTER(_CScenario)()
    rc = lib.gravity_scenario_load_ex(os.fsencode(pathname), int(inherit_dt), float(inherit_sim_width),
                                      int(max_objects), C.byref(sp))
    if not sp:
        raise MemoryError("gravity_scenario_load")
    try:
        s = sp.contents
        if rc != 0:
            raise ScenarioError([s.error_str[k].value.decode() for k in range(3)])
        n = s.n
        arr = lambda p, dt=np.float64: np.ctypeslib.as_array(p, shape=(n,)).astype(dt).copy() if n else np.zeros(0, dt)
        names = [s.name[i].value.decode() for i in range(n)]
        return Scenario(s.sim_name.decode(), s.dt, s.sim_width, names, arr(s.mass), arr(s.radius), arr(s.x),
                        arr(s.y), arr(s.vx), arr(s.vy), arr(s.disp_color, np.int32),
                        arr(s.max_path_disp_dflt, np.int32))
    finally:
        lib.gravity_scenario_free(sp)


def plummer(n, seed=1):
    """(mass, x, y, vx, vy) of the planar Plummer model, gravity_plummer_generate."""
    lib = load_host_library()
    out = [np.empty(n, dtype=np.float64) for _ in range(5)]
    if lib.gravity_plummer_generate(int(n), int(seed), *[a.ctypes.data_as(_dp) for a in out]) != 0:
        raise ValueError("gravity_plummer_generate(%d)" % n)
    return tuple(out)


def steps_until_day_change(sim_time, dt, last_day):
    return load_host_library().gravity_steps_until_day_change(int(sim_time), int(dt), int(last_day))


def snapshot_write(pathname, sim_time, dt, last_day, mass, x, y, vx, vy, path_tail=0, path_xy=None):
    """path_xy: the most recent PATH samples, oldest first, shape [samples, n, 2] (optional)."""
    lib = load_host_library()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (mass, x, y, vx, vy)]
    n = len(arrs[0])
    if path_xy is None:
        rc = lib.gravity_snapshot_write(os.fsencode(pathname), n, int(sim_time), int(dt), int(last_day),
                                        *[a.ctypes.data_as(_dp) for a in arrs])
    else:
        p = np.ascontiguousarray(path_xy, dtype=np.float64).reshape(-1, n, 2)
        rc = lib.gravity_snapshot_write_path(os.fsencode(pathname), n, int(sim_time), int(dt), int(last_day),
                                             *[a.ctypes.data_as(_dp) for a in arrs], int(path_tail), len(p),
                                             p.ctypes.data_as(_dp))
    if rc != 0:
        raise OSError("cannot write snapshot %s" % pathname)


def snapshot_read(pathname):
    """-> dict(n, sim_time, dt, last_day, mass, x, y, vx, vy, path_tail, path_xy [samples, n, 2])"""
    lib = load_host_library()
    n, t, dt, ld = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
    tail, count, path = C.c_int64(), C.c_int64(), _dp()
    ptrs = [_dp() for _ in range(5)]
    if lib.gravity_snapshot_read_path(os.fsencode(pathname), C.byref(n), C.byref(t), C.byref(dt), C.byref(ld),
                                      *[C.byref(p) for p in ptrs], C.byref(tail), C.byref(count), C.byref(path)) != 0:
        raise OSError("cannot read snapshot %s" % pathname)
    libc = C.CDLL(None)
    try:
        cols = [np.ctypeslib.as_array(p, shape=(n.value,)).copy() for p in ptrs]
        pxy = np.zeros((0, n.value, 2))
        if count.value > 0:
            pxy = np.ctypeslib.as_array(path, shape=(count.value, n.value, 2)).copy()
    finally:
        libc.free(C.cast(ptrs[0], C.c_void_p))
        if path:
            libc.free(C.cast(path, C.c_void_p))
    return dict(n=n.value, sim_time=t.value, dt=dt.value, last_day=ld.value, mass=cols[0], x=cols[1], y=cols[2],
                vx=cols[3], vy=cols[4], path_tail=tail.value, path_xy=pxy)
