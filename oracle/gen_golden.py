#!/usr/bin/env python3
"""oracle/gen_golden.py -- TEST INFRASTRUCTURE: regenerate tests/golden/*.json.

Runs the UNMODIFIED reference translation unit (oracle/_ref/ref_gravity, built
by oracle/Makefile from /root/reference/sim_gravity.c) on

  * the reference's own six scenario files (read from /root/reference at
    generation time only; identified in the fixture by md5, never copied), and
  * the scenario files authored for this repo under tests/golden/scenarios/,

and records the parsed initial state and the state after fixed step counts as
C99 hex floats.  The fixtures travel to the GPU box; /root/reference and this
script's inputs need not exist there.

usage (in the build container, /root/reference mounted):
    make -C oracle && python oracle/gen_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GRAVITY_REFERENCE_DIR", "/root/reference")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_gravity")
OUT = os.path.join(ROOT, "tests", "golden")

BASE_STEPS = [0, 1, 2, 3, 10, 100, 1000, 10000]
REFERENCE_SCENARIOS = {
    "two_body_circular": BASE_STEPS,
    "two_body_elliptical": BASE_STEPS,
    "three_body": BASE_STEPS,
    "gravity_boost": BASE_STEPS,
    "solar_system": BASE_STEPS,
    # BASELINE.json configs[1]: 10^5 steps, checked every 10^4th
    "solar_system_and_rogue_star": BASE_STEPS + [k * 10000 for k in range(2, 11)],
}
AUTHORED_STEPS = [0, 1, 2, 3, 10, 100, 1000]
AUTHORED_STEPS_LONG = {"solar_system_daily": [0, 1, 2, 3, 10, 100, 1000, 10000]}


def run(path, steps):
    p = subprocess.run([HARNESS, path] + [str(s) for s in steps],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.strip()]
    return p.returncode, lines


def fixture(name, path, steps, source):
    md5 = hashlib.md5(open(path, "rb").read()).hexdigest()
    rc, lines = run(path, steps)
    fx = {"scenario": name, "source": source, "file_md5": md5,
          "generator": "oracle/gen_golden.py via oracle/_ref/ref_gravity (unmodified sim_gravity.c, gcc -O2)"}
    if rc != 0:
        fx["error"] = lines[0]["error"][:1] + [os.path.basename(lines[0]["error"][1])] + lines[0]["error"][2:]
        return fx
    first = lines[0]
    fx.update({
        "n": first["n"], "dt": first["dt"], "sim_width": first["sim_width"],
        "names": [b["name"] for b in first["bodies"]],
        "mass": [b["mass"] for b in first["bodies"]],
        "radius": [b["radius"] for b in first["bodies"]],
        "checkpoints": [
            {"steps": l["steps"], "sim_time": l["sim_time"],
             # PATH ring of the commit block (sim_gravity.c:1248-1260): samples taken so far and
             # their digest (oracle/ref_harness/path_digest.h)
             "path_tail": l["path_tail"], "path_hash": l["path_hash"],
             "x": [b["x"] for b in l["bodies"]], "y": [b["y"] for b in l["bodies"]],
             "vx": [b["vx"] for b in l["bodies"]], "vy": [b["vy"] for b in l["bodies"]]}
            for l in lines],
    })
    return fx


def author_solar_system_daily(path):
    """The parsed solar-system state (tests/golden/ref_solar_system.json, 13 bodies) written
    back in the reference's format by libgravity_host with DT = 1 day: every commit is a day
    change, the case the device-side PATH ring exists for."""
    sys.path.insert(0, ROOT)
    from proj_entropy_b200 import host
    import numpy as np
    fx = json.load(open(os.path.join(OUT, "ref_solar_system.json")))
    cp = fx["checkpoints"][0]
    f = lambda key, src=cp: np.array([float.fromhex(v) for v in src[key]])
    n = fx["n"]
    sc = host.Scenario("SOLAR_SYSTEM_DAILY", 86400, float.fromhex(fx["sim_width"]), fx["names"], f("mass", fx),
                       f("radius", fx), f("x"), f("y"), f("vx"), f("vy"), np.full(n, 9, np.int32),
                       np.full(n, 400, np.int32))
    sc.save(path)


def main():
    if not os.path.exists(HARNESS):
        sys.exit("build oracle/_ref first: make -C oracle")
    os.makedirs(OUT, exist_ok=True)
    src_md5 = hashlib.md5(open(os.path.join(REF, "sim_gravity.c"), "rb").read()).hexdigest()
    for name, steps in REFERENCE_SCENARIOS.items():
        fx = fixture(name, os.path.join(REF, "sim_gravity", name), steps, "reference:sim_gravity/" + name)
        fx["sim_gravity_c_md5"] = src_md5
        json.dump(fx, open(os.path.join(OUT, "ref_%s.json" % name), "w"), indent=1)
        print("ref_%s.json" % name, fx.get("n"), fx.get("dt"))
    sdir = os.path.join(OUT, "scenarios")
    author_solar_system_daily(os.path.join(sdir, "solar_system_daily"))
    for name in sorted(os.listdir(sdir)):
        fx = fixture(name, os.path.join(sdir, name), AUTHORED_STEPS_LONG.get(name, AUTHORED_STEPS),
                     "authored:tests/golden/scenarios/" + name)
        fx["sim_gravity_c_md5"] = src_md5
        json.dump(fx, open(os.path.join(OUT, "own_%s.json" % name), "w"), indent=1)
        print("own_%s.json" % name, fx.get("n"), fx.get("dt"), fx.get("error"))


if __name__ == "__main__":
    main()
