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
             "x": [b["x"] for b in l["bodies"]], "y": [b["y"] for b in l["bodies"]],
             "vx": [b["vx"] for b in l["bodies"]], "vy": [b["vy"] for b in l["bodies"]]}
            for l in lines],
    })
    return fx


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
    for name in sorted(os.listdir(sdir)):
        fx = fixture(name, os.path.join(sdir, name), AUTHORED_STEPS, "authored:tests/golden/scenarios/" + name)
        fx["sim_gravity_c_md5"] = src_md5
        json.dump(fx, open(os.path.join(OUT, "own_%s.json" % name), "w"), indent=1)
        print("own_%s.json" % name, fx.get("n"), fx.get("dt"), fx.get("error"))


if __name__ == "__main__":
    main()
