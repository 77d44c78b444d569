"""The host C library under AddressSanitizer + UndefinedBehaviorSanitizer on a
corpus of scenario files: the fixtures, seeded random files and hand-made
pathological inputs.  The reader handles files users hand it, so it must stay
memory-safe where the reference is not (e.g. names longer than MAX_NAME overflow
a 32-byte buffer in sim_gravity_init, sim_gravity.c:341-342)."""
import glob
import os
import random
import subprocess

import pytest

from conftest import GOLDEN, REFERENCE_DIR, ROOT


def random_scenario(rng):
    """Looser than the reference-parity fuzzer: arbitrary junk is welcome here."""
    out = []
    for _ in range(rng.randint(0, 40)):
        k = rng.random()
        if k < 0.5:
            out.append(" " * rng.randint(0, 5) + " ".join(rng.choice(["x", "1e308", "-1e-320", "nan", "inf", "0x1p3", "RED", "", "7"])
                                                          for _ in range(rng.randint(0, 12))))
        elif k < 0.7:
            out.append("PARAM " + " ".join(rng.choice(["DT_LINUX", "SIM_WIDTH", "1e30", "-1", "SEC", "AU", "LY", "x" * 150])
                                           for _ in range(rng.randint(0, 5))))
        else:
            out.append("".join(chr(rng.randint(1, 255)) for _ in range(rng.randint(0, 80))))
    return "\n".join(out)


def test_reader_under_asan_and_ubsan(tmp_path):
    exe = str(tmp_path / "host_sanitize")
    cc = subprocess.run(["gcc", "-O1", "-g", "-fsanitize=address,undefined,float-cast-overflow", "-fno-sanitize-recover=all",
                         "-I", os.path.join(ROOT, "include"), "-o", exe,
                         os.path.join(ROOT, "tests", "host_sanitize_driver.c"),
                         os.path.join(ROOT, "proj_entropy_b200", "host", "gravity_host.c"), "-lm"],
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if cc.returncode != 0:
        pytest.skip("sanitizer runtimes not available: " + cc.stderr[-300:])
    corpus = sorted(glob.glob(os.path.join(GOLDEN, "scenarios", "*")))
    corpus += sorted(glob.glob(os.path.join(REFERENCE_DIR, "sim_gravity", "*")))
    rng = random.Random(77)
    for k in range(150):
        p = tmp_path / ("junk%d" % k)
        p.write_bytes(random_scenario(rng).encode("latin-1"))
        corpus.append(str(p))
    special = {
        "long_name": "a" * 500 + " 0 0 0 0 1E30 1 RED 1\n",
        "long_color": "a 0 0 0 0 1E30 1 " + "C" * 5000 + " 1\n",
        "deep_indent": "".join(" " * i + "b%d 1 1 1 1 1E20 1 RED 1\n" % i for i in range(130)),
        "huge_line": "x" * 300000 + "\n",
        "megabyte": ("b 1 2 3 4 5E20 6 RED 7\n" * 60000),               # > 1,000,000 bytes: truncated like the reference
        "nul_inside": "a 0 0 0 0 1E30 1 RED 1\n\0b 0 0 0 0 1E30 1 RED 1\n",
        "no_newline_at_end": "a 0 0 0 0 1E30 1 RED 1",
        "int_overflow_fields": "a 0 0 0 0 1E30 1 RED 99999999999999999999\nPARAM DT_LINUX 1e300 SEC\n",
        "many_objects": "".join("o%d %d 0 0 0 1E20 1 RED 1\n" % (i, i) for i in range(500)),
        "crlf": "PARAM DT_LINUX 5 SEC\r\na 0 0 0 0 1E30 1 RED 1\r\n",
    }
    for name, text in special.items():
        p = tmp_path / name
        p.write_bytes(text.encode("latin-1"))
        corpus.append(str(p))
    corpus.append(str(tmp_path / "does_not_exist"))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    p = subprocess.run([exe, str(tmp_path)] + corpus, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-3000:]
    assert "runtime error" not in p.stderr and "AddressSanitizer" not in p.stderr, p.stderr[-3000:]
    loaded = int(p.stdout.split()[1])
    assert loaded >= 8
