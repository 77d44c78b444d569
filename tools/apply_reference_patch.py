#!/usr/bin/env python3
"""Apply patches/sim_gravity_gpu.patch to a TEMPORARY copy of the reference's
sim_gravity.c (nothing from the reference is written into this repository).

The patch is an ed script (the format of `diff -e`, accepted by `ed -s` and by
`patch -e`): it names the reference's lines by number and carries only the NEW
text, so it contains no reference source.  Because it addresses lines by number
the input is checked against the md5 of the revision it was written for.  This
applicator implements the `a`, `c` and `d` commands itself, so neither `ed` nor
`patch` is needed.

usage: apply_reference_patch.py <reference/sim_gravity.c> <output.c> [--patch FILE]
"""
import hashlib
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPECTED_MD5 = "ae49d860396426fc73bba122d80534a5"      # sim_gravity.c as surveyed


def apply_ed_script(lines, script):
    """lines: list of text lines (with newlines); script: list of script lines.
    Commands must come bottom-up (as diff -e emits them) so that earlier line
    numbers stay valid."""
    k, last = 0, None
    while k < len(script):
        m = re.fullmatch(r"(\d+)(?:,(\d+))?([acd])\n?", script[k])
        if not m:
            raise ValueError("not an ed command: %r" % script[k])
        lo, hi, cmd = int(m.group(1)), int(m.group(2) or m.group(1)), m.group(3)
        if last is not None and hi >= last:
            raise ValueError("commands must address descending line numbers (%d after %d)" % (hi, last))
        last = lo
        k += 1
        text = []
        if cmd in "ac":
            while script[k].rstrip("\n") != ".":
                text.append(script[k] if script[k].endswith("\n") else script[k] + "\n")
                k += 1
            k += 1
        if hi > len(lines) or lo < (0 if cmd == "a" else 1):
            raise ValueError("line %d is outside the file" % hi)
        if cmd == "a":
            lines[lo:lo] = text
        elif cmd == "c":
            lines[lo - 1:hi] = text
        else:
            del lines[lo - 1:hi]
    return lines


def main(argv):
    if len(argv) < 3:
        sys.exit(__doc__)
    src, dst = argv[1], argv[2]
    patch = os.path.join(ROOT, "patches", "sim_gravity_gpu.patch")
    if "--patch" in argv:
        patch = argv[argv.index("--patch") + 1]
    data = open(src, "rb").read()
    md5 = hashlib.md5(data).hexdigest()
    if md5 != EXPECTED_MD5:
        sys.exit("%s has md5 %s, the patch was written for %s: line numbers would not match" % (src, md5, EXPECTED_MD5))
    if os.path.realpath(dst).startswith(os.path.realpath(ROOT) + os.sep) and "/oracle/_ref/" not in os.path.realpath(dst):
        sys.exit("refusing to write patched reference source inside the repository (use a temporary directory)")
    out = apply_ed_script(data.decode().splitlines(keepends=True), open(patch).read().splitlines(keepends=True))
    with open(dst, "w") as f:
        f.writelines(out)


if __name__ == "__main__":
    main(sys.argv)
