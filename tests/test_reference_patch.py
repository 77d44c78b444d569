"""patches/sim_gravity_gpu.patch applied to the REAL
reference translation unit.  oracle/Makefile applies it to a temporary copy of
/root/reference/sim_gravity.c and builds oracle/_ref/ref_gravity_gpu against
libgravity_cuda (harness_gpu.c); the binary travels to the GPU box.  There the
reference's own sim_gravity_init -> sim_gravity_thread -> commit block run with
max_thread = 1 and gravity_cuda_step_objects where the k-loop used to be, and
must reproduce the golden fixtures -- states, sim_time and the PATH rings -- bit
for bit."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

import proj_entropy_b200 as g
from conftest import REFERENCE_DIR, ROOT, bits_equal, load_fixture, ok_fixture_names, scenario_file_for

sys.path.insert(0, os.path.join(ROOT, "tools"))
import apply_reference_patch as arp

GPU_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_gravity_gpu")
GPU_HARNESS_BATCHED = os.path.join(ROOT, "oracle", "_ref", "ref_gravity_gpu_batched")
PATCH = os.path.join(ROOT, "patches", "sim_gravity_gpu.patch")
PATCH_BATCHED = os.path.join(ROOT, "patches", "sim_gravity_gpu_batched.patch")
have_harness = pytest.mark.skipif(not (os.path.exists(GPU_HARNESS) and os.path.exists(GPU_HARNESS_BATCHED)),
                                  reason="oracle/_ref/ref_gravity_gpu[_batched] not built")
have_reference = pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE_DIR, "sim_gravity.c")),
                                    reason="/root/reference is not mounted here")


def test_ed_script_applicator():
    lines = ["l%d\n" % k for k in range(1, 11)]
    script = "9,10c\nNINE-TEN\n.\n7d\n4a\nafter four\nand more\n.\n1,2c\nONE-TWO\n.\n0a\nTOP\n.\n".splitlines(keepends=True)
    out = arp.apply_ed_script(list(lines), script)
    assert out == ["TOP\n", "ONE-TWO\n", "l3\n", "l4\n", "after four\n", "and more\n", "l5\n", "l6\n", "l8\n", "NINE-TEN\n"]
    with pytest.raises(ValueError):
        arp.apply_ed_script(list(lines), "2a\nx\n.\n5d\n".splitlines(keepends=True))      # ascending: line numbers would shift
    with pytest.raises(ValueError):
        arp.apply_ed_script(list(lines), ["12d\n"])


@pytest.mark.parametrize("patch", [PATCH, PATCH_BATCHED])
def test_patch_is_an_ed_script_of_new_text_only(patch):
    """The patch names the reference's lines by number and carries only new text."""
    text = open(patch).read()
    assert "gravity_cuda_step_objects" in text and "gravity_cuda_create" in text and "gravity_cuda_destroy" in text
    assert "max_thread = 1;" in text
    ref = os.path.join(REFERENCE_DIR, "sim_gravity.c")
    if os.path.exists(ref):
        ref_lines = {l.strip() for l in open(ref).read().splitlines() if len(l.strip()) > 24}
        carried = [l for l in text.splitlines() if l.strip() in ref_lines]
        assert carried == [], carried


@have_reference
def test_patch_applies_to_the_surveyed_revision_only(tmp_path):
    out = tmp_path / "patched.c"
    arp.main(["x", os.path.join(REFERENCE_DIR, "sim_gravity.c"), str(out)])
    patched = out.read_text()
    assert patched.count("gravity_cuda_step_objects(gpu") == 1 and "static gravity_cuda_t * gpu;" in patched
    # what stays the reference's own: the day test, the PATH ring, the barrier function
    for kept in ("day_changed = day != last_day;", "obj->PATH[path_tail%MAX_PATH].X = obj->X;",
                 "int32_t sim_gravity_barrier(int32_t thread_id)", "__sync_fetch_and_add(&path_tail, 1);"):
        assert kept in patched
    # what is gone: the CPU inner loop and the NEXT_* copy
    assert "Ri      = sqrt(" not in patched and "obj->X  = obj->NEXT_X;" not in patched
    other = tmp_path / "other.c"
    other.write_text(patched)                        # a different revision: refused, line numbers would not match
    with pytest.raises(SystemExit):
        arp.main(["x", str(other), str(tmp_path / "twice.c")])
    with pytest.raises(SystemExit):                  # never writes reference source into the repository
        arp.main(["x", os.path.join(REFERENCE_DIR, "sim_gravity.c"), os.path.join(ROOT, "tests", "leak.c")])
    assert not os.path.exists(os.path.join(ROOT, "tests", "leak.c"))
    # the batching variant: same seam, the day memory moved to file scope, the clock advanced for the batch
    batched = tmp_path / "batched.c"
    arp.main(["x", os.path.join(REFERENCE_DIR, "sim_gravity.c"), str(batched), "--patch", PATCH_BATCHED])
    text = batched.read_text()
    assert text.count("gravity_cuda_step_objects(gpu") == 1 and "gravity_steps_until_day_change(sim.sim_time, dt, last_day)" in text
    assert text.count("static int32_t") >= 1 and "                static int32_t last_day;" not in text
    assert "day_changed = day != last_day;" in text and "obj->PATH[path_tail%MAX_PATH].X = obj->X;" in text


@have_harness
def test_patched_unit_reports_a_missing_gpu_through_the_reference_error_rows():
    """No GPU: sim_gravity_init fails the way its other failures do (ret = -1, error_str rows)."""
    if g.device_count() > 0:
        pytest.skip("a GPU is present")
    p = subprocess.run([GPU_HARNESS, os.path.join(ROOT, "tests", "golden", "scenarios", "random15"), "0", "5"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert p.returncode == 1
    err = json.loads(p.stdout.splitlines()[0])["error"]
    assert err[0] == "GPU Init Failed" and "no CPU fallback" in err[1]


@pytest.mark.gpu
@have_harness
@pytest.mark.parametrize("harness", [GPU_HARNESS, GPU_HARNESS_BATCHED], ids=["step_by_step", "batched"])
@pytest.mark.parametrize("name", ok_fixture_names())
def test_patched_reference_unit_reproduces_the_goldens_on_the_gpu(name, harness, tmp_path):
    """All six reference scenarios and the authored ones, every checkpoint (10^5 steps of the
    rogue-star scenario, 10^4 daily steps of the solar system): body state, sim_time, the number
    of PATH samples and the digest of the reference's own PATH rings.  Both patches: one shim call
    per step, and whole batches per call (up to the next commit that takes a PATH sample)."""
    fx = load_fixture(name)
    steps = [c["steps"] for c in fx["checkpoints"]]
    path = scenario_file_for(fx, tmp_path)
    p = subprocess.run([harness, path] + [str(s) for s in steps], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert p.returncode == 0, (p.stdout[-500:], p.stderr[-500:])
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert [l["steps"] for l in lines] == steps
    for l, cp in zip(lines, fx["checkpoints"]):
        assert l["max_thread"] == 1 and l["n"] == fx["n"] and l["dt"] == fx["dt"]
        assert l["sim_time"] == cp["sim_time"], (name, l["steps"])
        assert l["path_tail"] == cp["path_tail"] and l["path_hash"] == cp["path_hash"], (name, l["steps"])
        for key in ("x", "y", "vx", "vy"):
            got = [float.fromhex(b[key]) for b in l["bodies"]]
            assert bits_equal(got, cp[key + "_f"]), "%s step %d %s" % (name, l["steps"], key)
    last, final = lines[-1], fx["checkpoints"][-1]
    if harness == GPU_HARNESS:
        assert last["shim_calls"] == final["steps"]
    elif fx["dt"] > 0:
        # one call per requested checkpoint and per PATH sample at most (+1 for a batch cut by the limit)
        assert last["shim_calls"] <= len(steps) + final["path_tail"] + 1, (name, last["shim_calls"])
