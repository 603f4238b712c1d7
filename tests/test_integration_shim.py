"""integration/apply_shim.py against the reference tree of this container: every anchor the patches
hang on is still there, the patched files carry the calls, and nothing is written outside the
scratch copy. (The compiled result, entity.xc on libentity_b200.so, is tests/test_gpu_shim.py.)"""
import importlib.util
import os
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = ["src/engines/srpic/fieldsolvers.h", "src/engines/srpic/currents.h",
         "src/engines/srpic/particle_pusher.h", "src/framework/domain/metadomain_sort.cpp"]


def test_patches_apply_to_a_scratch_copy(tmp_path):
    if not os.path.isdir(REF):
        pytest.skip("no reference tree here")
    for f in FILES:
        os.makedirs(os.path.dirname(tmp_path / f), exist_ok=True)
        shutil.copy(os.path.join(REF, f), tmp_path / f)
    spec = importlib.util.spec_from_file_location("apply_shim", os.path.join(ROOT, "integration", "apply_shim.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    before = {f: os.path.getmtime(os.path.join(REF, f)) for f in FILES}
    m.main(str(tmp_path))
    m.main(str(tmp_path))  # idempotent: starts from the pristine files
    want = {"src/engines/srpic/fieldsolvers.h": ["eb200_faraday(", "eb200_ampere(", "eb200_currents_ampere("],
            "src/engines/srpic/currents.h": ["eb200_deposit(", "eb200_filter("],
            "src/engines/srpic/particle_pusher.h": ["eb200_push_deposit_sr(", "eb200_push_sr("],
            "src/framework/domain/metadomain_sort.cpp": ["eb200_sort_particles("]}
    for f, calls in want.items():
        text = open(tmp_path / f).read()
        for c in calls:
            assert text.count(c) == 1, (f, c, text.count(c))
        assert text.count("#define EB200_SHIM 1") == 1
    assert {f: os.path.getmtime(os.path.join(REF, f)) for f in FILES} == before
    with pytest.raises(AssertionError):
        m.main(REF)


def test_shim_header_uses_only_declared_entry_points():
    import re
    hdr = open(os.path.join(ROOT, "include", "entity_b200.h")).read()
    declared = set(re.findall(r"\b(eb200_[a-z0-9_]+)\s*\(", hdr))
    used = set()
    for f in ("eb200_shim.hpp", "apply_shim.py"):
        used |= set(re.findall(r"\b(eb200_[a-z0-9_]+)\(", open(os.path.join(ROOT, "integration", f)).read()))
    used -= {"eb200_shim"}
    assert used and used <= declared, used - declared
