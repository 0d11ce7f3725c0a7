"""CPU-side checks (no GPU needed): the C-ABI library loads and exports every symbol its header declares,
the drop-in API library exports the public Chipmunk2D names, and nothing works without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "chipmunk2d_b200", "lib")


def _need(path):
    if not os.path.exists(path):
        pytest.skip("%s not built (run __graft_entry__.build())" % path)
    return path


def _preprocessed(header):
    return subprocess.run(["gcc", "-E", "-P", "-I", os.path.join(ROOT, "include"), header], capture_output=True, text=True, check=True).stdout


def test_cabi_exports_every_declared_symbol():
    lib = C.CDLL(_need(os.path.join(LIB, "libcpb200.so")), mode=C.RTLD_LOCAL)
    src = _preprocessed(os.path.join(ROOT, "include", "cpb200.h"))
    names = sorted(set(re.findall(r"\b(cpb200_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_api_library_exports_the_public_chipmunk_names():
    lib = C.CDLL(_need(os.path.join(LIB, "libchipmunk_b200.so")), mode=C.RTLD_LOCAL)
    src = _preprocessed(os.path.join(ROOT, "include", "chipmunk", "chipmunk.h")) + _preprocessed(os.path.join(ROOT, "include", "chipmunk", "cpHastySpace.h"))
    # every non-inline function declaration: "... name(args);" outside of function bodies
    names = set(re.findall(r"\b(cp[A-Z][A-Za-z0-9_]*)\s*\((?!\s*\*)[^;{]*\)\s*;", src))
    inline = set(re.findall(r"static inline [A-Za-z ]+\b(cp[A-Za-z0-9_]+)\s*\(", src))
    names = sorted(n for n in names - inline if not n.endswith("Func"))
    assert len(names) > 300, len(names)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    for must in ("cpSpaceNew", "cpHastySpaceNew", "cpHastySpaceStep", "cpSpaceAddBody", "cpSpaceAddShape", "cpSpaceAddConstraint",
                 "cpSpaceStep", "cpArbiterGetContactPointSet", "cpArbiterGetNormal", "cpArbiterGetDepth", "cpArbiterTotalImpulse",
                 "cpBodyEachArbiter", "cpPinJointNew", "cpPivotJointNew", "cpDampedSpringNew", "cpGearJointNew", "cpSlideJointNew"):
        assert must in names


def test_no_cpu_fallback_world_creation_fails_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _need(os.path.join(LIB, "libcpb200.so"))
    from chipmunk2d_b200.engine import World, EngineError, load_engine
    assert load_engine().cpb200_device_available() == 0
    with pytest.raises(EngineError, match="no CPU fallback"):
        World(1)


def test_scene_blob_layouts_match_the_c_structs(tmp_path):
    """numpy mirrors in engine.py vs sizeof() of the C structs."""
    from chipmunk2d_b200 import engine as e
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "cpb200.h"\n#include "cpb_scene.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(cpb_scene_header),sizeof(cpb_scene_body),sizeof(cpb_scene_shape),sizeof(cpb_scene_joint),sizeof(cpb200_space_params),'
                   'sizeof(cpb200_body_desc),sizeof(cpb200_shape_desc),sizeof(cpb200_joint_desc),sizeof(cpb200_body_state),sizeof(cpb200_arbiter),'
                   'sizeof(cpb200_joint_state),sizeof(cpb200_stats));return 0;}\n')
    exe = str(tmp_path / "sz")
    subprocess.check_call(["gcc", "-o", exe, str(src), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "chipmunk2d_b200", "scenes")])
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mirrors = [e.SCENE_HEADER, e.SCENE_BODY, e.SCENE_SHAPE, e.SCENE_JOINT, e.SPACE_PARAMS, e.BODY_DESC, e.SHAPE_DESC, e.JOINT_DESC,
               e.BODY_STATE, e.ARBITER, e.JOINT_STATE, e.STATS]
    assert sizes == [m.itemsize for m in mirrors]


def test_generators_are_deterministic_and_well_formed():
    from chipmunk2d_b200.scenes import circle_pile, mixed_drop, batched_demo_scenes, golden_names, golden_scene
    a = circle_pile(5000, dense=True); b = circle_pile(5000, dense=True)
    assert a.blob == b.blob and a.n_dynamic() == 5000
    m = mixed_drop(3000)
    assert m.n_dynamic() == 3000 and len(m.joints) == 300
    assert set(np.unique(m.joints["type"])) == {2, 4}
    assert np.all(m.shapes["body"] < len(m.bodies))
    poly = m.shapes[m.shapes["type"] == 2]
    assert np.all(poly["vert_offset"] + poly["n_verts"] <= len(m.verts))
    scenes = batched_demo_scenes(4)
    assert [len(s.bodies) for s in scenes] == [107, 82, 107, 82]
    assert "SimpleTerrainCircles_1000" in golden_names()
    assert golden_scene("ComplexTerrainHexagons_1000").n_dynamic() == 1000
