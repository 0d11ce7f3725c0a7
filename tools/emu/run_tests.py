#!/usr/bin/env python
"""Developer tool: run the `-m gpu` tests against the CPB_EMU build (kernel logic check in the
GPU-less container).  Not used by the driver; the real tests run on the B200 box."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import chipmunk2d_b200.engine as engine  # noqa: E402

engine.ENGINE_LIB = os.path.join(ROOT, "tools/emu/_build/libcpb200_emu.so")
import chipmunk2d_b200.api as api  # noqa: E402

api.SCENE_LIB = os.path.join(ROOT, "tools/emu/_build/libscene_b200_emu.so")
sys.exit(pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + (sys.argv[1:] or [os.path.join(ROOT, "tests")])))
