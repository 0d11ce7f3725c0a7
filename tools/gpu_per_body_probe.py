"""Developer probe: where the per-object API loop (cpBodySetForce x N, cpSpaceStep, cpBodyGetPosition x N) spends its time.
`emu` as the second argument runs the host layer against the CPU emulator build (host-side costs only)."""
import os, sys
os.environ["CPB_E2E_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chipmunk2d_b200.api as api_mod
if len(sys.argv) > 2 and sys.argv[2] == "emu":
    import chipmunk2d_b200.engine as engine
    engine.ENGINE_LIB = os.path.join(ROOT, "tools/emu/_build/libcpb200_emu.so")
    api_mod.SCENE_LIB = os.path.join(ROOT, "tools/emu/_build/libscene_b200_emu.so")
from chipmunk2d_b200.scenes import circle_pile
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sc = circle_pile(n, dense=True, sleep=0.5)
api = api_mod.SceneSpace(api_mod.load_scene_lib(), sc.blob)
api.step(sc.dt, 2)
api.e2e_steps(sc.dt, 1)
sec, _ = api.e2e_steps(sc.dt, steps)
print("ms per step %.2f" % (1000 * sec / steps), flush=True)
api.space = None          # (freeing a million objects one by one at exit takes minutes)
os._exit(0)
