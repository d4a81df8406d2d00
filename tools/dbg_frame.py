import sys, importlib, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import oracle_api as O
from test_gpu_frame import bits_equal
snk = importlib.import_module("snake-vk_b200")
sc = snk.scenes.small_test_scene(96, 64)
osc = O.OracleScene(sc); fr = O.OracleFrame(osc); rt = snk.RT().Init(sc)
prev = None
for f in range(3):
    pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.05, 0, 0.1), np.float32) * f
    ubo = sc.ubo(f, cam_pos=pos); p = ubo if prev is None else prev
    fr.set(ubo, p, O.default_settings()); fr.gbuffer(); fr.render()
    rt.RecordRenderCmdBuf(ubo, p); prev = ubo
    for half in ["reservoirs_temporal", "reservoirs_final"]:
        a, b = rt.pt.read_aux(half), fr.read(half)
        eq = bits_equal(a, b)
        print(f, half, (~eq).sum())
        for (y, x) in np.argwhere(~eq)[:6]:
            print("   px", x, y, "\n     gpu", a[y, x], "\n     cpu", b[y, x])
    a, b = rt.pt.read_aux("direct")[..., :3], fr.read("direct")[..., :3]
    eq = bits_equal(a, b)
    print(f, "direct", (~eq).sum())
    for (y, x) in np.argwhere(~eq)[:8]:
        print("   px", x, y, a[y, x], b[y, x])
