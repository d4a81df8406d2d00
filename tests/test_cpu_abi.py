"""The C-ABI library loads on a CPU-only box, exports every symbol include/snkpt.h declares, keeps the wire-format
sizes of the header, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "snkpt.h")).read()
    return sorted(set(re.findall(r"SNKPT_API\s+[\w\s\*]+?\b(snkpt_\w+)\s*\(", src)))


def test_header_symbols_are_exported(snk):
    lib = snk.load_library()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(snk.api.EXPORTS) == names, "api.py EXPORTS and include/snkpt.h disagree"


def test_wire_format_sizes(snk):
    S = snk.scenes
    assert S.RESERVOIR_DTYPE.itemsize == 48 and S.RAY_DTYPE.itemsize == 32 and S.HIT_DTYPE.itemsize == 20
    assert S.MATERIAL_DTYPE.itemsize == 64 and S.UBO_DTYPE.itemsize == 240
    assert snk.api.SETTINGS_DTYPE.itemsize == 64 and snk.api.STATS_DTYPE.itemsize == 112 and snk.api.LAUNCH_TIMING_DTYPE.itemsize == 88
    s = snk.api.default_settings()
    assert (int(s["num_temporal"]), int(s["num_spatial"]), int(s["ris_iters"]), int(s["m_cap"])) == (1, 3, 32, 20)
    assert float(s["neighbour_radius"]) == 30.0 and abs(float(s["normal_threshold"]) - 0.95) < 1e-7


def test_no_cpu_fallback(snk):
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box the -m gpu suite exercises the real path
    try:
        snk.PathTracer(64, 64)
    except snk.SnkError as e:
        assert "no usable CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("snkpt_create succeeded without a CUDA device")
    lib = snk.load_library()
    assert lib.snkpt_render(None, None, None) == -1  # SNKPT_E_INVALID on a null context, never a crash


def test_light_blob_layout(snk):
    sc = snk.scenes.c2("instanced", width=64, height=36)
    blob = np.frombuffer(sc.light_bytes(), np.uint8)
    assert blob.size == 112 + 16 * 48
    assert int(blob[96:100].view(np.uint32)[0]) == 16 and int(blob[100:104].view(np.uint32)[0]) == 0
    d = blob[16:28].view(np.float32)
    assert abs(float(np.linalg.norm(d)) - 1.0) < 1e-6


def test_frame_jitter_halton_sequence(snk):
    """GetFrameJitter (RenderCommon.h:39-60): Halton(2,3) - 0.5, 8 phases at native resolution, 32 at half resolution"""
    def halton(i, b):
        f, r = 1.0, 0.0
        while i > 0:
            f *= b
            r += (i % b) / f
            i //= b
        return r
    seq = [snk.api.frame_jitter(f, 1920)[0] for f in range(17)]
    for f, j in enumerate(seq):
        k = f % 8 + 1
        assert abs(float(j[0]) - (halton(k, 2) - 0.5)) < 1e-6 and abs(float(j[1]) - (halton(k, 3) - 0.5)) < 1e-6
    assert np.array_equal(seq[0], seq[8]) and not np.array_equal(seq[0], seq[1])
    assert np.array_equal(snk.api.frame_jitter(0, 960, 1920)[0], snk.api.frame_jitter(32, 960, 1920)[0])      # 32 phases when upscaling 2x
    assert not np.array_equal(snk.api.frame_jitter(0, 960, 1920)[0], snk.api.frame_jitter(8, 960, 1920)[0])
    j, px = snk.api.frame_jitter(3, 1920)
    assert np.array_equal(px, np.float32(-0.5) * j)
    assert np.array_equal(snk.api.frame_jitter(5, 0, 1920)[0], np.zeros(2, np.float32))
