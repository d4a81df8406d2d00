"""snkpt_sort_uints_by_key (the one-sweep radix sort of csrc/bvh_build.cu, also behind the BVH builder's Morton sort and the
particle grid's key sort) against numpy's stable sort: the reference declares `SortUintsByKey(device_keys, device_values, count)`
(Core/src/scene/ParticleSystem.cpp:14) and sorts its particle keys with a bitonic chain; the contract here is ascending keys,
equal keys in input order. Sizes sit around the tile size (4096 keys per CTA) and the look-back chain (many tiles)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _keys(kind, n, rng):
    if kind == "random32":
        return rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    if kind == "few":       # 7 distinct keys: long runs of equal digits in every pass, stability visible in the values
        return (rng.integers(0, 7, n).astype(np.uint32) * np.uint32(0x01010101))
    if kind == "equal":
        return np.full(n, 0xDEADBEEF, np.uint32)
    if kind == "sorted":
        return np.sort(rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32))
    if kind == "reversed":
        return np.sort(rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32))[::-1].copy()
    if kind == "cells":     # particle cell keys: values below n (Particle.comp:51-55)
        return rng.integers(0, max(n, 1), n).astype(np.uint32)
    raise ValueError(kind)


@pytest.mark.parametrize("n", [0, 1, 2, 33, 4095, 4096, 4097, 12289, 100_000, 3_000_001])
def test_sort_matches_numpy_stable(snk, n):
    import torch
    pt = snk.PathTracer(16, 16)
    rng = np.random.default_rng(n + 11)
    for kind in ("random32", "few", "equal", "sorted", "reversed", "cells"):
        k = _keys(kind, n, rng)
        v = np.arange(n, dtype=np.uint32)
        dk = torch.from_numpy(k.view(np.int32).copy()).cuda()
        dv = torch.from_numpy(v.view(np.int32).copy()).cuda()
        torch.cuda.synchronize()
        pt.sort_uints_by_key(dk.data_ptr(), dv.data_ptr(), n)
        pt.synchronize()
        order = np.argsort(k, kind="stable")
        assert np.array_equal(dk.cpu().numpy().view(np.uint32), k[order]), f"{kind} n={n}: keys"
        assert np.array_equal(dv.cpu().numpy().view(np.uint32), v[order]), f"{kind} n={n}: values (stability)"
    pt.close()


def test_sort_repeated_calls_and_stream(snk):
    """the status words of the look-back are cleared per call; a caller stream is honoured"""
    import torch
    pt = snk.PathTracer(16, 16)
    rng = np.random.default_rng(3)
    st = torch.cuda.Stream()
    for n in (50_000, 9_000, 50_000, 777):
        k = _keys("random32", n, rng); v = np.arange(n, dtype=np.uint32)
        with torch.cuda.stream(st):
            dk = torch.from_numpy(k.view(np.int32).copy()).cuda(non_blocking=False)
            dv = torch.from_numpy(v.view(np.int32).copy()).cuda(non_blocking=False)
            pt.sort_uints_by_key(dk.data_ptr(), dv.data_ptr(), n, stream=st.cuda_stream)
        st.synchronize()
        order = np.argsort(k, kind="stable")
        assert np.array_equal(dk.cpu().numpy().view(np.uint32), k[order])
        assert np.array_equal(dv.cpu().numpy().view(np.uint32), v[order])
    pt.close()
