"""Host-side multi-GPU logic on CPU: band arithmetic, and a world-size-2 gloo run in which every rank renders
its row band with the CPU oracle, swaps final-reservoir halos through snake-vk_b200/sharding.py and the
gathered frame must equal the single-process frame bit for bit (the N > 1 path of bench.py --shard rows)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_bands_partition():
    sh = importlib.import_module("snake-vk_b200.sharding")
    for H, world in [(1080, 1), (1080, 2), (1080, 8), (2160, 8), (37, 5), (64, 7)]:
        b = sh.RowBands(H, world)
        rows = [b.band(r) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == H
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        sizes = [r1 - r0 for r0, r1 in rows]
        assert max(sizes) - min(sizes) <= 1
        for r in range(world):
            p0, p1 = b.pass1_rows(r)
            g0, g1 = b.gbuffer_rows(r)
            assert g0 <= p0 <= rows[r][0] and rows[r][1] <= p1 <= g1 and g0 >= 0 and g1 <= H
            # what I send to a peer is exactly what the peer expects to receive from me
            for peer, send, recv in b.exchange_plan(r):
                back = {p: (s, rc) for p, s, rc in b.exchange_plan(peer)}
                assert back[r][1] == send and back[r][0] == recv


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    snk_scenes = importlib.import_module("snake-vk_b200.scenes")
    sh = importlib.import_module("snake-vk_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sc = snk_scenes.small_test_scene(48, 32)
    H, W = sc.height, sc.width
    bands = sh.RowBands(H, world, radius=4, motion_margin=3)
    st = O.default_settings(neighbour_radius=4.0)
    fr = O.OracleFrame(O.OracleScene(sc))
    b0, b1 = bands.band(rank)
    prev = None
    for f in range(4):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.02, 0.0, 0.05), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        fr.set(ubo, ubo if prev is None else prev, st)
        fr.gbuffer(rows=bands.gbuffer_rows(rank), threads=2)
        fr.render_pass(1, rows=bands.pass1_rows(rank), threads=2)
        fr.render_pass(0, rows=(b0, b1), threads=2)
        prev = ubo
        res = fr.read("reservoirs_final")                      # (H, W) records of 48 B
        t = torch.from_numpy(res.view(np.uint8).reshape(H, W * 48))
        sh.exchange_final(bands, rank, [t], dist)
        fr.write_reservoirs(0, res)
    img = torch.from_numpy(fr.read("radiance"))
    sh.gather_rows(bands, rank, img, dist, dst=0)
    # sample sharding: mean over ranks of a per-rank constant image
    acc = torch.full((4, 4), float(rank + 1))
    sh.combine_samples(acc, dist, world)
    assert torch.allclose(acc, torch.full((4, 4), (world + 1) / 2.0))
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), img.numpy())
        np.save(os.path.join(out_dir, "res_band0.npy"), fr.read("reservoirs_final")[b0:b1])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bands_equal_whole_frame(snk, oracle, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # single-process whole frame
    sc = snk.scenes.small_test_scene(48, 32)
    fr = oracle.OracleFrame(oracle.OracleScene(sc))
    st = oracle.default_settings(neighbour_radius=4.0)
    prev = None
    for f in range(4):
        pos = np.asarray(sc.camera.pos, np.float32) + np.asarray((0.02, 0.0, 0.05), np.float32) * f
        ubo = sc.ubo(f, cam_pos=pos)
        fr.set(ubo, ubo if prev is None else prev, st)
        fr.gbuffer()
        fr.render()
        prev = ubo
    want = fr.read("radiance")
    got = np.load(tmp_path / "gathered.npy")
    assert want.tobytes() == got.tobytes(), f"{(want != got).any(-1).sum()} pixels differ between 2 bands and the whole frame"
    assert (want[..., :3] > 0).mean() > 0.3
    sh = importlib.import_module("snake-vk_b200.sharding")
    b0, b1 = sh.RowBands(32, 2, radius=4, motion_margin=3).band(0)
    assert fr.read("reservoirs_final")[b0:b1].tobytes() == np.load(tmp_path / "res_band0.npy").tobytes()
