"""Multi-GPU host logic of the frame: one process per GPU, scene + BVH replicated.

Two partitions:

* **row bands** (`RowBands`): GPU r renders the contiguous rows [b0, b1). ReSTIR couples rows: spatial reuse
  (pass 2) reads temporal reservoirs up to `radius` rows away, temporal reuse (pass 1) reads the previous
  frame's FINAL reservoirs one motion vector away. So every GPU runs the G-buffer and pass 1 on a widened band
  (`snkpt_set_halo`) and, after each frame, band neighbours swap the final reservoirs of the
  `radius + motion_margin` rows on each side of their common edge (`exchange_final`) — 3 SoA planes of 16 B per
  pixel, contiguous per row range, sent with point-to-point ops (NCCL over NVLink on GPUs, gloo in the CPU tests).
  No other data-path collective exists; the image is gathered only at readback (`gather_rows`).
* **samples** (`combine_samples`): every GPU renders the whole frame with its own seed; radiance is averaged with
  one all-reduce per frame (the accumulation step).

The module only needs `torch.distributed` tensors, so the same code drives the CUDA library (planes aliased
through `snkpt_device_ptr`) and the CPU oracle in the world-size-2 gloo tests.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class RowBands:
    height: int
    world: int
    radius: int = 30          # RayGen.rgen:488 neighbour radius in pixels
    motion_margin: int = 16   # rows a motion vector may cross per frame and still be reproduced exactly

    def band(self, rank: int) -> tuple[int, int]:
        """rows [b0, b1) of `rank`; the first H % world bands are one row taller"""
        q, r = divmod(self.height, self.world)
        b0 = rank * q + min(rank, r)
        return b0, b0 + q + (1 if rank < r else 0)

    @property
    def pass1_halo(self) -> int:
        return self.radius

    @property
    def gbuffer_halo(self) -> int:
        return self.radius + self.motion_margin + 1

    @property
    def exchange_rows(self) -> int:
        return self.radius + self.motion_margin

    def clamp(self, lo: int, hi: int) -> tuple[int, int]:
        return max(lo, 0), min(hi, self.height)

    def pass1_rows(self, rank: int) -> tuple[int, int]:
        b0, b1 = self.band(rank)
        return self.clamp(b0 - self.pass1_halo, b1 + self.pass1_halo)

    def gbuffer_rows(self, rank: int) -> tuple[int, int]:
        b0, b1 = self.band(rank)
        return self.clamp(b0 - self.gbuffer_halo, b1 + self.gbuffer_halo)

    def exchange_plan(self, rank: int) -> list[tuple[int, tuple[int, int], tuple[int, int]]]:
        """[(peer, rows_to_send, rows_to_receive)]: this rank sends the part of ITS band that lies inside a peer's
        history window [peer band ± exchange_rows] and receives the mirror image. With bands thinner than the
        window a rank talks to more than its two direct neighbours."""
        b0, b1 = self.band(rank)
        w0, w1 = self.clamp(b0 - self.exchange_rows, b1 + self.exchange_rows)
        plan = []
        for peer in range(self.world):
            if peer == rank:
                continue
            p0, p1 = self.band(peer)
            pw0, pw1 = self.clamp(p0 - self.exchange_rows, p1 + self.exchange_rows)
            send = (max(b0, pw0), min(b1, pw1))   # my rows the peer needs
            recv = (max(p0, w0), min(p1, w1))     # the peer's rows I need
            if send[0] < send[1] or recv[0] < recv[1]:
                plan.append((peer, send, recv))
        return plan


def exchange_final(bands: RowBands, rank: int, planes, dist, group=None):
    """Swap the final-reservoir rows between band neighbours. `planes` is a list of tensors shaped (H, ...)
    (row-major, one entry per SoA plane — or one AoS tensor); rows are sent / received in place."""
    ops = []
    for peer, send, recv in bands.exchange_plan(rank):
        for pl in planes:
            if send[0] < send[1]:
                ops.append(dist.P2POp(dist.isend, pl[send[0]:send[1]], peer, group))
            if recv[0] < recv[1]:
                ops.append(dist.P2POp(dist.irecv, pl[recv[0]:recv[1]], peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def gather_rows(bands: RowBands, rank: int, image, dist, dst: int = 0, group=None):
    """readback: rank `dst` receives every band of `image` (shape (H, W, C)) into its own copy"""
    ops = []
    if rank == dst:
        for peer in range(bands.world):
            if peer != dst:
                p0, p1 = bands.band(peer)
                ops.append(dist.P2POp(dist.irecv, image[p0:p1], peer, group))
    else:
        b0, b1 = bands.band(rank)
        ops.append(dist.P2POp(dist.isend, image[b0:b1], dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def combine_samples(radiance, dist, world: int, group=None):
    """sample sharding: mean of the per-GPU frames (one all-reduce = the accumulation step)"""
    dist.all_reduce(radiance, group=group)
    radiance /= world
    return radiance
