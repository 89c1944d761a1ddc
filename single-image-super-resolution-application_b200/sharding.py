"""Data-parallel plumbing around the forward pass: one process per GPU.

The reference is single-process / single-device (configs/model_config.py:114); images of a batch
are independent (casa / UnionAttention statistics are per sample), so a batch shards into
contiguous slices with NO collective on the data path and ONE all-gather of the SR outputs over
NCCL/NVLink at the end.  A single large frame is processed as overlapping tiles with the tiling
and overlap-add stitching of the only tiled-inference precedent in the reference tree
(参考资料/KAIR_master/main_test_swinir.py:256-285): tile t of the list goes to rank t % world.

Everything here is host logic + torch.distributed calls; it works with backend "gloo" on CPU
tensors (used by the world_size-2 CPU tests with a stand-in forward) and "nccl" on GPUs.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n items for `rank`; the first n % world ranks get one extra."""
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def tile_plan(h: int, w: int, tile: int, overlap: int) -> List[Tuple[int, int]]:
    """Top-left corners of the tiles, KAIR scheme (main_test_swinir.py:268-270):
    stride = tile - overlap; origins range(0, dim - tile, stride) + [dim - tile]."""
    tile_h, tile_w = min(tile, h), min(tile, w)
    stride_h, stride_w = max(tile_h - overlap, 1), max(tile_w - overlap, 1)
    hs = list(range(0, h - tile_h, stride_h)) + [h - tile_h]
    ws = list(range(0, w - tile_w, stride_w)) + [w - tile_w]
    return [(y, x) for y in hs for x in ws]


def stitch_tiles(tiles: Sequence[torch.Tensor], origins: Sequence[Tuple[int, int]], h: int, w: int, scale: int) -> torch.Tensor:
    """Overlap-add and divide by the coverage count (main_test_swinir.py:271-283)."""
    b, c = tiles[0].shape[:2]
    E = torch.zeros(b, c, h * scale, w * scale, dtype=tiles[0].dtype, device=tiles[0].device)
    Wt = torch.zeros_like(E)
    for t, (y, x) in zip(tiles, origins):
        th, tw = t.shape[2], t.shape[3]
        E[..., y * scale:y * scale + th, x * scale:x * scale + tw].add_(t)
        Wt[..., y * scale:y * scale + th, x * scale:x * scale + tw].add_(1.0)
    return E.div_(Wt)


class ShardedSR:
    """Run `forward` (a HiT_SIR module or any callable (b,c,h,w)->(b,c,sh,sw)) data-parallel over the
    default process group.  Weights are replicated (10.2 M parameters); nothing is exchanged until the
    output gather."""

    def __init__(self, forward: Callable[[torch.Tensor], torch.Tensor], scale: int, group=None):
        self.forward = forward
        self.scale = scale
        self.group = group

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if dist.is_initialized() else 0

    def forward_local(self, x_local: torch.Tensor) -> torch.Tensor:
        return self.forward(x_local)

    def forward_batch(self, x: torch.Tensor, gather: bool = True) -> torch.Tensor:
        """x: the GLOBAL batch (same tensor on every rank).  Returns the global SR batch on every rank
        (gather=True) or this rank's slice."""
        B = x.shape[0]
        lo, hi = shard_bounds(B, self.world, self.rank)
        y_local = self.forward(x[lo:hi]) if hi > lo else x.new_zeros((0, x.shape[1], x.shape[2] * self.scale, x.shape[3] * self.scale))
        if not gather or self.world == 1:
            return y_local
        return self.gather_batch(y_local, B)

    def gather_batch(self, y_local: torch.Tensor, B: int) -> torch.Tensor:
        world = self.world
        shape = (B,) + tuple(y_local.shape[1:])
        if B % world == 0:
            out = torch.empty(shape, dtype=y_local.dtype, device=y_local.device)
            dist.all_gather_into_tensor(out, y_local.contiguous(), group=self.group)
            return out
        # ragged: pad every slice to the largest one
        per = -(-B // world)
        pad = torch.zeros((per,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
        pad[:y_local.shape[0]] = y_local
        buf = torch.empty((world * per,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(buf, pad, group=self.group)
        parts = []
        for r in range(world):
            lo, hi = shard_bounds(B, world, r)
            parts.append(buf[r * per:r * per + (hi - lo)])
        return torch.cat(parts, 0)

    def forward_tiled(self, x: torch.Tensor, tile: int, overlap: int, dst_rank: Optional[int] = 0) -> Optional[torch.Tensor]:
        """One (or a few) large frame(s) x (b,c,H,W): tiles are dealt round-robin to the ranks, SR tiles are
        all-gathered and stitched.  Returns the stitched frame on `dst_rank` (None elsewhere), or on every
        rank when dst_rank is None.  Tiled output == reference(tile) per tile + KAIR stitching; it is NOT
        the full-frame forward (global casa/Fusion statistics couple every pixel)."""
        b, c, H, W = x.shape
        origins = tile_plan(H, W, tile, overlap)
        th, tw = min(tile, H), min(tile, W)
        world, rank = self.world, self.rank
        mine = [i for i in range(len(origins)) if i % world == rank]
        outs = []
        for i in mine:
            y0, x0 = origins[i]
            outs.append(self.forward(x[..., y0:y0 + th, x0:x0 + tw].contiguous()))
        per = -(-len(origins) // world)
        s = self.scale
        local = torch.zeros((per, b, c, th * s, tw * s), dtype=torch.float32, device=x.device)
        for k, o in enumerate(outs):
            local[k] = o
        if world > 1 and dst_rank is None:
            buf = torch.empty((world * per, b, c, th * s, tw * s), dtype=torch.float32, device=x.device)
            dist.all_gather_into_tensor(buf, local, group=self.group)
        elif world > 1:
            # only `dst_rank` stitches: a GATHER moves 1/world of what an all-gather would put on every link
            buf = torch.empty((world * per, b, c, th * s, tw * s), dtype=torch.float32, device=x.device) if rank == dst_rank else None
            dist.gather(local, list(buf.split(per)) if rank == dst_rank else None, dst=dst_rank, group=self.group)
        else:
            buf = local
        if dst_rank is not None and rank != dst_rank:
            return None
        tiles = [buf[(i % world) * per + i // world] for i in range(len(origins))]
        return stitch_tiles(tiles, origins, H, W, s)


class PeerGather:
    """Gather / all-gather of equal per-rank slices by peer-to-peer copies over NVLink instead of an NCCL kernel.

    mode="allgather": every rank ends up with every slice.  mode="gather": only rank `dst` does -- each rank writes its slice into row
    `rank` of dst's buffer and nothing else moves, 1/world of the all-gather's traffic per GPU (what "reassemble the outputs" needs).

    Every rank owns `slots` symmetric output buffers [world, *slice] (torch symmetric memory: the same allocation is mapped into every
    peer's address space).  `start(y, slot)` enqueues, on a side stream, a device-side barrier (all ranks have released the slot), one
    copy of this rank's slice into row `rank` of every rank's buffer, and a second barrier (every peer's row has landed here); the
    caller's stream only waits for it in `wait(slot)`.  A gather therefore overlaps the next forward pass -- the copies are plain
    device-to-device transfers that need no SMs of the persistent compute kernels, whereas an NCCL all-gather kernel competes with them
    for SMs -- which is what keeps 8-GPU weak scaling at the 1-GPU step time.  NVLink/NVSwitch only (same node); construct it
    collectively on every rank of the group."""

    def __init__(self, slice_shape, dtype, device, group=None, slots: int = 2, mode: str = "allgather", dst: int = 0):
        assert mode in ("allgather", "gather")
        self.mode, self.dst = mode, dst
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        name = self.group.group_name
        try:                                                    # older torch releases want the group opted in explicitly
            if not symm_mem.is_symm_mem_enabled_for_group(name):
                symm_mem.enable_symm_mem_for_group(name)
        except Exception:
            pass
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.slice_shape = tuple(slice_shape)
        self.full_shape = (self.world,) + self.slice_shape
        self.dtype = dtype
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.bufs, self.hdls, self.peers = [], [], []
        for _ in range(slots):
            buf = symm_mem.empty(self.full_shape, dtype=dtype, device=self.device)
            hdl = symm_mem.rendezvous(buf, self.group)
            self.bufs.append(buf)
            self.hdls.append(hdl)
            self.peers.append([hdl.get_buffer(p, self.full_shape, dtype) for p in range(self.world)])

    def start(self, y_local: torch.Tensor, slot: int) -> None:
        assert tuple(y_local.shape) == self.slice_shape and y_local.dtype == self.dtype
        y_local = y_local.contiguous()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))      # y_local is ready; this rank no longer reads the slot
        with torch.cuda.stream(self.stream):
            self.hdls[slot].barrier(channel=0)
            if self.mode == "gather":
                self.peers[slot][self.dst][self.rank].copy_(y_local, non_blocking=True)
            else:
                for k in range(self.world):                                   # start with the neighbour: spreads the load over the switch
                    p = (self.rank + k) % self.world
                    self.peers[slot][p][self.rank].copy_(y_local, non_blocking=True)
            self.hdls[slot].barrier(channel=1)
        y_local.record_stream(self.stream)

    def wait(self, slot: int) -> torch.Tensor:
        """Makes the current stream wait for the gather started on `slot`; returns the [world, *slice] buffer (valid until the slot's
        next `start`; in "gather" mode only rank `dst` holds the data)."""
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return self.bufs[slot]
