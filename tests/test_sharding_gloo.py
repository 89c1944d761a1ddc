"""Multi-process host logic of the N>1 path on CPU: world_size 2, backend gloo.

The forward is a stand-in (nearest x`scale` upsample + a per-sample statistic, so that wrong slicing or a
wrong gather order changes the result); the sharding / gather / tile-stitch code is the product code that
`bench.py --gpus N` and the tiled large-frame mode run with backend nccl.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hitsir_b200.sharding import ShardedSR, shard_bounds, stitch_tiles, tile_plan


def fake_sr(x: torch.Tensor, scale: int = 2) -> torch.Tensor:
    y = torch.nn.functional.interpolate(x, scale_factor=scale, mode="nearest")
    return y + x.mean(dim=(1, 2, 3), keepdim=True)          # per-sample global statistic, like casa


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, case: str, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = ShardedSR(fake_sr, 2)
        g = torch.Generator().manual_seed(7)
        if case == "batch_even":
            x = torch.rand(4, 3, 10, 12, generator=g)
            y = sh.forward_batch(x)
            assert torch.equal(y, fake_sr(x))
        elif case == "batch_ragged":
            x = torch.rand(5, 3, 9, 8, generator=g)
            y = sh.forward_batch(x)
            assert y.shape[0] == 5 and torch.equal(y, fake_sr(x))
            lo, hi = shard_bounds(5, world, rank)
            assert torch.equal(sh.forward_batch(x, gather=False), fake_sr(x[lo:hi]))
        elif case == "tiled":
            x = torch.rand(1, 3, 40, 52, generator=g)
            y = sh.forward_tiled(x, tile=24, overlap=8, dst_rank=None)
            origins = tile_plan(40, 52, 24, 8)
            ref = stitch_tiles([fake_sr(x[..., a:a + 24, b:b + 24]) for a, b in origins], origins, 40, 52, 2)
            assert torch.allclose(y, ref, atol=1e-6)
            only0 = sh.forward_tiled(x, tile=24, overlap=8, dst_rank=0)
            assert (only0 is None) == (rank != 0)
        open(os.path.join(out_dir, f"ok_{case}_{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["batch_even", "batch_ragged", "tiled"])
def test_world2_gloo(case, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok_{case}_{r}") for r in range(world))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_tile_plan_matches_kair_scheme():
    # 参考资料/KAIR_master/main_test_swinir.py:268-270 with tile 576, overlap 128 on a 1080x1920 frame
    origins = tile_plan(1080, 1920, 576, 128)
    assert sorted({y for y, _ in origins}) == [0, 448, 504]
    assert sorted({x for _, x in origins}) == [0, 448, 896, 1344]
    # a frame smaller than the tile is one tile
    assert tile_plan(40, 30, 64, 8) == [(0, 0)]


def test_stitch_is_identity_for_consistent_tiles():
    x = torch.rand(2, 3, 20, 28)
    origins = tile_plan(20, 28, 12, 4)
    tiles = [torch.nn.functional.interpolate(x[..., a:a + 12, b:b + 12], scale_factor=2, mode="nearest") for a, b in origins]
    y = stitch_tiles(tiles, origins, 20, 28, 2)
    assert torch.allclose(y, torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest"), atol=1e-6)
