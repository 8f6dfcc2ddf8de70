"""World-size-2 test of the multi-GPU plumbing on CPU (gloo): sample sharding + the one integer reduce.
The CUDA kernel cannot run here, so each rank's shard is rendered by the oracle on the renderer's own Philox stream
(keyed on the GLOBAL sample index), converted to the int64 fixed-point accumulation format, and combined with
reduce_accum() exactly as bench.py does with NCCL.  The union must equal the unsharded render."""
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
W, H, SPP, DEPTH, SEED = 48, 32, 8, 20, 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, out_path):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    import torch
    import torch.distributed as dist
    import oracle
    rtw = importlib.import_module("raytracing-one-weekend_b200")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = rtw.sample_shard(SPP, rank, world)
    p = oracle.port()
    s, _, rays = p.render_philox(p.scene_cover(), W, H, b, e, DEPTH, seed=SEED, nthreads=1)
    fx = torch.zeros((H, W, 4), dtype=torch.int64)
    fx[..., :3] = torch.from_numpy(np.rint(s * rtw.FIXED_POINT_ONE).astype(np.int64))
    fx[..., 3] = e - b
    rtw.reduce_accum(fx, dst=0)
    if rank == 0:
        np.save(out_path, fx.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_spp_shards_reduce_to_the_unsharded_image(rtw, port, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "reduced.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    full, _, _ = port.render_philox(port.scene_cover(), W, H, 0, SPP, DEPTH, seed=SEED, nthreads=2)
    want = np.rint(full * rtw.FIXED_POINT_ONE).astype(np.int64)
    assert np.all(got[..., 3] == SPP)
    assert np.abs(got[..., :3] - want).max() <= 2  # one rounding per shard, 2^-32 units
    # and a single-process "world" is a no-op
    import torch
    x = torch.ones(4, dtype=torch.int64)
    assert rtw.reduce_accum(x) is x


def _rows_worker(rank, world, port_no, full_path, out_path, tile_rows):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    import torch
    import torch.distributed as dist
    rtw = importlib.import_module("raytracing-one-weekend_b200")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.from_numpy(np.load(full_path))
    h = full.shape[0]
    lr = rtw.row_tile_local_rows(h, tile_rows, world)
    local = torch.zeros((lr, *full.shape[1:]), dtype=full.dtype)
    # what the kernel does for this rank: global row r lives in tile r // tile_rows; tiles rank, rank + world, ... are ours
    for r in range(h):
        t, w = divmod(r, tile_rows)
        if t % world == rank:
            local[(t // world) * tile_rows + w] = full[r]
    g = rtw.gather_row_tiles(local, dst=0)
    if rank == 0:
        np.save(out_path, rtw.untile_rows(g, h, tile_rows, world).numpy())
    else:
        assert g is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("tile_rows", [1, 5, 8])
def test_row_tiles_gather_to_the_unsplit_image(rtw, tmp_path, tile_rows):
    """The row-tile alternative of SURVEY 8(e) on CPU (gloo, world 2): packed per-rank buffers, one gather, tiles put back."""
    import torch.multiprocessing as mp
    rng = np.random.default_rng(tile_rows)
    full = rng.integers(0, 1 << 40, size=(H + 1, W, 4), dtype=np.int64)   # 33 rows: ragged against every tile size
    fp, out = str(tmp_path / "full.npy"), str(tmp_path / "rows.npy")
    np.save(fp, full)
    mp.spawn(_rows_worker, args=(2, _free_port(), fp, out, tile_rows), nprocs=2, join=True)
    assert np.array_equal(np.load(out), full)
    assert rtw.row_tile_local_rows(33, tile_rows, 2) == rtw.lib().rtw_row_tile_local_rows(33, tile_rows, 2)
    import torch
    x = torch.arange(24, dtype=torch.int64).reshape(6, 2, 2)
    assert torch.equal(rtw.untile_rows(rtw.gather_row_tiles(x), 6, 2, 1), x)   # world 1: identity
