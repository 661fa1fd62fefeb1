"""The N>1 path on CPU: two processes over gloo run the same partition / gather / assemble bookkeeping bench.py
runs over NCCL.  Each rank produces its interleaved tiles (from the oracle's frame, standing in for the kernel),
rank 0 gathers them with rt_b200.gather_parts and reassembles the frame, which must equal the undivided frame."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, aa, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import harness as H2
        rt = H2.rt_b200
        sc = H2.golden_scene("cornellbox")
        cam = sc.camera(0, w, h)
        full, _ = H2.OracleScene(sc).render(cam, aa, threads=2)
        L = rt.cuda_lib()
        stride = int(L.rt_part_bytes(C.byref(cam), 0, world))
        mine_np = rt.pack_tiles_host(full, rank, world)
        assert mine_np.size == int(L.rt_part_bytes(C.byref(cam), rank, world))
        mine = torch.zeros(stride, dtype=torch.uint8)
        mine[:mine_np.size] = torch.from_numpy(mine_np)
        all_parts = torch.zeros((world, stride), dtype=torch.uint8) if rank == 0 else None
        rt.gather_parts(dist, mine, all_parts, rank)
        # ray-count style reduction used by bench.py
        cnt = torch.tensor([len(rt.part_tile_ids(w, h, rank, world))], dtype=torch.int64)
        dist.all_reduce(cnt)
        if rank == 0:
            got = rt.assemble_tiles_host(all_parts.numpy(), w, h, world)
            tx, ty = rt.tile_grid(w, h, world)
            ok = np.array_equal(got, full) and int(cnt.item()) == tx * ty
            open(result_path, "w").write("ok" if ok else "mismatch")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 100, 70), (2, 64, 64), (3, 33, 95)])
def test_partition_gather_assemble_over_gloo(world, w, h, tmp_path):
    result = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), w, h, 2, result), nprocs=world, join=True)
    assert open(result).read() == "ok"
