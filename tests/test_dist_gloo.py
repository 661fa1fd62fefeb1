"""The N>1 path on CPU: two processes over gloo run the same partition / gather / assemble bookkeeping bench.py
runs over NCCL.  Each rank produces its interleaved row bands (from the oracle's frame, standing in for the kernel),
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
        stride = int(L.rt_part_bytes(C.byref(cam), aa, 0, world))
        bh = rt.band_height(cam, aa, world)
        mine_np = rt.pack_bands_host(full, bh, rank, world)
        assert mine_np.size == int(L.rt_part_bytes(C.byref(cam), aa, rank, world))
        mine = torch.zeros(stride, dtype=torch.uint8)
        mine[:mine_np.size] = torch.from_numpy(mine_np)
        all_parts = torch.zeros((world, stride), dtype=torch.uint8) if rank == 0 else None
        rt.gather_parts(dist, mine, all_parts, rank)
        # ray-count style reduction used by bench.py
        cnt = torch.tensor([len(rt.part_band_ids(h, bh, rank, world))], dtype=torch.int64)
        dist.all_reduce(cnt)
        if rank == 0:
            got = rt.assemble_bands_host(all_parts.numpy(), w, h, bh, world)
            ok = np.array_equal(got, full) and int(cnt.item()) == (h + bh - 1) // bh
            open(result_path, "w").write("ok" if ok else "mismatch")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h,aa", [(2, 100, 70, 2), (2, 64, 64, 1), (3, 33, 95, 2), (2, 40, 30, 8)])
def test_partition_gather_assemble_over_gloo(world, w, h, aa, tmp_path):
    result = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), w, h, aa, result), nprocs=world, join=True)
    assert open(result).read() == "ok"
