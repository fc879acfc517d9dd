"""Host logic of the context-parallel layer on CPU with the gloo backend (world_size 2): token slicing, the output
all-gather, the IPC-handle exchange and the VAE unit (chunk / tile) ownership + exchange.  No GPU, no compute kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ltx2_b200 import context_parallel as cp
        B, N, C = 2, 12, 5
        g = torch.Generator().manual_seed(0)
        lat = torch.randn(B, N, C, generator=g)
        pos = torch.randn(B, 3, N, 2, generator=g)
        ts_scalar = torch.rand(B, 1, generator=g)
        ts_token = torch.rand(B, N, generator=g)
        a, b = cp.token_range(N, rank, world)
        assert (a, b) == (rank * N // world, (rank + 1) * N // world)
        l, t1, p = cp.slice_tokens(lat, ts_scalar, pos, rank, world)
        assert torch.equal(l, lat[:, a:b]) and torch.equal(p, pos[:, :, a:b]) and torch.equal(t1, ts_scalar)
        _, t2, _ = cp.slice_tokens(lat, ts_token, pos, rank, world)
        assert torch.equal(t2, ts_token[:, a:b]) and l.is_contiguous() and p.is_contiguous()
        # a token-local "model": the gathered result must equal the unsharded computation, on every rank
        full = cp.gather_tokens(l * 2.0 + 1.0)
        assert torch.equal(full, lat * 2.0 + 1.0)
        handles = cp.exchange_handles(bytes([rank]) * 64)
        assert handles == b"".join(bytes([r]) * 64 for r in range(world))
        with pytest.raises(ValueError):
            cp.token_range(13, rank, world)
        # VAE decode units (chunks / tiles) are owned round-robin and exchanged to all ranks or to one rank
        from ltx2_b200 import video_vae as vv
        assert [vv.unit_owner(i, world) for i in range(5)] == [0, 1, 0, 1, 0]
        G = dist.group.WORLD
        for i in range(3):
            owner = vv.unit_owner(i, world)
            unit = torch.full((2, 3), float(10 + i)) if rank == owner else torch.zeros(2, 3)
            vv._exchange_unit(unit, owner, G, None)
            assert torch.equal(unit, torch.full((2, 3), float(10 + i)))
            unit = torch.full((2, 3), float(20 + i)) if rank == owner else torch.zeros(2, 3)
            vv._exchange_unit(unit, owner, G, 0)
            if rank == 0 or rank == owner:
                assert torch.equal(unit, torch.full((2, 3), float(20 + i)))
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_cp_host_logic_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
