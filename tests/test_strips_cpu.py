"""N > 1 host-side logic on CPU: two gloo ranks build their strip plans, pack their halo records (numpy restatement
of k_halo_pack), exchange them with the same torch.distributed code the GPU path uses, and we check with the oracle
that every candidate of every tester homed on a rank is available on that rank (own range or received halo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from resolv_b200 import strips


def test_plan_windows_cover_every_query_and_do_not_skip_ranks():
    H, m, span = 2048 * 4, 1, 1
    plans = [strips.make_plan(r, 4, H, m, span) for r in range(4)]
    assert plans[0].home_lo == 0 and plans[-1].home_hi == H
    for a, b in zip(plans, plans[1:]):
        assert a.home_hi == b.home_lo
        assert a.dn_row_min == b.row_lo and b.up_row_max == a.row_hi - 1
        assert b.row_lo >= a.home_lo and a.row_hi <= b.home_hi   # halos reach only the adjacent strip
    for p in plans:
        assert p.row_lo == max(0, p.home_lo - m) and p.row_hi == min(H, p.home_hi + span + m)
    with pytest.raises(ValueError):
        strips.make_plan(0, 8, 16, 1, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, size, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cap = 4096
        tiles = [strips.tile_world("cfg3", n, t, size=size) for t in range(world)]
        full = strips.concat_worlds(tiles, size * world)
        H = -(-size * world // full.cell_h)
        plan = strips.make_plan(rank, world, H, 1, 1)
        own = np.arange(rank * n, (rank + 1) * n)
        up, dn = strips.pack_reference(full.pos, full.local_aabb, full.cell_h, own, plan.up_row_max, plan.dn_row_min)
        send_up = torch.from_numpy(strips.encode_records(up, full.pos, cap).copy())
        send_dn = torch.from_numpy(strips.encode_records(dn, full.pos, cap).copy())
        recv_up, recv_dn = torch.zeros_like(send_up), torch.zeros_like(send_dn)
        strips.exchange(send_up, send_dn, recv_up, recv_dn, rank, world)
        got = []
        if rank > 0:
            got.append(strips.decode_records(recv_up.numpy())[0])
        if rank + 1 < world:
            got.append(strips.decode_records(recv_dn.numpy())[0])
        out[rank] = dict(sent_up=up, sent_dn=dn, got=np.concatenate(got) if got else np.zeros(0, np.int64),
                         plan=(plan.home_lo, plan.home_hi))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_exchange_a_sufficient_halo():
    from tests.util import OracleBatch
    n, size, world = 4000, 4096, 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n, size, out), nprocs=world, join=True)
        res = {r: out[r] for r in range(world)}
    # what rank 1 received from above is exactly what rank 0 sent down, and vice versa
    assert np.array_equal(np.sort(res[1]["got"]), np.sort(res[0]["sent_dn"]))
    assert np.array_equal(np.sort(res[0]["got"]), np.sort(res[1]["sent_up"]))
    assert 0 < len(res[0]["sent_dn"]) < n // 5
    # sufficiency, checked against the oracle on the whole world
    tiles = [strips.tile_world("cfg3", n, t, size=size) for t in range(world)]
    full = strips.concat_worlds(tiles, size * world)
    ob = OracleBatch(full)
    tot, cands, hits = ob.frame(1)
    sp = ob.spaces[0]
    H = sp.height_cells
    first_row = np.array([min(max(int(sp.cell_rect(i)[1]), 0), H - 1) for i in range(full.n)])
    for r in range(world):
        lo, hi = res[r]["plan"]
        have = np.zeros(full.n, dtype=bool)
        have[r * n:(r + 1) * n] = True
        have[res[r]["got"]] = True
        homed = (first_row[cands["a"]] >= lo) & (first_row[cands["a"]] < hi)
        assert have[cands["a"][homed]].all(), "a tester homed on this rank is not present on it"
        assert have[cands["b"][homed]].all(), "a candidate of a local tester was not delivered by the halo exchange"
    assert tot["candidates"] > 1000
