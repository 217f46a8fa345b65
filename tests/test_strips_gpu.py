"""Strip partitioning on ONE GPU: two ranks emulated in one process (two rsv_spaces on cuda:0, the neighbour
exchange replaced by device copies). The union of both ranks' results must equal the single-Space result on the
concatenated world: same candidate counts, same hit set with the same generation ranks, bit-identical MTV/contacts."""
import numpy as np
import pytest

from resolv_b200 import _abi, space_for_world, strips
from tests.util import bits

pytestmark = pytest.mark.gpu


def _hits_table(hits, con, to_global):
    live = (hits["count_rank"] & 0x7f) > 0
    h = hits[live]
    a, b = to_global(h["a"]), to_global(h["b"])
    n = (h["count_rank"] & 0x7f).astype(np.int64)
    rank = (h["count_rank"] >> 8).astype(np.int64)
    cons = [np.concatenate([con["p"][f:f + k], con["n"][f:f + k]], -1) for f, k in zip(h["first_contact"], n)]
    return a, b, n, rank, h["mtv"].copy(), cons


@pytest.mark.parametrize("kind,transport", [("cfg2", "copy"), ("cfg3", "copy"), ("cfg3", "peer"), ("cfg3", "fused"), ("cfg2", "fused")])
def test_two_strips_equal_one_space(kind, transport):
    import torch
    n, size, world = 20000, 9280, 2
    ranks = [strips.StripRank(kind, n, r, world, 0, size=size, cap_halo=8192) for r in range(world)]
    if transport in ("peer", "fused"):
        # the peer-memory transport with both "devices" in one process: mailboxes are connected by pointer (processes
        # on different GPUs map them through CUDA IPC instead); pushes, flag waits and applies are all stream-ordered
        boxes = [sr.mailbox()[0] for sr in ranks]
        ranks[0].connect(None, boxes[1])
        ranks[1].connect(boxes[0], None)
    ref_world = strips.concat_worlds([strips.tile_world(kind, n, t, size=size) for t in range(world)], size * world)
    ref = space_for_world(ref_world)
    try:
        for frame in range(3):
            if frame:   # every rank moves the shapes it owns; the reference moves everything
                ref_world.pos += ref_world.vel
                ref.set_positions(ref_world.pos)
                for r, sr in enumerate(ranks):
                    sr.ds.buf[_abi.BUF_POS][2 * n:4 * n] = ref_world.pos[r * n:(r + 1) * n].reshape(-1)
                    sr.ds.commit(1 << _abi.BUF_POS, n, 2 * n)
            if transport == "fused":
                # rsv_strip_frame_launch: the exchange is part of the frame (pack inside k_bounds, device-side waits);
                # send halves of both ranks first, as above
                for sr in ranks:
                    sr.step_launch(_abi.FRAME_DOWNLOAD, phases=1)
                for sr in ranks:
                    sr.step_launch(_abi.FRAME_DOWNLOAD, phases=2)
            elif transport == "peer":
                for sr in ranks:            # all send halves first: both ranks' streams may share a hardware queue
                    sr.exchange_peer(1)
                for sr in ranks:
                    sr.exchange_peer(2)
            else:
                for sr in ranks:
                    sr.pack()
                torch.cuda.synchronize()
                ranks[1].recv_up.copy_(ranks[0].send_dn)
                ranks[0].recv_dn.copy_(ranks[1].send_up)
                torch.cuda.synchronize()
            tables, P = [], 0
            for sr in ranks:
                if transport == "copy":
                    sr.apply()
                c = sr.ds.frame_finish() if transport == "fused" else sr.ds.frame(1, sr.flags | _abi.FRAME_DOWNLOAD, grow=True)
                P += c.n_candidates
                hits, con = sr.ds.results()
                tables.append(_hits_table(hits, con, sr.global_slot))
            cr = ref.frame(1, _abi.FRAME_DOWNLOAD, grow=True)
            rh, rc = ref.results()
            ra, rb, rn, rrank, rmtv, rcons = _hits_table(rh, rc, lambda x: np.asarray(x, dtype=np.int64))
            assert P == cr.n_candidates
            a = np.concatenate([t[0] for t in tables]); b = np.concatenate([t[1] for t in tables])
            nn = np.concatenate([t[2] for t in tables]); rk = np.concatenate([t[3] for t in tables])
            mtv = np.concatenate([t[4] for t in tables]); cons = sum((t[5] for t in tables), [])
            o, ro = np.lexsort((b, a)), np.lexsort((rb, ra))
            assert np.array_equal(a[o], ra[ro]) and np.array_equal(b[o], rb[ro]), "hit sets differ"
            assert np.array_equal(nn[o], rn[ro]) and np.array_equal(rk[o], rrank[ro]), "contact counts / generation ranks differ"
            assert np.array_equal(bits(mtv[o]), bits(rmtv[ro]))
            dc = np.concatenate([cons[i] for i in o]); rcc = np.concatenate([rcons[i] for i in ro])
            assert np.array_equal(bits(dc), bits(rcc))
            assert len(a) > 1000
            # the halo really is a small boundary layer
            if transport == "copy":
                assert int(ranks[0].send_dn[0].item() & 0xffffffff) < n // 20
    finally:
        for sr in ranks:
            sr.close()
        ref.close()


def test_shapes_that_migrate_into_the_neighbour_strip_are_rehomed_by_the_exchange():
    """"shapes that migrate across a strip boundary are re-homed in the same exchange". A shape is a tester
    where its first cell row is home, whoever owns its slot: owned shapes that wander deep into the neighbour's strip
    travel as halo records and run their IntersectionTest THERE (and only there); the union stays the single Space."""
    import torch
    n, size, world = 12000, 7200, 2
    ranks = [strips.StripRank("cfg3", n, r, world, 0, size=size, cap_halo=8192) for r in range(world)]
    ref_world = strips.concat_worlds([strips.tile_world("cfg3", n, t, size=size) for t in range(world)], size * world)
    ref = space_for_world(ref_world)
    rs = np.random.default_rng(3)
    try:
        movers = [rs.choice(n, 300, replace=False) + r * n for r in range(world)]        # global slots
        ref_world.pos[movers[0], 1] += size * rs.uniform(0.6, 1.2, 300)                     # rank 0's shapes go deep into strip 1
        ref_world.pos[movers[1], 1] -= size * rs.uniform(0.6, 1.2, 300)                     # and the other way round
        ref.set_positions(ref_world.pos)
        for r, sr in enumerate(ranks):
            sr.ds.buf[_abi.BUF_POS][2 * n:4 * n] = ref_world.pos[r * n:(r + 1) * n].reshape(-1)
            sr.ds.commit(1 << _abi.BUF_POS, n, 2 * n)
        for sr in ranks:
            sr.pack()
        torch.cuda.synchronize()
        ranks[1].recv_up.copy_(ranks[0].send_dn)
        ranks[0].recv_dn.copy_(ranks[1].send_up)
        torch.cuda.synchronize()
        tables, P, testers = [], 0, 0
        for sr in ranks:
            sr.apply()
            c = sr.ds.frame(1, sr.flags | _abi.FRAME_DOWNLOAD | _abi.FRAME_ALL_CANDIDATES, grow=True)
            P += c.n_candidates
            testers += c.n_testers
            hits, con = sr.ds.results()
            tables.append((sr.global_slot(hits["a"]), sr.global_slot(hits["b"]), (hits["count_rank"] >> 8).astype(np.int64),
                           (hits["count_rank"] & 0x7f).astype(np.int64), hits["mtv"].copy()))
        cr = ref.frame(1, _abi.FRAME_DOWNLOAD | _abi.FRAME_ALL_CANDIDATES, grow=True)
        rh, _ = ref.results()
        assert P == cr.n_candidates and testers == cr.n_testers == world * n
        a = np.concatenate([t[0] for t in tables]); b = np.concatenate([t[1] for t in tables])
        o, ro = np.lexsort((b, a)), np.lexsort((rh["b"], rh["a"]))
        assert np.array_equal(a[o], rh["a"][ro]) and np.array_equal(b[o], rh["b"][ro]), "candidate sets differ"
        assert np.array_equal(np.concatenate([t[2] for t in tables])[o], (rh["count_rank"][ro] >> 8)), "generation ranks differ"
        assert np.array_equal(np.concatenate([t[3] for t in tables])[o], (rh["count_rank"][ro] & 0x7f))
        assert np.array_equal(bits(np.concatenate([t[4] for t in tables])[o]), bits(rh["mtv"][ro]))
        # every record of a migrated shape was produced by the rank it moved to
        rows = np.floor((ref_world.pos[:, 1] + ref_world.local_aabb[:, 1]) / float(ref_world.cell_h))
        H = ranks[0].plan.H
        home_rank = (np.clip(rows, 0, H - 1) >= ranks[1].plan.home_lo).astype(np.int64)
        assert (home_rank[movers[0]] == 1).sum() > 250 and (home_rank[movers[1]] == 0).sum() > 250
        for r, t in enumerate(tables):
            assert np.all(home_rank[t[0]] == r), "a tester was run away from its home strip"
        assert sum(int(np.isin(t[0], np.concatenate(movers)).sum()) for t in tables) > 100
    finally:
        for sr in ranks:
            sr.close()
        ref.close()


@pytest.mark.parametrize("dda", [False, True])
def test_rays_across_two_strips(dda):
    """Every rank tests all rays against the shapes homed on it (RSV_LINE_HOME_ONLY); the union over ranks must be the
    single-Space LineTest: same (ray, shape) sets, bit-identical MTV / Center / key / contacts."""
    import torch
    from resolv_b200 import worlds
    n, size, world = 20000, 9280, 2
    ranks = [strips.StripRank("cfg3", n, r, world, 0, size=size, cap_halo=8192, max_rays=2000) for r in range(world)]
    ref_world = strips.concat_worlds([strips.tile_world("cfg3", n, t, size=size) for t in range(world)], size * world)
    rays = worlds.make_rays(6000, 77, ref_world.space_w, ref_world.space_h, 16, 1200)
    rays[:2000, 1] = size - 40 + (rays[:2000, 1] % 80)       # a third of the rays start right at the strip boundary
    ref = space_for_world(ref_world, max_rays=len(rays))
    fl = _abi.LINE_DDA if dda else 0
    try:
        for sr in ranks:
            sr.pack()
        torch.cuda.synchronize()
        ranks[1].recv_up.copy_(ranks[0].send_dn)
        ranks[0].recv_dn.copy_(ranks[1].send_up)
        torch.cuda.synchronize()
        got = {}
        for sr in ranks:
            sr.apply()
            sr.ds.frame(1, sr.flags | _abi.FRAME_GRID_ONLY)
            # (the ray capacity of a StripRank's space is whatever its default is: feed the rays in slices)
            step = 2000
            for lo in range(0, len(rays), step):
                hits, con, first, count = sr.linetest(rays[lo:lo + step], flags=fl)
                for h in hits:
                    k, c0 = int(h["count_rank"] & 0xff), int(h["first_contact"])
                    key = (lo + int(h["ray"]), int(h["b"]))
                    assert key not in got, "a (ray, shape) set was produced by both ranks"
                    got[key] = (bits(h["mtv"]).tolist(), bits(h["center"]).tolist(), int(bits(np.array([h["key"]]))[0]),
                                bits(np.concatenate([con["p"][c0:c0 + k], con["n"][c0:c0 + k]], -1)).tolist())
        ref.frame(1, _abi.FRAME_GRID_ONLY)
        ref.linetest(rays, flags=_abi.FRAME_DOWNLOAD | fl)
        rh, rc, _, _ = ref.ray_results()
        want = {}
        for h in rh:
            k, c0 = int(h["count_rank"] & 0xff), int(h["first_contact"])
            want[(int(h["ray"]), int(h["b"]))] = (bits(h["mtv"]).tolist(), bits(h["center"]).tolist(), int(bits(np.array([h["key"]]))[0]),
                                                   bits(np.concatenate([rc["p"][c0:c0 + k], rc["n"][c0:c0 + k]], -1)).tolist())
        assert len(want) > 3000
        assert got == want
    finally:
        for sr in ranks:
            sr.close()
        ref.close()


@pytest.mark.parametrize("transport", ["copy", "fused"])
def test_a_halo_batch_larger_than_its_buffer_fails_the_frame(transport):
    """A boundary population beyond the halo capacity must not pass silently (records past the capacity are dropped by
    the sender): the frame that uses the truncated batch reports RSV_OVF_GHOSTS / RSV_ERR_CAPACITY."""
    import torch
    n, size, world = 20000, 9280, 2
    ranks = [strips.StripRank("cfg2", n, r, world, 0, size=size, cap_halo=64) for r in range(world)]
    try:
        if transport == "fused":
            boxes = [sr.mailbox()[0] for sr in ranks]
            ranks[0].connect(None, boxes[1])
            ranks[1].connect(boxes[0], None)
            for sr in ranks:
                sr.step_launch(0, phases=1)
            for sr in ranks:
                sr.step_launch(0, phases=2)
            for sr in ranks:
                with pytest.raises(_abi.RsvError) as ei:
                    sr.ds.frame_finish()
                assert ei.value.code == _abi.RSV_ERR_CAPACITY
        else:
            for sr in ranks:
                sr.pack()
            torch.cuda.synchronize()
            assert int(ranks[0].send_dn[0].item() & 0xffffffff) > 64          # more boundary shapes than the buffer holds
            ranks[1].recv_up.copy_(ranks[0].send_dn)
            ranks[0].recv_dn.copy_(ranks[1].send_up)
            torch.cuda.synchronize()
            for sr in ranks:
                sr.apply()
                with pytest.raises(_abi.RsvError) as ei:
                    sr.ds.frame(1, sr.flags)
                assert ei.value.code == _abi.RSV_ERR_CAPACITY
    finally:
        for sr in ranks:
            sr.close()
