"""Shared helpers for the parity tests: build oracle Spaces from a worlds.World and canonicalise both sides."""
import numpy as np

from oracle.oracle import OracleSpace


class OracleBatch:
    """One OracleSpace per Space of a (possibly batched) World, with global<->local index maps."""

    def __init__(self, w):
        self.w = w
        self.spaces, self.glob, self.loc = [], [], np.zeros(w.n, dtype=np.int64)
        for s in range(w.n_spaces):
            idx = np.nonzero(w.space_idx == s)[0]
            sp = OracleSpace(w.space_w, w.space_h, w.cell_w, w.cell_h)
            # vertex pool must be re-based per space
            nv = w.nv[idx]
            voff = (np.cumsum(nv) - nv).astype(np.uint32)
            verts = np.concatenate([w.verts[o:o + n] for o, n in zip(w.vert_off[idx], nv)]) if nv.sum() else np.zeros((0, 2))
            sp.add_bulk(w.kind[idx], w.pos0[idx], w.pos[idx], w.radius[idx], voff, nv, verts, w.closed[idx], w.tags[idx])
            self.spaces.append(sp)
            self.glob.append(idx)
            self.loc[idx] = np.arange(len(idx))

    def set_positions(self, pos):
        for sp, idx in zip(self.spaces, self.glob):
            sp.set_positions(pos[idx])

    def frame(self, margin):
        """Returns (counts dict, candidate array (a,b,rank), hits dict) in global indices."""
        w = self.w
        tot = dict(testers=0, candidates=0, hits=0, contacts=0)
        ca, cb, cr, cd = [], [], [], []
        ha, hb, ho, hc, hm, hd, con = [], [], [], [], [], [], []
        for sp, idx in zip(self.spaces, self.glob):
            fc, cand, hits = sp.frame(margin, tester=w.tester[idx], use_any=w.use_any[idx], any_mask=w.any_mask[idx],
                                      none_mask=w.none_mask[idx])
            for k in tot:
                tot[k] += getattr(fc, k)
            ca.append(idx[cand["a"]]); cb.append(idx[cand["b"]]); cr.append(cand["rank"]); cd.append(cand["dist"])
            ha.append(idx[hits.a]); hb.append(idx[hits.b]); ho.append(hits.order); hc.append(hits.count)
            hm.append(hits.mtv); hd.append(hits.dist); con.append(hits.contacts)
        cat = np.concatenate
        cands = dict(a=cat(ca), b=cat(cb), rank=cat(cr), dist=cat(cd))
        hits = dict(a=cat(ha), b=cat(hb), order=cat(ho), count=cat(hc), mtv=cat(hm), dist=cat(hd), contacts=cat(con))
        hits["first"] = np.cumsum(hits["count"]) - hits["count"]
        return tot, cands, hits

    def cells(self):
        """Per global cell (spaces concatenated, row-major): sorted tuple of global shape indices."""
        out = []
        for sp, idx in zip(self.spaces, self.glob):
            counts, entries = sp.cells()
            ends = np.cumsum(counts)
            starts = ends - counts
            for s, e in zip(starts, ends):
                out.append(tuple(sorted(idx[entries[s:e]].tolist())))
        return out


def device_cells(ds, n_entries):
    end, ent = ds.debug_cells(n_entries)
    start = np.concatenate([[0], end[:-1]])
    return [tuple(ent[s:e].tolist()) for s, e in zip(start, end)]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _check_results(c, hits, con, tot, oh, w=None):
    """Hit set bit-exact, contact counts equal, MTV / contacts within 1e-9; returns whether they are bit-identical.

    Contacts: sets of at most 12 points arrive sorted as the reference sorts them. Beyond 12, sort.Slice is Go's
    pdqsort_func, which the HOST layers run on the device's generation-order output (include/resolv_b200.h: rsv_contact);
    here the oracle's restatement of that sort (oracle.go_sort) plays the host, so the comparison is exact for those too."""
    from oracle.oracle import go_sort
    n = hits["count_rank"] & 0x7f
    live = n > 0
    assert c.n_testers == tot["testers"]
    assert c.n_hits == tot["hits"] == int(live.sum()), (c.n_hits, tot["hits"], int(live.sum()))
    assert c.n_contacts == tot["contacts"]
    dh = hits[live]
    dn = n[live]
    # canonical order: by (a, b)
    do = np.lexsort((dh["b"], dh["a"]))
    oo = np.lexsort((oh["b"], oh["a"]))
    assert np.array_equal(dh["a"][do], oh["a"][oo]) and np.array_equal(dh["b"][do], oh["b"][oo]), "hit set mismatch"
    assert np.array_equal(dn[do], oh["count"][oo]), "contact counts mismatch"
    dm = dh["mtv"][do]
    om = oh["mtv"][oo]
    assert np.all(np.abs(dm - om) <= 1e-9), "MTV beyond 1e-9"
    la = None if w is None else np.asarray(w.local_aabb, dtype=np.float64).reshape(-1, 4)

    def host_sort(rows, a, b):
        """What the host layer does with a device set of more than 12 contacts (csrc/host/resolv.hpp: Space::fillSet)."""
        if len(rows) <= 12 or w is None:
            return rows
        if w.kind[a] == 0:
            return rows                                  # circleConvexTest / circleCircleTest do not sort
        if w.kind[b] == 0:
            cx, cy = w.pos[b]                            # convexCircleTest: by distance to circle.position
        else:                                            # convexConvexTest: by distance to convexA.Center()
            mnx, mny, mxx, mxy = la[a, 0] + w.pos[a, 0], la[a, 1] + w.pos[a, 1], la[a, 2] + w.pos[a, 0], la[a, 3] + w.pos[a, 1]
            cx, cy = mnx + (mxx - mnx) * 0.5, mny + (mxy - mny) * 0.5
        dx, dy = rows[:, 0] - cx, rows[:, 1] - cy
        return rows[go_sort(dx * dx + dy * dy)]

    dcon = np.concatenate([host_sort(np.concatenate([con["p"][f:f + k], con["n"][f:f + k]], -1), a, b)
                           for f, k, a, b in zip(dh["first_contact"][do], dn[do], dh["a"][do], dh["b"][do])]) if len(do) else np.zeros((0, 4))
    ocon = np.concatenate([oh["contacts"][f:f + k] for f, k in zip(oh["first"][oo], oh["count"][oo])]) if len(oo) else np.zeros((0, 4))
    if w is None:   # no world at hand: sets beyond 12 points can only be compared as multisets
        def canon(rows, counts):
            out, at = [], 0
            for k in counts:
                r = rows[at:at + k]
                out.append(r[np.lexsort(r.T[::-1])] if k > 12 else r)
                at += k
            return np.concatenate(out) if out else rows
        dcon, ocon = canon(dcon, dn[do]), canon(ocon, oh["count"][oo])
    same_nan = np.isnan(dcon) == np.isnan(ocon)
    assert same_nan.all()
    assert np.all((np.abs(dcon - ocon) <= 1e-9) | np.isnan(ocon)), "contacts beyond 1e-9"
    return bool(np.array_equal(bits(dm), bits(om)) and np.array_equal(bits(dcon), bits(ocon)))


def _records(hits, gated_only=True):
    """(a, b, rank) rows of a frame's records, in buffer order."""
    h = hits[(hits["count_rank"] & 0x80) == 0] if gated_only else hits
    return np.stack([h["a"], h["b"], h["count_rank"] >> 8], -1).astype(np.int64)


def _check_grouping(rec):
    """Records are grouped by tester, a tester's records in generation order (ascending rank)."""
    if not len(rec):
        return
    a = rec[:, 0]
    starts = np.concatenate([[True], a[1:] != a[:-1]])
    assert len(np.unique(a[starts])) == int(starts.sum()), "a tester's records are not contiguous"
    same = ~starts[1:]
    assert np.all(rec[1:, 2][same] > rec[:-1, 2][same]), "a tester's records are not in generation order"


def compare_frame(ds, ob, margin, check_candidates=True, check_rank=True, extra_flags=0):
    """Runs one frame on both sides and asserts: candidate set (with generation rank) bit-exact, hit set bit-exact,
    contacts/MTV within 1e-9 absolute (the north-star tolerance) — and reports whether they are bit-identical.

    The device side runs the frame three ways: the literal walk dumping every candidate (RSV_FRAME_ALL_CANDIDATES),
    the literal walk recording only gated pairs (RSV_FRAME_CANDIDATE_WALK), and the default cell-pair path
    (rsv_cellpairs.cuh), whose records must be the gated records of the dump — same pairs, same ranks."""
    from resolv_b200 import _abi
    tot, cands, oh = ob.frame(margin)
    exact = True
    gated_ref = None
    c_all = None
    if check_candidates:
        c = ds.frame(margin, _abi.FRAME_DOWNLOAD | _abi.FRAME_ALL_CANDIDATES | extra_flags, grow=True)
        hits, con = ds.results()
        assert c.n_candidates == tot["candidates"], (c.n_candidates, tot["candidates"])
        # The generation rank is defined for a Space whose cells are in append order (freshly built). After moves
        # the reference's in-cell order depends on its swap-remove history (cell.go:26-38), so against a moved
        # oracle only the candidate SET is compared (check_rank=False).
        rank = hits["count_rank"] >> 8
        dev = np.stack([hits["a"], hits["b"], rank if check_rank else 0 * rank], -1).astype(np.int64)
        ora = np.stack([cands["a"], cands["b"], cands["rank"] if check_rank else 0 * cands["rank"]], -1).astype(np.int64)
        dev = dev[np.lexsort((dev[:, 1], dev[:, 0]))]
        ora = ora[np.lexsort((ora[:, 1], ora[:, 0]))]
        assert dev.shape == ora.shape and np.array_equal(dev, ora), "candidate set / generation rank mismatch"
        exact &= _check_results(c, hits, con, tot, oh, ob.w)
        gated_ref = _records(hits)
        gated_ref = gated_ref[np.lexsort((gated_ref[:, 1], gated_ref[:, 0]))]
        c_all = c
        # the literal walk without the dump: float-gated records, the exact gate applied per record afterwards
        c = ds.frame(margin, _abi.FRAME_DOWNLOAD | _abi.FRAME_CANDIDATE_WALK | extra_flags, grow=True)
        hits, con = ds.results()
        assert c.n_candidates == tot["candidates"], (c.n_candidates, tot["candidates"])
        exact &= _check_results(c, hits, con, tot, oh, ob.w)
        rec = _records(hits)
        _check_grouping(rec)
        assert np.array_equal(rec[np.lexsort((rec[:, 1], rec[:, 0]))], gated_ref), "walk records != gated candidates of the dump"
    # the default frame
    c = ds.frame(margin, _abi.FRAME_DOWNLOAD | extra_flags, grow=True)
    hits, con = ds.results()
    exact &= _check_results(c, hits, con, tot, oh, ob.w)
    rec = _records(hits, gated_only=False)
    _check_grouping(rec)
    srt = rec[np.lexsort((rec[:, 1], rec[:, 0]))]
    if gated_ref is not None:
        if not (extra_flags & _abi.FRAME_INCREMENTAL):
            assert c.n_gated == len(gated_ref) and not np.any(hits["count_rank"] & 0x80), "cell-pair records are exactly the gated pairs"
        srt = srt[(hits["count_rank"][np.lexsort((rec[:, 1], rec[:, 0]))] & 0x80) == 0]
        assert np.array_equal(srt, gated_ref), "default-frame records != gated candidates of the dump (pairs / ranks)"
    elif check_rank:
        # every record is one of the oracle's candidates, with the oracle's generation rank
        key = lambda a, b: a.astype(np.int64) * (1 << 32) + b.astype(np.int64)
        ok = np.argsort(key(cands["a"], cands["b"]))
        pos = np.searchsorted(key(cands["a"], cands["b"])[ok], key(srt[:, 0], srt[:, 1]))
        assert np.all(pos < len(ok))
        assert np.array_equal(cands["rank"][ok][pos], srt[:, 2]) and np.array_equal(cands["b"][ok][pos], srt[:, 1])
    return (c_all if c_all is not None else c), exact
