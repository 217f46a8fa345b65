"""The closed-form identities behind RSV_FRAME_OWN_RECT (rsv_pairs.cuh: k_pref_*, k_pairs_own), checked on the CPU
against the oracle's candidate lists: candidate count of a tester and generation rank of every shape met in the
tester's own cell rect, from three prefix counts over the entry flags (scripts/proto_pairs_v2.py is the numpy
prototype of the kernel's arithmetic). Walk semantics: shape.go:220-237, cell.go:86-121."""
import importlib.util
import os

import numpy as np
import pytest

from resolv_b200 import worlds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("proto_pairs_v2", os.path.join(ROOT, "scripts", "proto_pairs_v2.py"))
proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(proto)


@pytest.mark.parametrize("name,margin", [("dense", 1), ("dense", 3), ("big", 0), ("big", 2), ("bouncer", 1)])
def test_candidate_counts_and_ranks_from_prefix_counts(name, margin):
    if name == "dense":
        w = worlds.make_rects(800, space=768, cell=32)
    elif name == "big":
        w = worlds.make_rects(250, space=512, cell=16, lo=20.0, hi=90.0)
    else:
        w = worlds.make_bouncer()
    w.use_any[:] = 0
    w.none_mask[:] = 0
    w.pos[:8] = np.array([[-5.0, 40.0], [3.0, -2.0], [w.space_w + 4.0, 100.0], [w.space_w - 3.0, w.space_h - 2.0]] * 2)
    checked, own, grown = proto.check(w, margin)       # asserts every count and every rank against the oracle
    assert checked > 300 and own <= grown
