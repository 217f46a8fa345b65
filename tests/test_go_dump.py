"""Pinning kit for the oracle: text worlds + the dumps the REAL reference must produce.

The Go harness (parity/go) cannot run in this image (no Go toolchain). What runs here: the world files round-trip, the
committed expected dumps are what the oracle produces today, and - when somebody has dropped reference dumps into
parity/dumps/ (`go run ./dump ../worlds/X.world > ../dumps/X.dump`) - they are compared line by line, which is the test
that turns "parity unpinned" into "pinned".
"""
import glob
import os

import numpy as np
import pytest

from parity import worldfile

PARITY = os.path.dirname(os.path.abspath(worldfile.__file__))
WORLDS = sorted(p for p in glob.glob(os.path.join(PARITY, "worlds", "*.world")) if "bench_" not in os.path.basename(p))


def test_world_files_are_committed():
    assert len(WORLDS) >= 4


def test_world_file_roundtrip(tmp_path):
    w, pos, rays = worldfile.cases()[2]      # the mixed, tag-filtered world
    p = tmp_path / "w.world"
    worldfile.write_world(p, w, pos, rays)
    w2, pos2, rays2 = worldfile.read_world(p)
    for f in ("kind", "pos", "pos0", "radius", "tags", "nv", "vert_off", "verts", "closed", "tester", "use_any", "any_mask",
              "none_mask", "local_aabb"):
        a, b = getattr(w, f), getattr(w2, f)
        assert a.shape == b.shape and a.tobytes() == np.asarray(b, dtype=a.dtype).tobytes(), f
    assert (w2.space_w, w2.space_h, w2.cell_w, w2.cell_h, w2.margin, w2.name) == (w.space_w, w.space_h, w.cell_w, w.cell_h, w.margin, w.name)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(pos, pos2)) and rays.tobytes() == rays2.tobytes()


@pytest.mark.parametrize("path", WORLDS, ids=[os.path.basename(p) for p in WORLDS])
def test_expected_dump_is_what_the_oracle_produces(path):
    w, pos, rays = worldfile.read_world(path)
    with open(os.path.join(PARITY, "expected", w.name + ".dump")) as fh:
        expected = fh.read()
    assert worldfile.oracle_dump(w, pos, rays) == expected, "oracle drifted from parity/expected (python parity/worldfile.py export)"


def test_reference_dumps_match_the_oracle():
    dumps = sorted(glob.glob(os.path.join(PARITY, "dumps", "*.dump")))
    if not dumps:
        pytest.skip("no reference dumps in parity/dumps/ (needs Go: see parity/README.md); the oracle stays 'parity unpinned'")
    for d in dumps:
        with open(d) as fh:
            got = fh.read().splitlines()
        with open(os.path.join(PARITY, "expected", os.path.basename(d))) as fh:
            want = fh.read().splitlines()
        for i, (g, x) in enumerate(zip(got, want)):
            assert g == x, f"{os.path.basename(d)} line {i + 1}: reference says {g[:160]!r}, oracle says {x[:160]!r}"
        assert len(got) == len(want)
