"""Small end-to-end run for compute-sanitizer (memcheck): frames on mixed / batched worlds, both line iterators."""
import sys
sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds
for w in (worlds.make_mixed(6000, space=2000, cell=32), worlds.make_platformer_batch(4), worlds.make_bouncer(100)):
    ds = space_for_world(w)
    for f in range(2):
        worlds.advance(w, f + 1)
        ds.set_positions(w.pos)
        c = ds.frame(w.margin, _abi.FRAME_DOWNLOAD | (_abi.FRAME_ALL_CANDIDATES if f else 0) | _abi.FRAME_TESTER_TABLE, grow=True)
    print(w.name, c.as_dict())
    if w.n_spaces == 1:
        rays = worlds.make_rays(3000, 3, w.space_w, w.space_h, 8, 300)
        for fl in (0, _abi.LINE_DDA):
            rc = ds.linetest(rays, flags=_abi.FRAME_DOWNLOAD | fl)
            print("  rays", rc.as_dict())
    ds.close()
print("done")
