"""Quick independent check of the candidate walk against the CPU oracle (no torch import): candidate sets with
generation ranks, hit sets, contacts and MTVs on a few small worlds, with and without the unfiltered fast paths."""
import sys
import time

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402
from tests.util import OracleBatch, compare_frame  # noqa: E402

t0 = time.time()
for name, w, margins in (("mixed filtered 4096", worlds.make_mixed(4096, space=4096, cell=32), (1, 2)),
                         ("rects 20000", worlds.make_rects(20000, space=9280, cell=32), (1,)),
                         ("bouncer", worlds.make_bouncer(), (1,)),
                         ("platformer batch 8", worlds.make_platformer_batch(8), (4,))):
    ds = space_for_world(w, max_pairs=64 * w.n, max_contacts=16 * w.n)
    ob = OracleBatch(w)
    unfiltered = not (w.use_any.any() or w.none_mask.any())
    for m in margins:
        c, exact = compare_frame(ds, ob, m)                                           # every candidate, with ranks
        c2, exact2 = compare_frame(ds, ob, m, check_candidates=False)                 # the production walk
        line = f"{name:22s} margin {m}: P={c.n_candidates} H={c.n_hits} bit_exact={exact and exact2}"
        if unfiltered:
            for fl in (_abi.FRAME_NO_TAG_FILTERS, _abi.FRAME_NO_TAG_FILTERS | _abi.FRAME_OWN_RECT):
                c3, exact3 = compare_frame(ds, ob, m, check_candidates=False, extra_flags=fl)
                line += f" flags={fl:#x}:{exact3}"
        print(line, flush=True)
    ds.close()
print(f"ORACLE CHECK OK ({time.time() - t0:.1f} s)")
