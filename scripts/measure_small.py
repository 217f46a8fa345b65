import sys, time
sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds
for w in (worlds.make_bouncer(294), worlds.make_platformer_batch(1), worlds.make_platformer_batch(64)):
    ds = space_for_world(w)
    fl = _abi.FRAME_DOWNLOAD
    for _ in range(20):
        ds.frame(w.margin, fl, grow=True)
    t0 = time.perf_counter()
    K = 200
    for _ in range(K):
        ds.commit(1 << _abi.BUF_POS)
        c = ds.frame(w.margin, fl)
    t = (time.perf_counter() - t0) / K
    ds.timer_start()
    for _ in range(K):
        ds.frame_launch(w.margin, 0)
    ms = ds.timer_stop() / K
    ds.frame_finish()
    print(f"{w.name:24s} N={w.n:6d} e2e(commit+frame+download) {t*1e6:7.1f} us/frame   device-only {ms*1e3:7.1f} us/frame   {c.as_dict()}")
    ds.close()
for w in (worlds.make_bouncer(294), worlds.make_platformer_batch(64)):
    ds = space_for_world(w)
    for _ in range(5):
        ds.frame(w.margin, _abi.FRAME_TIMING, grow=True)
    print(w.name, {k: round(v, 1) for k, v in ds.timing().items()})
    ds.close()
