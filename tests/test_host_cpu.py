"""Host-layer logic that needs no GPU: the C++ port of Go's sort.Slice (pdqsort_func) in resolv.hpp must produce the
same permutation as the oracle's restatement (and hence as oracle/second_opinion.py's, which is checked against the
oracle elsewhere) — including for equal keys beyond 12 elements, where the algorithm is not stable."""
import os
import subprocess

import numpy as np

from oracle.oracle import go_sort

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gosort_port_matches_the_oracle(tmp_path):
    exe = str(tmp_path / "test_gosort")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I/usr/local/cuda/include", os.path.join(ROOT, "tests/host/test_gosort.cpp"), "-o", exe])
    rs = np.random.default_rng(7)
    arrays = []
    for n in list(range(0, 40)) + [49, 50, 51, 63, 64, 65, 100, 127, 128, 200, 333, 400]:
        for kind in range(6):
            if kind == 0:
                a = rs.integers(0, max(n // 3, 2), n).astype(np.float64)        # many exact ties
            elif kind == 1:
                a = rs.random(n)
            elif kind == 2:
                a = np.sort(rs.integers(0, 10, n)).astype(np.float64)            # already sorted, with ties
            elif kind == 3:
                a = np.sort(rs.random(n))[::-1].copy()                            # strictly decreasing
            elif kind == 4:
                a = np.zeros(n)                                                   # all equal
            else:
                a = np.concatenate([np.arange(n // 2), np.arange(n - n // 2)]).astype(np.float64)   # two runs
            arrays.append(a)
    text = "".join(f"{len(a)} " + " ".join(repr(float(x)) for x in a) + "\n" for a in arrays)
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(out) == len(arrays)
    unstable = 0
    for a, ln in zip(arrays, out):
        got = np.array([int(x) for x in ln.split()], dtype=np.int64)
        want = go_sort(a).astype(np.int64)
        assert np.array_equal(got, want), (len(a), a[:16], got[:16], want[:16])
        if len(a) and not np.array_equal(want, np.argsort(a, kind="stable")):
            unstable += 1
    assert unstable > 10      # the cases really exercise the non-stable regime
