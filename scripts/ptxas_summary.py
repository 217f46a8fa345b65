"""Registers / spills / shared memory per kernel from a `make -B` log of resolv_b200/csrc (ptxas -v)."""
import re
import subprocess
import sys
log = subprocess.run(["make", "-C", "resolv_b200/csrc"] + sys.argv[1:], capture_output=True, text=True)
t = log.stdout + log.stderr
if "error" in t:
    print("\n".join(l for l in t.splitlines() if "error" in l or "rror:" in l)[:3000])
pat = re.compile(r"Compiling entry function '([^']+)'[^\n]*\n[^\n]*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n[^\n]*Used (\d+) registers[^\n]*?(?:, (\d+) bytes smem)?\n")
want = sys.argv[1:] and [a for a in sys.argv[1:] if not a.startswith("-")]
for m in pat.finditer(t):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"{name:45s} regs {m.group(5):>3s} spill {m.group(3)}/{m.group(4)} smem {m.group(6) or 0}")
