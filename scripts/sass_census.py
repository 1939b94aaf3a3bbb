"""SASS opcode census of libupp_geom.so (no GPU needed): which Blackwell-specific instructions the kernels compile to.
    python scripts/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "iccv2025-upp_b200", "lib", "libupp_geom.so")
WATCH = ["UBLKCP", "UTMALDG", "UTMAPF", "SYNCS", "FFMA2", "FADD2", "FMUL2", "CREDUX", "REDUX", "MATCH", "VOTE", "FMNMX3", "VIMNMX3",
         "ST.ASYNC", "STAS", "UCGABAR", "MEMBAR", "ERRBAR", "REDG", "ATOMG", "ATOMS", "RED.E", "BAR.SYNC", "SHFL", "LDS", "STS",
         "LDG", "STG", "NANOSLEEP", "UTC", "LDTM", "HMMA", "HGMMA", "LDL", "STL"]

out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        per[cur][m.group(1)] += 1

tot = collections.Counter()
for c in per.values():
    tot.update(c)
print(f"libupp_geom.so: {len(per)} kernels, {sum(tot.values())} SASS instructions (cuobjdump -sass, sm_100a)\n")
print("library totals of the opcodes that carry the design (prefix match):")
for w in WATCH:
    n = sum(v for k, v in tot.items() if k.startswith(w))
    print(f"  {w:10s} {n}")
print("\n(no UTC*MMA / LDTM / HMMA: K = 3 distances are not a tensor-core contraction -- BASELINE.json north_star)\n")
print("per kernel: instructions | TMA bulk (UBLKCP) / tensor (UTMALDG) | packed fp32x2 | CREDUX+REDUX | MATCH | atomics (REDG/ATOMG/RED) | ST.ASYNC-class | LDL+STL")
for name, c in per.items():
    g = lambda p: sum(v for k, v in c.items() if k.startswith(p))  # noqa: E731
    print(f"  {name[:110]:110s} {sum(c.values()):6d} | {g('UBLKCP'):2d}/{g('UTMALDG'):2d} | {g('FFMA2') + g('FADD2') + g('FMUL2'):4d} | "
          f"{g('CREDUX') + g('REDUX'):3d} | {g('MATCH'):2d} | {g('REDG') + g('ATOMG') + g('RED.'):2d} | {g('ST.ASYNC') + g('STAS'):2d} | {g('LDL') + g('STL'):2d}")
