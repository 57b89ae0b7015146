"""SASS opcode histogram of the built library (no GPU needed): python scripts/sass_histogram.py > profiles/r2_sass_opcodes.txt
Per kernel family: the Blackwell-specific / asynchronous-memory opcodes that show the hand-written tcgen05 / TMEM / TMA path."""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "sahs-deformable-nerf_b200", "lib", "libsahs_b200.so")
SPECIAL = re.compile(r"^(UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|ELECT|REDG|ATOMG|NANOSLEEP|CCTL|FENCE|UTCBAR|LDGSTS)")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
per, allops, fn = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        fam = next((k for k in ("field_fwd_duo", "field_fwd", "field_bwd", "field_wgrad", "spade_conv", "sample_pdf_merge64",
                                "sample_pdf_merge", "composite_fwd", "composite_bwd") if k in name), "other")
        fn = fam
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        op = m.group(1)
        allops[op] += 1
        if SPECIAL.match(op):
            per[fn][op] += 1
print(f"SASS opcode histogram of sahs-deformable-nerf_b200/lib/libsahs_b200.so (cuobjdump -sass, sm_100a; scripts/sass_histogram.py)\n")
for fam in sorted(per):
    print(f"{fam}: " + ", ".join(f"{op} x{n}" for op, n in per[fam].most_common()))
print("\nall opcodes (top 30):")
for op, n in allops.most_common(30):
    print(f"{n:8d}  {op}")
