#!/usr/bin/env python3
"""Build-time guard for the arithmetic contract.

The parity-critical kernels must not contain fused multiply-adds on f32 data:
ptxas contracts mul+add (even mul.rn.f32x2 + add.rn.f32x2) unless prevented.
Scans the SASS of libhnsw_b200.so and fails if a scalar FFMA (outside the IEEE division /
square-root expansions) or a packed multiply FMUL2 appears in any kernel whose name matches
the distance / search / build / brute-force families.  FFMA2 is allowed: the sources never
emit a packed multiply (so none can be contracted with a following add); every FFMA2 comes
from an explicit fma.rn.f32x2 that restates ONE rounded product exactly (csrc/dist.cuh:
fma(2^23+c, delta, -2^23*delta) == rn(c*delta), fma(t, t, -0.0) == rn(t*t)).
Prints the per-kernel instruction mix as evidence.
"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "hnsw_rs_b200/libhnsw_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
critical = re.compile(r"search_kernel|build_kernel|bf_chunk|dist_query_many|dist_pairs|dist_one_to_many|dist_full|quantise_kernel")
cur = None
mix = collections.defaultdict(collections.Counter)
seq = collections.defaultdict(list)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        mix[cur][m.group(1).split(".")[0]] += 1
        seq[cur].append(m.group(1))
bad = 0
for fn, c in sorted(mix.items()):
    if not critical.search(fn):
        continue
    ops = seq[fn]
    # FFMAs that belong to the IEEE division / square-root expansions (__fdiv_rn, __fsqrt_rn:
    # FCHK / MUFU.RCP / MUFU.RSQ followed by Newton steps) are part of a correctly rounded
    # single operation; anything else would be a contracted a*b+c.
    stray = 0
    # code after the kernel's last EXIT is the out-of-line slow path of those expansions
    last_exit = max([i for i, o in enumerate(ops) if o.startswith("EXIT")] or [len(ops)])
    for i, op in enumerate(ops):
        if i > last_exit:
            break
        if op.split(".")[0] == "FFMA":
            near = ops[max(0, i - 28):i + 6]
            if not any(o.startswith(("MUFU.RCP", "MUFU.RSQ", "FCHK")) for o in near):
                stray += 1
    print(f"{fn[:88]:88s} FADD2={c.get('FADD2',0):4d} FMUL={c.get('FMUL',0):4d} FADD={c.get('FADD',0):4d} "
          f"PRMT={c.get('PRMT',0):4d} FFMA(div/sqrt)={c.get('FFMA',0)-stray} FFMA2={c.get('FFMA2',0)} stray={stray} total={sum(c.values())}")
    if stray or c.get("FMUL2", 0):
        bad += 1
if bad:
    print(f"FAIL: {bad} parity-critical kernel(s) contain contracted multiply-adds")
    sys.exit(1)
print("OK: no contracted scalar FFMA and no packed multiplies in parity-critical kernels")
