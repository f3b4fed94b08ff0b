#!/usr/bin/env python
"""Per-kernel SASS evidence that the scan kernels are sm_100a tcgen05 / TMA / TMEM code (B200_PROFILING.md):
counts of UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG (TMA loads), LDTM / STTM (tcgen05.ld / st), UTCBAR
(tcgen05.commit), SYNCS (mbarrier), UCGABAR (cluster barriers), ACQBULK / PREEXIT (griddepcontrol) and plain
HMMA (legacy mma.sync: must be 0) in every kernel of the built library.
usage: python scripts/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "ragroute_b200", "csrc", "libragroute_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR", "ACQBULK", "PREEXIT", "HMMA", "LDGSTS", "REDUX", "MATCH", "DFMA"]
kern, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            counts[kern]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[kern][o] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"library: {os.path.relpath(lib)}   arch: {sorted(arch)}   kernels: {len(counts)}")
print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
for (k, c), d in zip(counts.items(), demangle):
    name = re.sub(r"\(anonymous namespace\)::", "", d)
    name = re.sub(r"\(.*", "", name)[:70]
    print(f"{name:70s} {c['_total']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
