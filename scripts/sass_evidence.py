#!/usr/bin/env python3
"""SASS / resource evidence of the hot kernels, read from the built library with cuobjdump (no GPU needed).

    python scripts/sass_evidence.py > profiles/r02_sass.md

For every kernel family: registers / stack / static shared memory of each instantiation (cuobjdump -res-usage) and, for
the instantiations the BASELINE configs run, a histogram of the SASS mnemonics that matter for an HBM-bound fp64 path
(256-bit global loads / stores, cp.async = LDGSTS, TMA bulk copies = UBLKCP, mbarrier = SYNCS, L2 prefetch = CCTL /
PREFETCH, DFMA / DADD / DMUL, shuffles, barriers) plus the first lines of the longest straight-line DFMA run -- the
in-register K-step loop of the fused passes.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tetrakis-sim_b200", "lib", "libtetrakis_b200.so")

# demangled-name fragments of the instantiations the BASELINE configurations run
FOCUS = [
    ("lat_fused2d_kernel<4, true, double, false>", "cfg2 pass (no damping, no cross)"),
    ("lat_fused2d_kernel<4, true, double, true>", "2-D pass of the cross scheme (lattices with defects)"),
    ("fx_bulk3d_kernel<16, true, 2>", "3-D bulk pass, K=16 ticks unrolled, no damping (cfg3 / cfg4)"),
    ("fx_cross_step_kernel<1>", "cross cells + line advance, plain lattice rows (cfg3 / cfg4)"),
    ("fx_cross_step_kernel<0>", "cross cells, rows with detail records"),
    ("lat_step_kernel<1, 1, true, 0, 2>", "3-D step kernel (slabs; one step per launch)"),
    ("lat_step_kernel<1, 8, false, 0, 3>", "2-D step kernel (one step per launch)"),
    ("ensemble_kernel(", "cfg5 ensemble"),
    ("graph_step_kernel", "generic sliced-ELL graph kernel"),
    ("fft_stockham", "Bluestein / Stockham FFT stage"),
]
PATTERNS = collections.OrderedDict([
    ("LDG.E*.256 (256-bit global load)", r"\bLDG\.E[.\w]*\.256"),
    ("LDG.E*.128", r"\bLDG\.E[.\w]*\.128"),
    ("LDG.E*.64", r"\bLDG\.E[.\w]*\.64"),
    ("STG.E*.256 (256-bit global store)", r"\bSTG\.E[.\w]*\.256"),
    ("STG.E*.128", r"\bSTG\.E[.\w]*\.128"),
    ("STG.E*.64", r"\bSTG\.E[.\w]*\.64"),
    ("LDGSTS (cp.async)", r"\bLDGSTS"),
    ("UBLKCP (cp.async.bulk, TMA)", r"\bUBLKCP"),
    ("UTMALDG (tensor-map TMA)", r"\bUTMALDG"),
    ("SYNCS (mbarrier)", r"\bSYNCS"),
    ("CCTL / PREFETCH (L2 prefetch)", r"\b(CCTL|PREFETCH)"),
    ("LDS", r"\bLDS"),
    ("STS", r"\bSTS"),
    ("DFMA", r"\bDFMA"),
    ("DADD", r"\bDADD"),
    ("DMUL", r"\bDMUL"),
    ("SHFL", r"\bSHFL"),
    ("BAR.SYNC", r"\bBAR\.SYNC"),
    ("LDL / STL (local memory)", r"\b(LDL|STL)"),
    ("UTC*MMA / HMMA (tensor cores)", r"\b(UTC\w*MMA|HMMA|DMMA)"),
])


def run(*a):
    return subprocess.run(a, capture_output=True, text=True, check=True).stdout


def demangle(names):
    out = run("c++filt", *names).splitlines()
    return dict(zip(names, out))


def main():
    res = run("cuobjdump", "-res-usage", LIB)
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = dict(kv.split(":") for kv in line.split() if ":" in kv and not kv.startswith("CONSTANT"))
            cur = None
    dm = demangle(list(usage))
    sass = run("cuobjdump", "-sass", LIB)
    bodies, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            bodies[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            bodies[name].append(re.sub(r"/\*[0-9a-f]{4,}\*/", "", line).split("/*")[0].strip())

    print("# SASS and resource evidence (round 2)\n")
    print(f"Read from `{os.path.relpath(LIB, ROOT)}` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a -lineinfo`) with")
    print("`cuobjdump -res-usage` / `cuobjdump -sass` by `scripts/sass_evidence.py`; no GPU involved.\n")
    fam = collections.OrderedDict()
    for mangled, u in usage.items():
        d = dm[mangled]
        short = re.sub(r"^void ", "", d).split("(")[0]
        f = re.sub(r"<.*", "", short).replace("tk::", "")
        fam.setdefault(f, []).append((short.replace("tk::", ""), u))
    print("## Registers / stack / static shared memory per kernel family\n")
    print("| family | instantiations | registers (min-max) | stack bytes (max) | static smem (max) |\n|---|---|---|---|---|")
    for f, lst in fam.items():
        regs = [int(u["REG"]) for _, u in lst]
        print(f"| `{f}` | {len(lst)} | {min(regs)}-{max(regs)} | {max(int(u['STACK']) for _, u in lst)} | "
              f"{max(int(u['SHARED']) for _, u in lst)} |")
    print("\nStack bytes on the step / cross kernels belong to the out-of-line detail path (rows that hold defect records);")
    print("the streaming path of the same kernels uses none (no LDL/STL inside the row loop, see the histograms).\n")

    print("## Mnemonic histograms of the instantiations the BASELINE configs run\n")
    totals = collections.Counter()
    for mangled, body in bodies.items():
        for label, pat in PATTERNS.items():
            totals[label] += sum(1 for l in body if re.search(pat, l))
    for frag, what in FOCUS:
        hits = [m for m in bodies if frag in dm.get(m, "")]
        if not hits:
            continue
        m = hits[0]
        body = bodies[m]
        u = usage.get(m, {})
        print(f"### `{dm[m].split('(')[0].replace('void ', '').replace('tk::', '')}` -- {what}\n")
        print(f"{len(body)} SASS instructions, {u.get('REG', '?')} registers, stack {u.get('STACK', '?')} B, "
              f"static smem {u.get('SHARED', '?')} B\n")
        print("| mnemonic | count |\n|---|---|")
        for label, pat in PATTERNS.items():
            n = sum(1 for l in body if re.search(pat, l))
            if n:
                print(f"| {label} | {n} |")
        # longest run of consecutive DFMA instructions
        best, cur, start = (0, 0), 0, 0
        for i, l in enumerate(body):
            if re.search(r"\bDFMA", l):
                if cur == 0:
                    start = i
                cur += 1
                if cur > best[0]:
                    best = (cur, start)
            else:
                cur = 0
        if best[0] >= 8:
            print(f"\nLongest straight run of DFMA: {best[0]} instructions; its first 8:\n\n```")
            for l in body[best[1]:best[1] + 8]:
                print("    " + l)
            print("```")
        wide = [l for l in body if re.search(r"\b(LDG|STG)\.E[.\w]*\.256|LDGSTS|UBLKCP|CCTL|PREFETCH", l)][:6]
        if wide:
            print("\nMemory instructions (first 6 of the wide / asynchronous ones):\n\n```")
            for l in wide:
                print("    " + l)
            print("```")
        print()
    print("## Whole library\n\n| mnemonic | count |\n|---|---|")
    for label in PATTERNS:
        print(f"| {label} | {totals[label]} |")
    print("\nNo tensor-core instruction anywhere: the path has no contraction (DESIGN.md section 2); the work is fp64 FMA")
    print("on streamed state, bounded by HBM.")


if __name__ == "__main__":
    sys.exit(main())
