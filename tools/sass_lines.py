"""Short SASS excerpts that show the Blackwell-specific instructions of this library's kernels (run here, no GPU):
   python tools/sass_lines.py > profiles/r02_sass_excerpt.txt
UBLKCP = cp.async.bulk (1-D TMA), SYNCS.* = mbarrier operations, per B200_PROFILING.md."""
import re, subprocess, sys, os
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pysages_b200", "lib", "libpysages_b200.so")
KERNELS = [
    ("_ZN3psb11step_kernelILi1EfLi2ELi1EEEvPKNS_5ModelENS_4SnapE", "step_kernel<HOOMD, float, METAD_DIRECT, general>: pass H, hills staged by 1-D TMA bulk copies (cfg4)"),
    ("_ZN3psb11step_kernelILi1EfLi1ELi3EEEvPKNS_5ModelENS_4SnapE", "step_kernel<HOOMD, float, HARMONIC, stream>: opt-in TMA-staged slot-ordered class (per-warp rings, bulk stores)"),
]
pat = re.compile(r"UBLKCP|SYNCS|UTMA|FENCE\.VIEW\.ASYNC|CCTL")
for sym, title in KERNELS:
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if re.search(r"/\*[0-9a-f]{4,}\*/", l)]
    hits = [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines if pat.search(l)]
    counts = {}
    for l in hits:
        m = re.search(r"(UBLKCP[.\w]*|SYNCS[.\w]*|UTMA\w*)", l)
        if m: counts[m.group(1)] = counts.get(m.group(1), 0) + 1
    print(f"== {title}\n   symbol {sym}: {len(lines)} SASS instructions; " + ", ".join(f"{k} x{v}" for k, v in sorted(counts.items())))
    for l in hits[:14]:
        print("   " + l.strip())
    print()
