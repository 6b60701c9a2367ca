"""Join an ncu SASS source page (csv) with nvdisasm -g line info: stall samples per source line.
   python tools/ncu_lines.py <src_sass.csv> <nvdisasm -g output> <kernel symbol>"""
import csv, re, sys, collections
csvp, sassp, sym = sys.argv[1:4]
# offsets -> (file, line)
line_of = {}
cur = None; inside = False; nins = 0
for l in open(sassp):
    if l.startswith(".text." + sym + ":"): inside = True; continue
    if inside and l.startswith("//-----"): break
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: line_of[int(m.group(1), 16)] = cur; nins += 1
rows = list(csv.reader(open(csvp)))
blk = None
for i, r in enumerate(rows):
    if r and r[0] == "Address": hdr = r; start = i + 1; break
h = {n: i for i, n in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows[start:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    a = int(r[0], 16)
    if base is None: base = a
    key = line_of.get(a - base, ("?", 0))
    c = agg[key]
    c["n"] += int(r[h["# Samples"]]); c["inst"] += int(r[h["Instructions Executed"]]); c["sass"] += 1
    for s in ("stall_no_inst", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_branch_resolving", "stall_membar"):
        c[s] += int(r[h[s]])
tot = sum(c["n"] for c in agg.values())
print("instructions in kernel:", nins, "samples:", tot)
src = {}
for (f, ln), c in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    if f not in src:
        try: src[f] = open("/root/repo/pysages_b200/csrc/" + f).read().split("\n")
        except OSError: src[f] = []
    text = src[f][ln - 1].strip()[:70] if 0 < ln <= len(src[f]) else ""
    print(f"{c['n']:5d} noinst={c['stall_no_inst']:4d} long={c['stall_long_sb']:4d} short={c['stall_short_sb']:4d} bar={c['stall_barrier']:4d} wait={c['stall_wait']:4d} sass={c['sass']:5d} exec={c['inst']:7d} {f}:{ln}: {text}")
