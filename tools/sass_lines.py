"""SASS instruction count per source line for one kernel: python tools/sass_lines.py <nvdisasm -g out> <symbol> [top]"""
import re, sys, collections
sassp, sym = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cnt = collections.Counter(); cur = None; inside = False; n = 0
for l in open(sassp):
    if l.startswith(".text." + sym + ":"): inside = True; continue
    if inside and l.startswith("//-----"): break
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l): cnt[cur] += 1; n += 1
print("instructions:", n, "=", n * 16, "bytes")
src = {}
for (f, ln), c in cnt.most_common(top):
    if f not in src:
        try: src[f] = open("/root/repo/pysages_b200/csrc/" + f).read().split("\n")
        except OSError: src[f] = []
    text = src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ""
    print(f"{c:6d} {f}:{ln}: {text}")
