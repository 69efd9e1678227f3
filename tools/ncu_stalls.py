"""Stall samples per SOURCE LINE (with the top reasons) for one kernel of an ncu report.
usage: ncu_stalls.py report.ncu-rep lib.so kernel_substr mangled_prefix [top]"""
import collections, csv, io, re, subprocess, os, tempfile, sys
rep, lib, kname, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur); continue
    if r and r[0] == "Address":
        cur["hdr"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
d = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, "kernels.sm_100a.cubin")], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if re.match(r"\s*\.section\s+\.text\." + re.escape(mangled), l))
end = next((i for i in range(start + 1, len(dis)) if re.match(r"\s*\.section\s+\.text\.", dis[i])), len(dis))
seq, curl = [], None
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: curl = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+", l): seq.append(curl)
sec = [s for s in secs if kname in s["name"] and "hdr" in s][0]
h = sec["hdr"]; isamp = h.index("# Samples"); ie = h.index("Instructions Executed")
reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
assert len(seq) == len(sec["rows"]), (len(seq), len(sec["rows"]))
byline = collections.Counter(); byreason = collections.Counter(); lr = collections.defaultdict(collections.Counter); ex = collections.Counter()
tot = 0
for ln, r in zip(seq, sec["rows"]):
    n = int(r[isamp]); byline[ln] += n; tot += n; ex[ln] += int(r[ie])
    for c in reasons:
        v = int(r[h.index(c)] or 0); byreason[c] += v; lr[ln][c] += v
print("==", kname, "samples", tot, "instructions", sum(ex.values()))
print("  reasons:", [(k.replace("stall_", ""), v) for k, v in byreason.most_common(9)])
for ln, n in byline.most_common(top):
    print("  %6d %5.1f%%  exec %9d  %s:%s  %s" % (n, 100 * n / tot, ex[ln], ln[0], ln[1], [(k.replace("stall_", ""), v) for k, v in lr[ln].most_common(3)]))
