"""Executed warp-instructions per SOURCE LINE for one kernel of an ncu report (needs --import-source on and a
-lineinfo build): aligns the report's SASS page with `nvdisasm -g` of the same cubin by instruction order.
usage: ncu_lines.py report.ncu-rep lib.so kernel_substr mangled_prefix"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, lib, kname, mangled = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur); continue
    if r and r[0] == "Address":
        cur["hdr"] = r; continue
    if cur is not None and r: cur["rows"].append(r)
sec = [s for s in secs if kname in s["name"] and "hdr" in s][0]
h = sec["hdr"]; ie = h.index("Instructions Executed")
execd = [(r[1].strip(), int(r[ie])) for r in sec["rows"]]
d = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, "kernels.sm_100a.cubin")], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if re.match(r"\s*\.section\s+\.text\." + re.escape(mangled), l))
end = next(i for i in range(start + 1, len(dis)) if re.match(r"\s*\.section\s+\.text\.", dis[i]))
seq, curl = [], None
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: curl = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+", l): seq.append(curl)
assert len(seq) == len(execd), (len(seq), len(execd))
cnt = collections.Counter(); tot = 0
for ln, (ins, n) in zip(seq, execd):
    cnt[ln] += n; tot += n
print("total", tot)
src = {}
for (f, ln), n in cnt.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 40):
    path = next((p for p in (f"dacapo_b200/csrc/{f}",) if os.path.exists(p)), None)
    text = ""
    if path:
        src.setdefault(path, open(path).read().splitlines())
        text = src[path][ln - 1].strip()[:90]
    print(f"{n:10d} {100*n/tot:5.1f}%  {f}:{ln}  {text}")
