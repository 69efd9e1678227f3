"""Static SASS summary of libB200_HEVM.so (opcode families per kernel) -> profiles/rNN_sass_summary.md.
usage: python tools/sass_summary.py [out.md]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
txt = subprocess.run(["cuobjdump", "-sass", str(REPO / "dacapo_b200" / "libB200_HEVM.so")], capture_output=True, text=True).stdout
out = ["# SASS summary of libB200_HEVM.so (sm_100a) -- `cuobjdump -sass dacapo_b200/libB200_HEVM.so`, N = 2^15 instantiations", "",
       "Instruction counts are STATIC (instructions in the kernel image), opcode families by prefix.  What to look for:",
       "* `UTMALDG` = `cp.async.bulk.tensor` (TMA) and `SYNCS` = mbarrier operations -- the cluster NTT (`k_ntt_fwd_cluster`) loads its strided",
       "  128 x 4 pass-A tiles with TMA; `UCGABAR_*` = cluster barrier; its `ST` through `MAPA`-mapped addresses are the distributed-shared-memory scatter.",
       "* `LDGSTS` = `cp.async` (twiddle / row staging in every NTT kernel).",
       "* there is no `HMMA` / `UTC*MMA`: the path is 64-bit integer arithmetic (IMAD.WIDE / IMAD / IADD3), no tensor cores (DESIGN.md section 3).",
       "* `NANOSLEEP` + `ATOMG`/`RED` = the ticket / completion counters of the single-launch key switch (`k_ks_fused`) and the epoch flags of the",
       "  peer-to-peer exchange (`k_p2p_push` / `k_p2p_wait`, whose stores go to CUDA-IPC mapped peer memory over NVLink).", "",
       "| kernel | instructions | IMAD.WIDE | IMAD (other) | IADD3 | LOP3+SHF | LDS/STS | LDG/STG/LD/ST | LDGSTS | UTMALDG | SYNCS | UCGABAR | MAPA | BAR | ATOMG/RED | NANOSLEEP |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if not any(k in name for k in ("Li7E", "k_elementwise", "k_mulp_add_n", "k_p2p", "k_ntt_fwd_cluster", "k_sample_enc")):
        continue
    ops = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter()
    for o in ops:
        b = o.split(".")[0]
        key = ("IMAD.WIDE" if o.startswith("IMAD.WIDE") else "IMAD" if b == "IMAD" else "LOP3SHF" if b in ("LOP3", "SHF") else "LDSSTS" if b in ("LDS", "STS")
               else "LDGSTG" if b in ("LDG", "STG", "LD", "ST") else "ATOM" if b in ("ATOMG", "RED", "ATOM") else "UCGABAR" if b.startswith("UCGABAR") else b)
        c[key] += 1
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    out.append(f"| `{short}` | {len(ops)} | {c['IMAD.WIDE']} | {c['IMAD']} | {c['IADD3']} | {c['LOP3SHF']} | {c['LDSSTS']} | {c['LDGSTG']} | {c['LDGSTS']} | "
               f"{c['UTMALDG']} | {c['SYNCS']} | {c['UCGABAR']} | {c['MAPA']} | {c['BAR']} | {c['ATOM']} | {c['NANOSLEEP']} |")
dst = Path(sys.argv[1]) if len(sys.argv) > 1 else REPO / "profiles" / "r02_sass_summary.md"
dst.write_text("\n".join(out) + "\n")
print(dst)
