#!/usr/bin/env python
"""profiles/<out>_k2_sass.txt: what the in-tree libbandup_gpu.so contains for K2 -- target
architecture, the mnemonics that prove the bulk-copy (TMA engine) ring and its mbarriers, and the
hot loop of one instantiation.  No GPU needed (cuobjdump reads the fat binary).

    python scripts/sass_evidence.py r02
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "bandup_b200" / "lib" / "libbandup_gpu.so"
out = sys.argv[1] if len(sys.argv) > 1 else "r02"

elf = subprocess.run(["cuobjdump", "-lelf", str(LIB)], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
digest = (ROOT / "bandup_b200" / "lib" / "libbandup_gpu.digest").read_text().strip()

funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(line)
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))

lines = [f"# SASS evidence for libbandup_gpu.so (source digest {digest[:16]})", "",
         "cuobjdump -lelf:", *["  " + ln for ln in elf.strip().splitlines()], "",
         f"architectures in the fat binary: {', '.join(arch)}", "",
         "## mnemonic counts per kernel (cuobjdump -sass)", "",
         "| kernel | instructions | UBLKCP (bulk async copy, TMA engine) | SYNCS (mbarrier) | LDS.128 | LDG | DADD | F2F.F64.F32 | BAR | registers see build.log |",
         "|---|---|---|---|---|---|---|---|---|---|"]


def demangle(n):
    r = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
    i = r.rfind(">(")
    r = r[:i + 1] if i >= 0 else r[:r.find("(")] if "(" in r else r
    return r.replace("void ", "").replace("bandup::", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")


hot = None
for name, body in funcs.items():
    if not body:
        continue
    txt = "\n".join(body)
    cnt = lambda pat: len(re.findall(pat, txt))            # noqa: E731
    d = demangle(name)
    lines.append(f"| `{d}` | {len(body)} | {cnt(r'UBLKCP')} | {cnt(r'SYNCS')} | {cnt(r'LDS\.128')} | {cnt(r'LDG')} | "
                 f"{cnt(r'DADD')} | {cnt(r'F2F\.F64\.F32')} | {cnt(r'BAR\.')} | |")
    if d == "k2_weights_reduce<0, 0>":
        hot = (d, body)

if hot:
    d, body = hot
    # the producer's copy loop and one full consumer stage: from the first UBLKCP to the DADD pair after the next LDS.128 group
    idx = [i for i, ln in enumerate(body) if "UBLKCP" in ln]
    lds = [i for i, ln in enumerate(body) if "LDS.128" in ln]
    lines += ["", f"## `{d}`: the producer's bulk copy (expect_tx arrive, UBLKCP, phase flip)", "", "```"]
    a = max(0, idx[0] - 14)
    lines += [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln) for ln in body[a:idx[0] + 6] if re.search(r"/\*[0-9a-f]{4}\*/\s+\S", ln)]
    lines += ["```", "", "## the same kernel: one full 16 KB stage on the consumer side (slot ids, try_wait, 4 x LDS.128, arrive, |C|^2, selection)", "", "```"]
    a = max(0, lds[0] - 18)
    lines += [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln) for ln in body[a:lds[0] + 70] if re.search(r"/\*[0-9a-f]{4}\*/\s+\S", ln)]
    lines += ["```"]
(ROOT / "profiles" / f"{out}_k2_sass.txt").write_text("\n".join(lines) + "\n")
print("wrote", ROOT / "profiles" / f"{out}_k2_sass.txt", len(lines), "lines")
