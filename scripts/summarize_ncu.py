#!/usr/bin/env python
"""Turns gpurun_out/<tag>_launches.csv (+ <tag>_k2.ncu-rep) into the tracked summaries under
profiles/: a per-kernel launch table, the K2 ncu metrics, and profiles/k2_traffic.json (the
dram bytes per launch that bench.py reports as roofline.traffic).

    python scripts/summarize_ncu.py r3 r01 [workload]
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, out = sys.argv[1], sys.argv[2]
workload = sys.argv[3] if len(sys.argv) > 3 else "sige555"
G = ROOT / "gpurun_out"
P = ROOT / "profiles"
P.mkdir(exist_ok=True)

rows = [r for r in csv.reader(open(G / f"{tag}_launches.csv")) if r and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("bandup::", "").replace("<unnamed>::", "")
    per.setdefault(name, []).append((float(r[-1]), r[7], r[8]))
total = sum(t for v in per.values() for t, _, _ in v)
lines = [f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none`",
         "", "Cold-cache, serialised launches: compare SHARES, not absolutes.", "",
         "| kernel | launches | mean us | share of listed time | block | grid |", "|---|---|---|---|---|---|"]
for k, v in per.items():
    ts = [t for t, _, _ in v]
    lines.append(f"| `{k}` | {len(v)} | {sum(ts) / len(ts) / 1e3:.2f} | {100 * sum(ts) / total:.1f} % | {v[0][1]} | {v[0][2]} |")
own = {k: v for k, v in per.items() if k.startswith("k")}
hot = {k: v for k, v in own.items() if "synth" not in k}
tot_hot = sum(t for v in hot.values() for t, _, _ in v)
lines += ["", "Share within the hot path only (K1, K2, K3; the synthetic generator and torch fills excluded):", ""]
for k, v in hot.items():
    lines.append(f"* `{k}`: {100 * sum(t for t, _, _ in v) / tot_hot:.1f} %")
(P / f"{out}_launches.md").write_text("\n".join(lines) + "\n")

rep = G / f"{tag}_k2.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr = rr[0]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    md = [f"# ncu --set full, k2_weights_reduce ({tag}, {workload})", "",
          "`ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 2`", "",
          "| metric | unit | launch 1 | launch 2 |", "|---|---|---|---|"]
    vals = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            v = [r[i] for r in rr[2:]]
            vals[w] = (rr[1][i], v)
            md.append(f"| `{w}` | {rr[1][i]} | " + " | ".join(v) + " |")

    def to_bytes(x, unit):
        return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    rd = [to_bytes(x, vals["dram__bytes_read.sum"][0]) for x in vals["dram__bytes_read.sum"][1]]
    wr = [to_bytes(x, vals["dram__bytes_write.sum"][0]) for x in vals["dram__bytes_write.sum"][1]]
    traffic = sum(rd) / len(rd) + sum(wr) / len(wr)
    md += ["", f"DRAM traffic per launch (read + write, mean of the captured launches): **{traffic / 1e9:.4f} GB**."]
    (P / f"{out}_k2_ncu.md").write_text("\n".join(md) + "\n")
    # the capture belongs to ONE build of the library: bench.py reports it as roofline.traffic only
    # while bandup_b200/lib/libbandup_gpu.digest (a hash of every source and flag) still says the same
    digest_file = G / f"{tag}_lib.digest"
    digest = (digest_file if digest_file.exists() else ROOT / "bandup_b200" / "lib" / "libbandup_gpu.digest").read_text().strip()
    (P / "k2_traffic.json").write_text(json.dumps(
        {"workload": workload, "dram_bytes_per_launch": traffic, "lib_digest": digest,
         "source": f"profiles/{out}_k2_ncu.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}, indent=1) + "\n")
    det = subprocess.run(["ncu", "-i", str(rep), "--page", "details"], capture_output=True, text=True).stdout
    (P / f"{out}_k2_ncu_details.txt").write_text(det)
print("wrote profiles/%s_*" % out)

# ---- the other captures of gpu_final.sh: spinor K2, complex128 K2, the K1/K2/K3 chain of the small cells
WANT2 = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
         "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
         "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]
extra = [("k2_spinor", "K2 on the MoS2 8x8 spinor workload (C3, 26 SC-kpts per launch with 1-7 pc-kpts each: register accumulators and shared-memory bins mixed)"),
         ("k2_c128", "K2 on complex128 coefficients (C5 shape, `bench.py --complex128`)"),
         ("c1", "K1, K2, K3 of one step of the graphene 4x4 job (C1, 23 SC-kpts in one batch)"),
         ("c2", "K1, K2, K3 of one step of the Si 3x3x3 job (C2, 99 SC-kpts in one batch)")]
md = [f"# ncu --set full, the other captures of `scripts/gpu_final.sh {tag}`", "",
      "Per-launch times under ncu are cold-cache and serialised (no dependent-launch overlap): compare shares and",
      "bytes, not absolutes.", ""]
for suffix, title in extra:
    rep = G / f"{tag}_{suffix}.ncu-rep"
    if not rep.exists():
        continue
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        continue
    hdr = rr[0]
    names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "")
             for r in rr[2:]]
    md += [f"## {title}", "", "| metric | unit | " + " | ".join(f"`{n}`" for n in names) + " |", "|---|---|" + "---|" * len(names)]
    for w in WANT2:
        if w in hdr:
            i = hdr.index(w)
            md.append(f"| `{w}` | {rr[1][i]} | " + " | ".join(r[i] for r in rr[2:]) + " |")
    md.append("")
(P / f"{out}_other_ncu.md").write_text("\n".join(md) + "\n")
print("wrote profiles/%s_other_ncu.md" % out)
