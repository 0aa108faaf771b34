"""Compact speed-of-light table from `ncu -i <rep> --page details --csv` exports (no GPU needed to read a report): per kernel launch the
duration, DRAM / L1 / L2 / compute throughput as % of peak, achieved occupancy, registers, issue-slot use, DRAM bytes.
Usage: ncu_sol_table.py out.md rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [("Duration", "duration"), ("DRAM Throughput", "DRAM %"), ("L1/TEX Cache Throughput", "L1 %"), ("L2 Cache Throughput", "L2 %"), ("Compute (SM) Throughput", "SM %"),
        ("Achieved Occupancy", "occupancy %"), ("Registers Per Thread", "regs"), ("Issue Slots Busy", "issue slots %"), ("Executed Ipc Active", "IPC"),
        ("Theoretical Occupancy", "theoretical occ %"), ("Block Limit Registers", "CTA limit (regs)"), ("Block Limit Shared Mem", "CTA limit (smem)")]
rows = []
for rep in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    per = {}
    for r in csv.DictReader(io.StringIO(out)):
        key = (rep.split("/")[-1], r["ID"], r["Kernel Name"].split("(")[0], r["Grid Size"], r["Block Size"])
        per.setdefault(key, {})
        for metric, col in WANT:
            if r["Metric Name"] == metric and col not in per[key]:
                per[key][col] = f'{r["Metric Value"]} {r["Metric Unit"]}'.strip() if col == "duration" else r["Metric Value"]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    if len(rr) > 2:
        hdr, units = rr[0], rr[1]
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in rr[2:]:
            for key in per:
                if key[0] == rep.split("/")[-1] and key[1] == r[0]:
                    per[key]["DRAM MB"] = f"{float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]:.0f}"
    rows += sorted(per.items(), key=lambda kv: int(kv[0][1]))
cols = [c for _, c in WANT] + ["DRAM MB"]
with open(sys.argv[1], "w") as f:
    f.write("Speed-of-light summary of the committed `ncu --set full --clock-control none` captures (`tools/ncu_sol_table.py`; per-launch times under ncu are\nserialised and cold-cache - read the percentages, not the absolute durations; DRAM MB = dram__bytes_read.sum + dram__bytes_write.sum of the launch).\n\n")
    f.write("| report | kernel | grid x block | " + " | ".join(cols) + " |\n|---|---|---|" + "---|" * len(cols) + "\n")
    for (rep, _id, name, grid, block), m in rows:
        f.write(f"| {rep} | `{name}` | {grid} x {block} | " + " | ".join(m.get(c, "-") for c in cols) + " |\n")
print("wrote", sys.argv[1], len(rows), "launches")
