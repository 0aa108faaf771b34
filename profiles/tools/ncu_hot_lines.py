"""Join an `ncu --page source --csv` export (SASS rows) with `nvdisasm -g -c` line info of the same kernel:
instructions executed and stall samples per source line.  usage: ncu_hot_lines.py <src.csv> <disasm.txt> [topn]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = []; seen = set()
for r in rows[2:]:
    if len(r) != len(hdr) or r[idx["Address"]] in seen or r[idx["Address"]] == "Address": continue
    seen.add(r[idx["Address"]]); data.append(r)
base = min(int(r[idx["Address"]], 16) for r in data)
line_of = {}; cur = None; infunc = True
for l in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", l)
    if m:
        off = int(m.group(1), 16)
        if off in line_of: break  # next function
        line_of[off] = cur
agg = collections.defaultdict(lambda: [0.0, 0.0])
tot_i = tot_s = 0.0
for r in data:
    off = int(r[idx["Address"]], 16) - base
    try: i = float(r[idx["Instructions Executed"]]); s = float(r[idx["# Samples"]])
    except ValueError: continue
    k = line_of.get(off); agg[k][0] += i; agg[k][1] += s; tot_i += i; tot_s += s
src = {}
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print(f"total warp-inst {tot_i:.3e}, samples {tot_s:.0f}")
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    text = ""
    if k:
        try:
            if k[0] not in src: src[k[0]] = open("/root/repo/pcodec_b200/csrc/" + k[0]).read().split("\n")
            text = src[k[0]][k[1] - 1].strip()[:100]
        except Exception: pass
    print(f"{str(k):34s} inst {100*i/tot_i:5.1f}%  samp {100*s/max(tot_s,1):5.1f}%  {text}")
