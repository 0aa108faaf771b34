"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS instructions."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
def f(r, k):
    try: return float(r[idx[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data); inst = sum(f(r, "Instructions Executed") for r in data)
print(f"samples {tot:.0f}  warp-instructions {inst:.3e}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall reasons (all samples):", ", ".join(f"{k[6:]} {100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
exc = sum(f(r, "L1 Wavefronts Shared Excessive") for r in data); wf = sum(f(r, "L1 Wavefronts Shared") for r in data)
print(f"shared wavefronts {wf:.3e} of which excessive {exc:.3e}")
print("top instructions by samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
    top = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{r[idx['Address']][-5:]} {100*f(r,'# Samples')/max(tot,1):5.2f}%  inst {f(r,'Instructions Executed'):.2e}  wf {f(r,'L1 Wavefronts Shared'):.2e}/{f(r,'L1 Wavefronts Shared Ideal'):.2e}  {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}  {r[idx['Source']][:70]}")
