# offline (no GPU): turn the reports of call w into the committed evidence files
set -e
cd "$(dirname "$0")/../.."
for f in pytest.txt e2e_in_place.txt bench.json bench_reference.json launches.csv; do [ -f gpurun_out/r02_w_$f ] && cp gpurun_out/r02_w_$f profiles/r02_w_$f; done
ncu -i gpurun_out/r02_w_fused.ncu-rep --page details --csv > profiles/r02_w_fused_narrow_kernel_ncu_details.csv
ncu -i gpurun_out/r02_w_comp.ncu-rep --page details --csv > profiles/r02_w_compress_kernels_ncu_details.csv
python profiles/tools/ncu_sol_table.py profiles/r02_w_sol_summary.md gpurun_out/r02_w_fused.ncu-rep gpurun_out/r02_w_comp.ncu-rep
ncu -i gpurun_out/r02_w_fused.ncu-rep --page source --csv > gpurun_out/r02_w_fused_source.csv 2>/dev/null || true
python profiles/tools/ncu_source_summary.py gpurun_out/r02_w_fused_source.csv > profiles/r02_w_fused_hot_instructions.txt 2>/dev/null || true
python - <<'PY'
import csv, io, json, subprocess
raw = subprocess.run(["ncu", "-i", "gpurun_out/r02_w_fused.ncu-rep", "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, row = rr[0], rr[1], rr[2]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
total = int(float(row[ir]) * scale[units[ir]] + float(row[iw]) * scale[units[iw]])
json.dump({"dram_bytes_per_launch": {"fused_narrow_kernel": total},
           "source": "profiles/r02_w_fused_narrow_kernel_ncu_details.csv: ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of the one launch, "
                     "1024 chunks x 2^18 u64 (C2 data), profiles/tools/r02_w_call.sh"}, open("profiles/roofline_traffic.json", "w"), indent=1)
print("fused_narrow_kernel DRAM bytes per launch", total)
PY
