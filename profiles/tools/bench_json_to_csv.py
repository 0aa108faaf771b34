"""bench.py JSON lines -> rows of the reference bench tool's results CSV (pcodec_b200/benchfmt.py), so committed measurements can
be laid beside docs/benchmark_results/*.csv of the reference.  compress_dt / decompress_dt = seconds for one pass over the whole job
(uncompressed bytes / MB/s), as pco_cli reports them.   Usage: bench_json_to_csv.py out.csv bench1.json [bench2.json ...]"""
import json
import os
import sys

sys.path.insert(0, os.getcwd())
from pcodec_b200 import DeltaSpec, ModeSpec, benchfmt

rows = []
for path in sys.argv[2:]:
    d = json.loads(open(path).read().strip().split("\n")[-1])
    n_chunks = d["config"]["chunks_per_gpu"] * d["n_gpus"]
    unc = n_chunks * d["config"]["chunk_n"] * 8
    rows.append(dict(input=f"c2_u64_cumsum_geometric_{n_chunks}x2^18@{d['n_gpus']}xB200_resident", codec=benchfmt.PcoCodec(8, DeltaSpec.try_consecutive(1), ModeSpec.classic()),
                     compress_dt=unc / 1e6 / d["compress_mb_s"], decompress_dt=unc / 1e6 / d["decompress_mb_s"],
                     compressed_size=d["compressed_bytes_per_gpu"] * d["n_gpus"], uncompressed_size=unc))
print(benchfmt.merge_results_csv(sys.argv[1], rows), "rows in", sys.argv[1])
