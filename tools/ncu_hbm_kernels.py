"""Turn the ncu CSV of tools/bench_geometry.py (metrics gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum) into the per-kernel HBM table the north_star asks for: kernel time, algorithmic bytes (from the
tool's own JSON), DRAM bytes, achieved GB/s = algorithmic bytes / kernel time, fraction of the measured copy peak.
usage: python tools/ncu_hbm_kernels.py <ncu.csv> <bench_geometry.json> [peak_gbs]"""
import csv, json, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
alg = json.load(open(sys.argv[2]))
peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6451.2
k = defaultdict(lambda: defaultdict(list))
for r in rows:
    name, metric, value = r[4], r[-3], float(r[-1].replace(",", ""))
    unit = r[-2]
    if metric.startswith("gpu__time_duration"):
        value *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)      # -> ms
    if metric.startswith("dram__bytes"):
        value *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    k[name.split("(")[0].split("<")[0]][metric].append(value)
want = {"generate_rays4_kernel": "generate_rays", "sample_points_rays_kernel": ("sample_points", "sample_points_jitter"),
        "encode_rows_kernel": ("positional_encoding_L10", "positional_encoding_L4"), "composite4_kernel": ("composite", "composite_with_weights"),
        "importance_warp_kernel": "importance_sample", "merge_warp_kernel": "merge_samples",
        "hierarchical_samples_warp_kernel": ("hierarchical_samples_u_given", "hierarchical_samples_philox"),
        "composite_white4_kernel": "composite_white", "ray_batch_kernel": "ray_batch_4096"}
out = {}
for kern, m in k.items():
    base = kern.split("::")[-1].replace("void ", "").strip()
    if base not in want:
        continue
    t = sorted(m["gpu__time_duration.sum"])
    rd, wr = sorted(m.get("dram__bytes_read.sum", [0])), sorted(m.get("dram__bytes_write.sum", [0]))
    out[base] = {"launches": len(t), "ms_min": t[0], "ms_median": t[len(t) // 2], "dram_read_MB_median": rd[len(rd) // 2] / 1e6,
                 "dram_write_MB_median": wr[len(wr) // 2] / 1e6, "variants": want[base]}
print(json.dumps({"peak_gbs": peak, "kernels": out, "algorithmic": alg}, indent=1))
