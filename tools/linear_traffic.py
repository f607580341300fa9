"""Turn an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over the MLP-layer
kernels of ONE forward+loss step) into profiles/linear_traffic.json, stamped with the sha of the kernel sources it was
measured on (bench.py reports `roofline.traffic` only when that sha matches the sources it runs).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k 'regex:linear_tc|sa_stack' --csv --log-file gpurun_out/linear_dram.csv python tools/one_step.py
    python tools/linear_traffic.py gpurun_out/linear_dram.csv [launches per step]
"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows:
    key = int(r[ix["ID"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    name = r[ix["Metric Name"]]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    elif name.startswith("gpu__time_duration"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)      # -> microseconds
    per.setdefault(key, {"kernel": r[ix["Kernel Name"]][:40]})[name] = v
launches = [per[k] for k in sorted(per)]
n_step = int(sys.argv[2]) if len(sys.argv) > 2 else len(launches)
launches = launches[-n_step:]
total = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in launches)
dur_us = sum(l.get("gpu__time_duration.sum", 0.0) for l in launches)
out = {"dram_bytes_per_step": total, "launches": len(launches), "kernel_source_sha": kernel_source_sha(),
       "ncu_duration_us_per_step": dur_us,
       "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum over the {len(launches)} tensor-core MLP-layer launches of "
                 f"one forward+loss step ({os.path.basename(path)}), same kernel sources as this build",
       "per_launch": launches}
for d in ("profiles", "gpurun_out"):
    os.makedirs(os.path.join(ROOT, d), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, d, "linear_traffic.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("dram_bytes_per_step", "ncu_duration_us_per_step", "launches", "kernel_source_sha")}))
