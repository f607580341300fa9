#!/bin/bash
# One GPU-box pass that produces every number profiles/README.md quotes for a round (run from the repo root under
# gpurun; outputs under gpurun_out/<tag>_*):  tools/measure_all.sh r2k
tag=${1:-rX}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests -q -m gpu > $o/${tag}_tests.log 2>&1; tail -3 $o/${tag}_tests.log
# DRAM traffic of the tensor-core layer launches of one step, stamped with the kernel-source sha (-> roofline.traffic)
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k 'regex:linear_tc|sa_stack' --csv --log-file $o/${tag}_linear_dram.csv python tools/one_step.py 1 > /dev/null 2>&1
python tools/linear_traffic.py $o/${tag}_linear_dram.csv 17
# launch list of two steps (per-launch durations, cold cache, serialised)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python tools/one_step.py 2 > /dev/null 2>&1
timeout 400 python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -2 $o/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/${tag}_reference.json 2> $o/${tag}_reference.err
timeout 300 python bench.py --workload train --steps 20 --warmup 5 > $o/${tag}_train.json 2>> $o/${tag}_bench.err
timeout 300 python bench.py --workload igr --steps 5 --warmup 3 > $o/${tag}_igr.json 2>> $o/${tag}_bench.err
timeout 300 python bench.py --workload stress --steps 10 --warmup 3 > $o/${tag}_stress.json 2>> $o/${tag}_bench.err
# eval-mode sa1: the one-kernel level against the per-layer kernels
timeout 300 python tools/bench_sa_stack.py 2>> $o/${tag}_bench.err | tail -1 > $o/${tag}_sa_stack.json
python - <<P
import json
for n in ("bench", "reference", "train", "igr", "stress"):
    try:
        d = json.load(open("$o/${tag}_%s.json" % n))
        print(n, round(d["value"], 1), d["unit"], round(d["ms_per_step"], 3), "ms", "e2e", d.get("e2e", {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(n, "FAILED", e)
P
