"""Turn an ``ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`` log of ONE training step
(tools/gpu_traffic.sh) into profiles/r2_ncu_traffic.json: DRAM bytes of all tensor-core convolution launches of the step and the
per-launch DRAM bytes of the fused photometric forward / backward kernels.  bench.py reports these as ``roofline.traffic``."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "traffic.csv")
lines = [l for l in open(src) if not l.startswith("==")]
per = collections.defaultdict(lambda: collections.defaultdict(float))   # launch id -> metric -> value
names = {}
for row in csv.DictReader(lines):
    m = row.get("Metric Name")
    if m not in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
    per[row["ID"]][m] = v * scale
    names[row["ID"]] = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "")
conv = [i for i, n in names.items() if "conv_tc" in n]
pf = [i for i, n in names.items() if "photometric_fwd" in n]
pb = [i for i, n in names.items() if "photometric_bwd" in n]
tot = lambda ids: sum(per[i]["dram__bytes_read.sum"] + per[i]["dram__bytes_write.sum"] for i in ids)
out = {
    "conv_step_dram_bytes": tot(conv), "conv_launches": len(conv), "conv_step_us": sum(per[i]["gpu__time_duration.sum"] for i in conv),
    "conv_source": "profiles/r2_ncu_traffic.json <- ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d conv_tc_* launches of one eager "
                   "single-stream step of `bench.py --steps 1` (tools/gpu_traffic.sh)" % len(conv),
    "photometric_fwd_dram_bytes": tot(pf) / max(len(pf), 1), "photometric_fwd_launches": len(pf),
    "photometric_fwd_us": sum(per[i]["gpu__time_duration.sum"] for i in pf) / max(len(pf), 1),
    "photometric_bwd_dram_bytes": tot(pb) / max(len(pb), 1), "photometric_bwd_launches": len(pb),
    "photometric_bwd_us": sum(per[i]["gpu__time_duration.sum"] for i in pb) / max(len(pb), 1),
    "photometric_source": "same capture: mean over the photometric launches of the step (4 scales)",
}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
