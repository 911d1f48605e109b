#!/usr/bin/env python
"""Merge the three per-kernel measurements of one denoise step (768x512) into profiles/per_kernel_r2.csv:
  * in-graph duration and start-to-next-start slot (globaltimer stamps written by the kernels inside the captured 17-step
    graph: bench.py --ops-out),
  * the ncu launch list of one eager step (cold caches, serialised: compare shares, not absolutes),
  * tensor-pipe utilisation and DRAM bytes from the ncu --set full captures (where that launch was captured).
Usage: per_kernel_table.py per_kernel_ingraph.csv launches.csv out.csv ncu_summary.csv[:class[:first[+k]|last]] ..."""
import csv
import sys


def launch_list(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
    return [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows[h + 2:] if len(r) > vi]


ingraph = list(csv.DictReader(open(sys.argv[1])))
step = launch_list(sys.argv[2])[-len(ingraph):]
assert len(step) == len(ingraph)
full = [None] * len(step)
for spec in sys.argv[4:]:
    path, cls, where = (spec.split(":") + ["", "first"])[:3]
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], rows[1:]
    col = {name.split(" [")[0]: i for i, name in enumerate(hdr)}
    idx = [i for i, (n, _) in enumerate(step) if any(c in n for c in cls.split("|"))]
    off = int(where.split("+")[1]) if "+" in where else 0  # "first+1": the capture started one launch into the step
    idx = idx[off:off + len(data)] if where.startswith("first") else idx[-len(data):]
    for i, r in zip(idx, data):
        assert r[0].split("(")[0] == step[i][0].split("(")[0], (r[0], step[i][0])
        full[i] = (float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
                   float(r[col["dram__bytes_read.sum"]]) * (1e-6 if "[byte]" in hdr[col["dram__bytes_read.sum"]] else 1.0),
                   float(r[col["dram__bytes_write.sum"]]) * (1e-6 if "[byte]" in hdr[col["dram__bytes_write.sum"]] else 1.0))
with open(sys.argv[3], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["op", "kernel", "gflop", "algorithmic_mbytes", "us_in_graph", "slot_us_in_graph", "tflops_in_graph", "gbs_in_graph",
                "us_ncu_cold", "tensor_pipe_pct_ncu", "dram_read_mb_ncu", "dram_write_mb_ncu"])
    for r, (name, us), fl in zip(ingraph, step, full):
        k = name.replace("void ", "").split("(")[0]
        w.writerow([r["op"], k, r["gflop"], r["mbytes"], r["us_in_graph"], r["slot_us"], r["tflops"], r["gbs"], f"{us:.2f}"]
                   + ([f"{fl[0]:.1f}", f"{fl[1]:.2f}", f"{fl[2]:.2f}"] if fl else ["", "", ""]))
tot = sum(float(r["us_in_graph"]) for r in ingraph)
print(f"{len(ingraph)} kernels: in-graph {tot:.1f} us (slots {sum(float(r['slot_us']) for r in ingraph):.1f}), ncu cold {sum(u for _, u in step):.1f} us")
