#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list of
`bench.py --no-cuda-graph` into per-kernel time / launches per training step.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rN_step_launches_summary.txt

Steps are delimited by the fused AdamW launches (one per step): only the kernels between the first and the last
`adamw_kernel` are counted.  ncu serialises launches and runs them cold-cache, so only SHARES are meaningful."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.reader(lines):
        if r[0] == "ID" or len(r) < 15 or r[12] != "gpu__time_duration.sum":
            continue
        ns = float(r[14].replace(",", "")) * ({"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[13], 1.0))
        rows.append((r[4], ns))
    marks = [i for i, (k, _) in enumerate(rows) if "adamw_kernel" in k]
    if len(marks) < 2:
        raise SystemExit(f"need >= 2 adamw_kernel launches to delimit steps, found {len(marks)} in {len(rows)} launches")
    steps = len(marks) - 1
    seg = rows[marks[0] + 1: marks[-1] + 1]
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in seg:
        k = re.sub(r"^void ", "", k).replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        agg[k][0] += 1
        agg[k][1] += ns
    total = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --no-cuda-graph (eager), cfg-2,",
          f"{steps} steps between AdamW launches")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"# total {total / steps / 1e3:.0f} us/step (serialised), {len(seg) / steps:.0f} launches/step")
    fam = defaultdict(float)
    for k, (n, ns) in agg.items():
        name = k.split("(")[0].split("<")[0]
        fam["gemm_tcgen05" if ("gemm_tcgen05" in name or "gemm_dw_grouped" in name) else "flash_attn" if name.startswith("flash_") else
            "layernorm/colsum/cast" if any(s in name for s in ("add_dropout_ln", "colsum", "add_cast")) else
            "set abstraction + pointops" if any(s in name for s in ("sa_", "knn_", "fps_")) else
            "optimizer" if any(s in name for s in ("adamw", "sumsq", "_slices_kernel")) else
            "batchnorm / ffn gate / heads / tokens (own)" if any(s in name for s in ("bn_", "ffn_", "act_heads", "coord_embed", "fill_head",
                                                                                   "unet", "groupnorm", "mish", "grid_", "spconv"))
            else "ATen / library glue"] += ns
    for f_, ns in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"# family {f_:28s} {ns / steps / 1e3:9.1f} us/step {100 * ns / total:5.1f}%")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{ns / steps / 1e3:9.1f} us/step {n / steps:7.1f} launches/step {100 * ns / total:5.1f}%  {k[:140]}")


if __name__ == "__main__":
    main(sys.argv[1])
