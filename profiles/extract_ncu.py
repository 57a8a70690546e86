"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) into the few metrics
DESIGN.md / bench.py quote: duration, DRAM bytes, tensor-pipe activity, registers.  Writes a markdown table and the
dominant kernel's DRAM traffic as JSON (bench.py reads profiles/r1_ncu_dominant.json for roofline.traffic)."""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active": "hmma_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}


def main(path, md_out, json_out):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[2:]:
        d = {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", ""), "grid": r[col["Grid Size"]]}
        for m, k in WANT.items():
            if m in col:
                v = float(r[col[m]].replace(",", "") or 0)
                d[k] = v * UNIT_SCALE.get(units[col[m]], 1)
        recs.append(d)
    with open(md_out, "w") as f:
        f.write("| kernel | grid | us | DRAM read MB | DRAM write MB | tensor pipe % | SM % | DRAM % | regs | dyn smem KB |\n")
        f.write("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for d in recs:
            f.write(f"| `{d['kernel']}` | {d['grid']} | {d.get('us', 0):.1f} | {d.get('dram_read', 0) / 1e6:.2f} | "
                    f"{d.get('dram_write', 0) / 1e6:.2f} | {d.get('tensor_pipe_pct', 0):.1f} | {d.get('sm_pct', 0):.1f} | "
                    f"{d.get('dram_pct', 0):.1f} | {int(d.get('regs', 0))} | {d.get('dyn_smem', 0) / 1e3:.0f} |\n")
    def ctas(d):
        g = [int(x) for x in d["grid"].strip("() ").split(",")]
        return g[0] * g[1] * g[2]

    # dominant launch = the forward launch with the most CTAs (most algorithmic FLOPs), longest among equals
    dom = max(recs, key=lambda d: (ctas(d), d.get("us", 0)) if "forward" in d["kernel"] else (-1, 0))
    json.dump({"kernel": dom["kernel"], "grid": dom["grid"], "us_under_ncu": dom.get("us"),
               "dram_bytes_per_launch": dom.get("dram_read", 0) + dom.get("dram_write", 0),
               "tensor_pipe_pct": dom.get("tensor_pipe_pct"),
               "source": "ncu --set full --clock-control none, profiles/run_c2_update.py (eager c2 update, bf16 mode)"},
              open(json_out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
