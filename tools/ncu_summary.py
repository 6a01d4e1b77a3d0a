"""Condense an `ncu --set full` report into the numbers DESIGN.md / bench.py quote: per kernel duration, DRAM bytes,
pipe utilisation, issue slots, registers, plus the top stall reasons of the source view.

    python tools/ncu_summary.py report.ncu-rep out.json
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return x


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
    # the source page holds one section per profiled launch: a "Kernel Name" line, the column header, the instructions
    names = [i for i, l in enumerate(src) if l.startswith('"Kernel Name"')]
    starts = [i for i, l in enumerate(src) if l.startswith('"Address"')] + [len(src) + 1]
    n_k = len(rows) - 2
    per = len(names) // n_k if n_k and len(names) % n_k == 0 else 0      # 1, or 2 (SASS view + source-line view) sections per launch
    aligned = per > 0 and len(starts) - 1 == len(names)
    kernels = []
    for n, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        k = {"kernel": d["Kernel Name"].split("(")[0], "metrics": {}}
        for key in KEYS:
            if key in d:
                k["metrics"][key] = {"value": num(d[key]), "unit": unit.get(key, "")}
        rd, wr = k["metrics"].get("dram__bytes_read.sum"), k["metrics"].get("dram__bytes_write.sum")
        if rd and wr:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            k["dram_bytes"] = rd["value"] * scale.get(rd["unit"], 1.0) + wr["value"] * scale.get(wr["unit"], 1.0)
        if aligned:
            sec = per * n
            end = names[sec + 1] if sec + 1 < len(names) else len(src)
            srows = list(csv.reader(src[starts[sec]:end]))
            sh = srows[0]
            ix = {name: i for i, name in enumerate(sh)}
            stalls = [c for c in sh if c.startswith("stall_") and "Not Issued" not in c]
            agg = collections.Counter()
            total = 0
            for sr in srows[1:]:
                if len(sr) != len(sh):
                    continue
                total += int(sr[ix["# Samples"]])
                for c in stalls:
                    agg[c] += int(sr[ix[c]])
            k["warp_samples"] = total
            k["top_stalls"] = dict(agg.most_common(6))
        kernels.append(k)
    json.dump({"report": rep, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        m = k["metrics"]
        print(k["kernel"], m.get("gpu__time_duration.sum"), "dram", k.get("dram_bytes"), k.get("top_stalls"))


if __name__ == "__main__":
    main()
