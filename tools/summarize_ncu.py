"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/launches_r1.md
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/k1_full_r1.md
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            try:
                seq.append((r[ki], float(r[vi].replace(",", ""))))
            except ValueError:
                pass
    mine = [(n, v) for n, v in seq if any(s in n for s in ("k1_kernel", "k1_max_kernel", "k2_kernel", "search_", "argmax", "rows_kernel",
                                                           "zero_fill_kernel", "phase_each"))]
    # the last full step = the launches after the last-but-one K1 store (12 launches per step in mode=single)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: "
                "compare SHARES, not absolutes).\n\n")
        f.write(f"{len(seq)} launches captured, {len(mine)} of them ours (the rest: torch's synthetic-data generation).\n\n")
        tot = {}
        for n, v in mine:
            key = n.split("(")[0][:90]
            tot.setdefault(key, [0, 0.0])
            tot[key][0] += 1
            tot[key][1] += v
        total = sum(v for _, v in tot.values())
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {c} | {v/1e6:.3f} | {100*v/total:.1f} % |\n")
        f.write("\nLast 13 launches (one chain step):\n\n| # | kernel | µs |\n|---|---|---:|\n")
        for i, (n, v) in enumerate(mine[-13:]):
            f.write(f"| {i} | `{n.split('(')[0][:90]}` | {v/1e3:.1f} |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    units = rows[1]
    kn = h.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({src})\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[kn]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            d = {}
            for k in KEYS:
                if k in h:
                    i = h.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")
                    d[k] = r[i]
            try:
                t = float(d["gpu__time_duration.sum"].replace(",", ""))
                unit_t = units[h.index("gpu__time_duration.sum")]
                scale_t = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit_t, 1e-9)
                rb = float(d["dram__bytes_read.sum"].replace(",", ""))
                wb = float(d["dram__bytes_write.sum"].replace(",", ""))
                ub = units[h.index("dram__bytes_read.sum")]
                scale_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(ub, 1)
                f.write(f"\nDRAM traffic {(rb+wb)*scale_b/1e9:.3f} GB in {t*scale_t*1e3:.3f} ms = "
                        f"{(rb+wb)*scale_b/(t*scale_t)/1e9:.0f} GB/s (under ncu: cold clocks, not a bench number)\n\n")
            except (KeyError, ValueError):
                f.write("\n")
            stalls = []
            for i, k in enumerate(h):
                if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                    try:
                        stalls.append((float(r[i]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                    except ValueError:
                        pass
            tot = sum(v for v, _ in stalls) or 1.0
            f.write("stall samples: " + ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in sorted(stalls, reverse=True)[:7]) + "\n\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
