"""Summarise an `ncu --page raw --csv` export and a launch-list CSV into the small tables kept under profiles/.
Usage: python profiles/summarize_ncu.py <raw.csv> <launches.csv> <out_prefix>"""
import csv
import json
import sys
from collections import OrderedDict, defaultdict

KEYS = OrderedDict([
    ("gpu__time_duration.sum", "time_ms"),
    ("dram__bytes_read.sum", "dram_read_GB"),
    ("dram__bytes_write.sum", "dram_write_GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
])


def main(raw, launches, out):
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = []
    for r in data:
        k = OrderedDict(kernel=r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("(int)", ""))
        for src, dst in KEYS.items():
            if src in idx:
                v = r[idx[src]].replace(",", "")
                try:
                    v = float(v)
                    if units[idx[src]] == "Gbyte" or dst.endswith("_GB"):
                        scale = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}.get(units[idx[src]], 1.0)
                        v *= scale
                    if dst == "time_ms":
                        v *= {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[idx[src]], 1.0)
                except ValueError:
                    pass
                k[dst] = v
        kernels.append(k)
    json.dump(kernels, open(out + "_kernels.json", "w"), indent=1)
    # launch list: per-kernel totals of one iteration
    tot = defaultdict(lambda: [0, 0.0])
    lines = [l for l in csv.reader(open(launches)) if len(l) > 5]
    h = lines[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    for l in lines[1:]:
        name = l[ki].split("(")[0].replace("void ", "").replace("(int)", "")
        t = float(l[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(l[ui], 1e-6)
        tot[name][0] += 1
        tot[name][1] += t
    total = sum(v[1] for v in tot.values())
    with open(out + "_launch_shares.md", "w") as f:
        f.write("| kernel | launches | ms (ncu, cold, serialised) | share |\n|---|---|---|---|\n")
        for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {name} | {c} | {t:.3f} | {t / total:.3f} |\n")
        f.write(f"| total | {sum(v[0] for v in tot.values())} | {total:.3f} | 1.000 |\n")
    print(open(out + "_launch_shares.md").read())
    for k in kernels:
        print({kk: (round(v, 3) if isinstance(v, float) else v) for kk, v in k.items()})


if __name__ == "__main__":
    main(*sys.argv[1:4])
