"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/ (run in the build
container; ncu reads .ncu-rep files without a GPU).

    python tools/ncu_summary.py launches gpurun_out/r1b_bench_launches.csv profiles/r1b_bench_launches_summary.md "<command>"
    python tools/ncu_summary.py full gpurun_out/r1b_nb2_full.ncu-rep profiles/r1b_nb2_full_summary.md "<command>"
"""
import collections, csv, io, json, re, subprocess, sys

FULL_METRICS = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_static launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
sm__cycles_elapsed.avg.per_second sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__throughput.avg.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
l1tex__t_requests_pipe_lsu_mem_global_op_red.sum l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum""".split()


def launches(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1000 if r["Metric Unit"] == "ns" else (v * 1000 if r["Metric Unit"] == "ms" else v)
        agg.setdefault(re.sub(r"\(.*", "", r["Kernel Name"]), []).append(v)
    own = {k: v for k, v in agg.items() if not k.startswith("void at::")}
    tot = sum(sum(v) for v in own.values())
    out = [f"# ncu launch list of `{cmd}`", "",
           "(raw CSV next to this file; per-launch times are cold-cache and serialised, so only SHARES are compared; torch "
           "fill kernels -- the 256 MiB L2 flush between steps and `force.zero_()` of the e2e leg -- are excluded)", "",
           "| kernel | launches | mean us | share of own-kernel time |", "|---|---:|---:|---:|"]
    for k, v in sorted(own.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / tot:.1f} % |")
    open(dst, "w").write("\n".join(out) + "\n")


def full(src, dst, cmd):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = [f"# `ncu --set full` of `{vals[hdr.index('Kernel Name')]}`", "", f"Command: `{cmd}`", "", "| metric | value |", "|---|---|"]
    for m in FULL_METRICS:
        if m in hdr:
            out.append(f"| `{m}` | {vals[hdr.index(m)]} {units[hdr.index(m)]} |")
    st = [(h, float(vals[i].replace(",", "") or 0)) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    out += ["", "Warps stalled per issued instruction (top reasons):", "", "| reason | warps per issue |", "|---|---:|"]
    for h, v in sorted(st, key=lambda x: -x[1])[:10]:
        out.append(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.2f} |")
    open(dst, "w").write("\n".join(out) + "\n")
    rd = float(vals[hdr.index("dram__bytes_read.sum")].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(vals[hdr.index("dram__bytes_write.sum")].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
    return rd, wr


if __name__ == "__main__":
    kind, src, dst, cmd = sys.argv[1:5]
    if kind == "launches":
        launches(src, dst, cmd)
    else:
        rd, wr = full(src, dst, cmd)
        print(json.dumps({"dram_bytes_read": rd, "dram_bytes_write": wr}))
