"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py <rep> [kernel-substring]
Prints per captured launch: duration, DRAM bytes, achieved DRAM GB/s, issue rate, pipe utilisation and
the warp-stall breakdown (cycles per issued instruction)."""
import csv, io, subprocess, sys

def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    def g(r, name, default=float("nan")):
        i = col.get(name)
        if i is None or r[i] == "":
            return default
        try:
            return float(r[i].replace(",", ""))
        except ValueError:
            return r[i]
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if filt and filt not in name:
            continue
        dur = g(r, "gpu__time_duration.sum"); du = units[col["gpu__time_duration.sum"]]
        scale = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(du, 1e-6)
        def bytes_of(n):
            v = g(r, n); u = units[col[n]]
            return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
        print(f"kernel: {name[:110]}")
        print(f"  grid {g(r,'launch__grid_size'):.0f} x block {g(r,'launch__block_size'):.0f}, regs/thread {g(r,'launch__registers_per_thread'):.0f}, "
              f"duration {dur:.1f} {du}, SM clock {g(r,'sm__cycles_elapsed.avg.per_second'):.3f} GHz")
        print(f"  dram read {rd/1e6:.1f} MB + write {wr/1e6:.1f} MB = {(rd+wr)/1e6:.1f} MB -> {(rd+wr)/(dur*scale)/1e9:.0f} GB/s "
              f"(dram throughput {g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f}% of peak)")
        print(f"  issue slots busy {g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}%  "
              f"fma-heavy pipe {g(r,'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f}% of elapsed  "
              f"alu pipe {g(r,'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):.1f}%  "
              f"fma inst {g(r,'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.1f}%  "
              f"lsu {g(r,'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f}%")
        print(f"  warps active/SMSP {g(r,'smsp__warps_active.avg.per_cycle_active'):.2f}, inst executed {g(r,'smsp__inst_executed.sum'):.3g}, "
              f"local loads {g(r,'smsp__sass_inst_executed_op_local_ld.sum',0):.0f} stores {g(r,'smsp__sass_inst_executed_op_local_st.sum',0):.0f}, "
              f"smem bank conflicts {g(r,'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',0):.0f}")
        stalls = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                stalls.append((g(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        stalls.sort(reverse=True)
        print("  stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in stalls if v >= 0.03))

if __name__ == "__main__":
    main()
