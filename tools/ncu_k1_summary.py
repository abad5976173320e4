"""Boils an `ncu --page raw --csv` export of the K1 launch down to the numbers bench.py quotes
(profiles/r02_k1_traffic.json).  Usage: ncu_k1_summary.py metrics.csv out.json"""
import csv, json, sys


def main(src, dst):
    rows = list(csv.reader(open(src)))
    names, units, vals = rows[0], rows[1], rows[2]
    m = {}
    for n, u, v in zip(names, units, vals):
        try:
            m[n] = (float(v.replace(",", "")), u)
        except ValueError:
            pass

    def g(name, scale=1.0):
        if name not in m:
            return None
        v, u = m[name]
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u, 1.0)
        return v * mult * scale

    rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
    cyc = g("sm__cycles_elapsed.avg")
    wf = g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    inst = g("sm__inst_executed.sum") or g("smsp__inst_executed.sum")
    out = {
        "source": "ncu --set full --clock-control none, one K1 launch of `python bench.py --steps 1 --warmup 3` (tools/ncu_k1.sh)",
        "kernel_ms_under_ncu": g("gpu__time_duration.sum"),
        "dram_bytes_per_launch": (rd or 0) + (wr or 0),
        "dram_bytes_read": rd, "dram_bytes_written": wr,
        "sm_cycles_elapsed": cyc,
        "smem_wavefronts": wf,
        "smem_wavefronts_per_clk_per_sm": (wf / (cyc * 148)) if wf and cyc else None,
        "smem_wavefront_pipe_pct": g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "smem_bank_conflicts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "issue_slots_active_pct": g("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fmaheavy_pipe_pct": g("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "fp64_pipe_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "lsu_pipe_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "threads_active_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "registers_per_thread": g("launch__registers_per_thread"),
        "sm_clock_ghz_under_ncu": g("sm__cycles_elapsed.avg.per_second"),
    }
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
