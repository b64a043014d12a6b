#!/usr/bin/env python
"""ncu report -> one CSV row per profiled launch with the metrics the roofline discussion uses.
usage: ncu_summary.py report.ncu-rep [out.csv]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"), ("sm__cycles_elapsed.avg", "sm_cycles"),
    ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local_ld"), ("smsp__sass_inst_executed_op_local_st.sum", "local_st"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = io.StringIO()
    w = csv.writer(out)
    w.writerow([name + (f" [{units[idx[m]]}]" if m in idx and units[idx[m]] else "") for m, name in WANT])
    for r in data:
        vals = []
        for m, _ in WANT:
            v = r[idx[m]] if m in idx else ""
            if m == "Kernel Name":
                v = v.replace("void dsg::", "").split("(")[0]
            vals.append(v)
        w.writerow(vals)
    text = out.getvalue()
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
