"""Key metrics per kernel from an `ncu --page raw --csv` export: python tools/ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us", 1e-3), ("launch__registers_per_thread", "regs", 1), ("launch__block_size", "thr", 1), ("launch__grid_size", "grid", 1),
        ("launch__shared_mem_per_block_dynamic", "smemKB", 1 / 1024), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%", 1), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64i%", 1),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%", 1), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("dram__bytes_read.sum", "rdGB", 1e-9), ("dram__bytes_write.sum", "wrGB", 1e-9),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf", 1e-6), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shwave", 1e-6),
        ("smsp__inst_executed_op_local_ld.sum", "LDL", 1e-6), ("smsp__inst_executed_op_local_st.sum", "STL", 1e-6)]
units = rows[1]
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[col["Kernel Name"]][:70]
    out = []
    for m, lab, sc in want:
        if m in col and r[col[m]] not in ("", "n/a"):
            v = float(r[col[m]].replace(",", ""))
            u = units[col[m]]
            if lab == "us":
                v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
                sc = 1
            if lab in ("rdGB", "wrGB"):
                v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            if lab == "smemKB":
                v = v * {"byte": 1, "Kbyte": 1024, "Mbyte": 1024 ** 2}.get(u, 1)
            out.append("%s %.4g" % (lab, v * sc))
    print(name, "|", " ".join(out))
