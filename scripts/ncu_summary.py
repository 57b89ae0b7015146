"""Condense `ncu --page raw --csv` output into a short per-kernel summary (profiles/*.txt)."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second", "TMA load rate"),
    ("gpc__cycles_elapsed.avg.per_second", "SM clock"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = name.split("(")[0][-90:]
        print(f"kernel: {short}")
        if "FieldCfg" in name:
            print("   template:", name[name.index("FieldCfg"):][:60])
        for k, label in KEYS:
            if k in ix and r[ix[k]] != "":
                print(f"   {label:24s} {r[ix[k]]} {units[ix[k]]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
