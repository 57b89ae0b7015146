"""Summarise an `ncu --set full` report (run here, no GPU needed): one block per captured launch with the metrics the
rooflines use, and profiles/ncu_traffic.json = DRAM bytes (read + written) per launch, keyed by the bench's stage names
(bench.py reads it for `roofline.traffic`).

    python scripts/ncu_summary.py gpurun_out/r2_render.ncu-rep profiles/r2_render_ncu_summary.txt profiles/ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("cluster", "launch__cluster_size"),
    ("regs/thread", "launch__registers_per_thread"), ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue active %", "smsp__issue_active.avg.pct"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("shared-memory data pipe (LSU wavefronts) %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("shared-memory bank conflicts (ld / st)", None),
    ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("dram throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("SM clock", "sm__cycles_elapsed.avg.per_second"),
]


def main(rep, out_txt, out_json):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"ncu --set full --clock-control none, report {rep.split('/')[-1]} (scripts/ncu_render_capture.sh: one rendered 512x512 frame,",
             "audio/person_2_auto: coarse field, composite, sample_pdf+merge, fine field, composite; single-kernel replays, not power-capped)", ""]
    traffic = {}
    seen = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        lines.append(f"kernel: {name[:150]}")
        for label, key in METRICS:
            if key is None:
                ld = r[col["l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"]]
                st = r[col["l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"]]
                lines.append(f"   {label:44s} {ld} / {st}")
                continue
            if key in col:
                lines.append(f"   {label:44s} {r[col[key]]} {units[col[key]]}")
        rd, wr = float(r[col["dram__bytes_read.sum"]]), float(r[col["dram__bytes_write.sum"]])
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd * scale[units[col["dram__bytes_read.sum"]]] + wr * scale[units[col["dram__bytes_write.sum"]]]
        k = "field" if "field_fwd" in name else ("composite" if "composite" in name else ("sample_pdf_merge" if "sample_pdf" in name else None))
        if k:
            n = seen.get(k, 0)
            seen[k] = n + 1
            key = {("field", 0): "field_fwd_kernel_coarse", ("field", 1): "field_fwd_kernel_fine", ("composite", 0): "composite_coarse",
                   ("composite", 1): "composite_fine", ("sample_pdf_merge", 0): "sample_pdf_merge"}.get((k, n))
            if key and tot == tot:          # (nan: counters missing for that launch)
                traffic[key] = tot
        lines.append("")
    open(out_txt, "w").write("\n".join(lines))
    json.dump(traffic, open(out_json, "w"), indent=1)
    print("\n".join(lines))
    print(traffic)


if __name__ == "__main__":
    main(*sys.argv[1:4])
