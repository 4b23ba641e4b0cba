"""Condense an `ncu --page raw --csv` dump into the per-kernel table kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv && python tools/ncu_summary.py raw.csv > profiles/<name>.csv
"""
import csv
import re
import sys

KEYS = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_throughput_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
]


def kernel_name(full):
    m = re.search(r"(\w+)(<[^>]*>)?\(", full.replace("<unnamed>::", ""))
    return (m.group(1) + (m.group(2) or "")) if m else full[:60]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(k), name) for k, name in KEYS if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([name + (f" [{units[i]}]" if units[i] else "") for i, name in cols])
    for r in rows[2:]:
        w.writerow([kernel_name(r[i]) if name == "kernel" else r[i] for i, name in cols])


if __name__ == "__main__":
    main(sys.argv[1])
