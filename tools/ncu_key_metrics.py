"""Condense `ncu -i X.ncu-rep --page raw --csv` files into the handful of numbers quoted in profiles/README.md.
usage: python tools/ncu_key_metrics.py out.json name=file.csv [name=file.csv ...]"""
import csv
import json
import sys

KEYS = {
    "duration": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "tensor_pipe_cycles_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smem_tensor_operand_wavefronts_pct": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "registers_per_thread": "launch__registers_per_thread",
    "grid": "launch__grid_size",
}


def load(fn):
    rows = list(csv.reader(open(fn)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:120]}
        for k, m in KEYS.items():
            c = [i for i, h in enumerate(hdr) if h == m or h.endswith("." + m)]
            if c and r[c[0]] != "":
                d[k] = f"{r[c[0]]} {units[c[0]]}".strip()
        out.append(d)
    return out


if __name__ == "__main__":
    res = {}
    for a in sys.argv[2:]:
        name, fn = a.split("=")
        res[name] = load(fn)
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(res, indent=1))
