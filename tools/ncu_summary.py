"""Summarise the `ncu --page raw --csv` export of one distance-kernel launch into the JSON that
bench.py reads `roofline.traffic` from (profiles/distance_kernel_ncu_summary.json).

    python tools/ncu_summary.py gpurun_out/TAG_distance_kernel_ncu_raw.csv "what was captured" > profiles/distance_kernel_ncu_summary.json
"""
import csv
import json
import sys

UNIT = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1e-3, 'ms': 1.0, 'ns': 1e-6}


def main():
    path, what = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]

    def g(name, default=None):
        if name not in hdr:
            return default
        i = hdr.index(name)
        try:
            return float(vals[i].replace(',', '')) * UNIT.get(units[i], 1.0)
        except ValueError:
            return default

    out = {
        "capture": "%s (%s; kernel %s)" % (path.replace('gpurun_out/', 'profiles/'), what, vals[hdr.index('Kernel Name')]),
        "dram_bytes_per_launch": (g('dram__bytes_read.sum', 0.0) + g('dram__bytes_write.sum', 0.0)),
        "dram_bytes_read": g('dram__bytes_read.sum'),
        "dram_bytes_written": g('dram__bytes_write.sum'),
        "duration_ms": g('gpu__time_duration.sum'),
        "sm_clock_ghz": g('smsp__cycles_elapsed.avg.per_second'),
        "cycles": g('sm__cycles_elapsed.max'),
        "tensor_pipe_active_pct": g('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'),
        "issue_active_pct": g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
        "warp_instructions": g('smsp__inst_executed.sum'),
        "l2_to_sm_gbytes": (g('l1tex__m_xbar2l1tex_read_bytes.sum', 0.0) or 0.0) / 1e9,
        "l2_hit_pct": g('lts__t_sector_hit_rate.pct'),
        "registers": g('launch__registers_per_thread'),
        "grid": g('launch__grid_size'),
        "ctas_per_sm": g('launch__occupancy_limit_registers'),
    }
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
