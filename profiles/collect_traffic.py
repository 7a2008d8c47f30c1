"""Turn an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) into profiles/r01_traffic.json:
per kernel launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), duration, DRAM %, tensor-pipe %, L2 hit rate.
Launches are matched to bench.py's per-call keys by order within each kernel family (the bench issues them in a fixed order)."""
import csv
import json
import sys


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main(csv_path, out_path):
    rows = list(csv.reader(open(csv_path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale_units=True):
        i = col.get(name)
        if i is None:
            return None
        v = num(r[i])
        if v is None:
            return None
        u = units[i]
        if scale_units:
            mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
            v *= mult
        return v

    out = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        out.append({
            "kernel": name[:100],
            "grid": r[col["Grid Size"]] if "Grid Size" in col else None,
            "duration_ms": get(r, "gpu__time_duration.sum"),
            "dram_bytes_per_launch": None if rd is None or wr is None else rd + wr,
            "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_pct_of_peak": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
            "tensor_pipe_pct": get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
            "l2_hit_pct": get(r, "lts__t_sector_hit_rate.pct", False),
            "sm_warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
            "registers": get(r, "launch__registers_per_thread", False),
        })
    json.dump({"source": csv_path, "launches": out}, open(out_path, "w"), indent=1)
    for e in out:
        print(f"{e['duration_ms'] or 0:9.3f} ms  dram {((e['dram_bytes_per_launch'] or 0) / 1e6):10.1f} MB  dram% {e['dram_pct_of_peak']}  tensor% {e['tensor_pipe_pct']}  {e['kernel'][:70]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
