"""Build profiles/r01_traffic.json: per-kernel DRAM bytes / DRAM % / tensor-pipe % of one train step.
Inputs: the raw-page CSV of an `ncu --set full` capture of the plane / rollout kernels of ONE step
(`ncu -i X.ncu-rep --page raw --csv`), and a bench.py JSON line of the same code (its `kernel_order` lists the C-ABI calls
of a step in issue order).  Launches are matched to bench keys by issue order within the captured kernel families.
  python profiles/match_traffic.py gpurun_out/r01_full_raw.csv gpurun_out/r01_bench_final.json profiles/r01_traffic.json "<how>"
"""
import csv
import json
import sys

FAMILIES = ("pl_conv_down", "pl_conv_up", "pl_conv_wgrad", "rollout_tc_fwd", "rollout_tc_bwd", "rollout_fwd", "rollout_bwd")


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main(csv_path, bench_path, out_path, how):
    rows = list(csv.reader(l for l in open(csv_path) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, scale=True):
        i = col.get(name)
        if i is None:
            return None
        v = num(r[i])
        if v is None:
            return None
        if scale:
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(units[i], 1.0)
        return v

    text = open(bench_path).read()
    bench = json.loads(text[text.index("{"):])
    keys = []
    for k in bench["kernel_order"]:
        fam = k.split(":")[0]
        if fam in ("rollout_tc_fwd", "rollout_tc_bwd"):
            keys.append("rollout_tc_pack")           # the weight-stream pack kernel precedes each rollout kernel (and matches the -k regex)
        if fam in FAMILIES:
            keys.append(k)
    launches = rows[2:]

    def kind(name_or_key):
        for tag in ("rollout_tc_pack", "rollout_tc_fwd", "rollout_tc_bwd", "wgrad", "rollout"):
            if tag in name_or_key:
                return tag
        return "fwd"

    # the capture is a window of consecutive launches of the cyclic per-step sequence: find its rotation
    n = len(keys)
    rot = [o for o in range(n) if all(kind(launches[i][col["Kernel Name"]]) == kind(keys[(o + i) % n]) for i in range(len(launches)))]
    assert len(rot) == 1, ("cannot align the capture with the step", rot)
    out = {}
    for i, r in enumerate(launches):
        k = keys[(rot[0] + i) % n]
        if k == "rollout_tc_pack":
            continue
        name = r[col["Kernel Name"]]
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        out[k] = {
            "grid": r[col["Grid Size"]] if "Grid Size" in col else None,
            "duration_ms": get(r, "gpu__time_duration.sum"),
            "dram_bytes_per_launch": None if rd is None or wr is None else rd + wr,
            "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_pct_of_peak": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
            "tensor_pipe_pct": get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
            "l2_hit_pct": get(r, "lts__t_sector_hit_rate.pct", False),
            "sm_warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
            "registers": get(r, "launch__registers_per_thread", False),
            "sass_kernel": name[:100],
        }
        e = out[k]
        print(f"{e['duration_ms'] or 0:8.3f} ms  dram {((e['dram_bytes_per_launch'] or 0) / 1e6):9.1f} MB  dram% {e['dram_pct_of_peak']}  tensor% {e['tensor_pipe_pct']}  {k}")
    json.dump({"how": how, "kernels": out}, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:5])
