"""Per-kernel counters of one `ncu --set full` capture of a precompute step -> profiles/kernel_counters.json.

    ncu -i gpurun_out/prof_X.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_counters.py /tmp/raw.csv <segments per launch> "<source note>" [out.json]

For every profiled launch: duration, DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and the FP64 thread
instructions executed (smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on, reported by ncu as a rate per
elapsed cycle, multiplied back by the elapsed SM cycles).  FP64 flop = 2 dfma + dmul + dadd.  Launches of the same kernel
are summed per step.  bench.py reads the file for `roofline.traffic`, `roofline.traffic_total` and `roofline.fp64`
(flop per segment is a property of the code; the achieved rate uses the live CUDA-event times).
"""
import csv
import json
import re
import sys

SHORT = re.compile(r"(?:void\s+)?(?:bpc::)?(k_[a-z0-9_]+)")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    raw, segs = sys.argv[1], int(sys.argv[2])
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    out = sys.argv[4] if len(sys.argv) > 4 else "profiles/kernel_counters.json"
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name):
        return num(r[col[name]]) if name in col else 0.0

    def scaled(r, name):                      # ncu prints bytes with a unit column (byte / Kbyte / Mbyte / Gbyte)
        v = get(r, name)
        u = units[col[name]].lower() if name in col else ""
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3}.get(u, 1.0)

    kernels = {}
    for r in rows[2:]:
        m = SHORT.search(r[col["Kernel Name"]])
        if not m:
            continue
        k = m.group(1)
        cyc = get(r, "sm__cycles_elapsed.max") or get(r, "smsp__cycles_elapsed.max")
        tunit = units[col["gpu__time_duration.sum"]].lower()
        ms = get(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(tunit, 1e-6)
        e = kernels.setdefault(k, {"launches_per_step": 0, "ncu_ms": 0.0, "dram_bytes": 0.0, "dfma": 0.0, "dmul": 0.0,
                                   "dadd": 0.0, "fp64_pipe_pct_x_ms": 0.0, "regs": 0})
        e["launches_per_step"] += 1
        e["ncu_ms"] += ms
        e["dram_bytes"] += scaled(r, "dram__bytes_read.sum") + scaled(r, "dram__bytes_write.sum")
        for op in ("dfma", "dmul", "dadd"):
            e[op] += get(r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") * cyc
        e["fp64_pipe_pct_x_ms"] += get(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") * ms
        e["regs"] = max(e["regs"], int(get(r, "launch__registers_per_thread")))
    res = {"source": note, "segments_per_launch": segs, "kernels": {}}
    tot_b = tot_f = tot_ms = 0.0
    for k, e in kernels.items():
        flop = 2 * e["dfma"] + e["dmul"] + e["dadd"]
        res["kernels"][k] = {
            "launches_per_step": e["launches_per_step"], "ncu_ms": round(e["ncu_ms"], 5), "regs": e["regs"],
            "dram_bytes_per_step": e["dram_bytes"], "dram_bytes_per_segment": e["dram_bytes"] / segs,
            "fp64_thread_inst_per_segment": {op: e[op] / segs for op in ("dfma", "dmul", "dadd")},
            "fp64_flop_per_segment": flop / segs,
            # share of the FP64 pipe's issue slots: every FP64 thread instruction (FMA or not) takes one of 64 lanes/clk/SM
            "fp64_pipe_busy_pct": round(e["fp64_pipe_pct_x_ms"] / e["ncu_ms"], 2) if e["ncu_ms"] else 0.0,
        }
        tot_b += e["dram_bytes"]; tot_f += flop; tot_ms += e["ncu_ms"]
    res["total_dram_bytes_per_step"] = tot_b
    res["total_dram_bytes_per_segment"] = tot_b / segs
    res["total_fp64_flop_per_segment"] = tot_f / segs
    res["total_fp64_inst_per_segment"] = sum(e["dfma"] + e["dmul"] + e["dadd"] for e in kernels.values()) / segs
    res["sum_ncu_ms"] = tot_ms
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: (round(v["ncu_ms"], 3), round(v["fp64_flop_per_segment"] / 1e6, 3), v["fp64_pipe_busy_pct"],
                          round(v["dram_bytes_per_segment"] / 1e3, 1)) for k, v in res["kernels"].items()}, indent=1))
    print("per segment: FP64 MFLOP", tot_f / segs / 1e6, "DRAM KB", tot_b / segs / 1e3, "sum ms", tot_ms)


if __name__ == "__main__":
    main()
