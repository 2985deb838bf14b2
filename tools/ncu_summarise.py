#!/usr/bin/env python
"""Summarise an `ncu --set full` report of one training step into the CSV / JSON kept under profiles/.

    python tools/ncu_summarise.py gpurun_out/r2_c3_step.ncu-rep profiles/r2_step.csv [profiles/r2_traffic.json]

Runs here (no GPU needed): `ncu -i <rep> --page raw --csv`.  The JSON records the SHA-256 of csrc/gemm_tc.cu and the
commit the capture was taken at, so bench.py only quotes `roofline.traffic` while the profiled kernel is the shipped one."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "duration_us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
        ("smsp__issue_active.avg.pct", "issue_active_pct"), ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
        ("launch__registers_per_thread", "regs"), ("sm__cycles_elapsed.avg.per_second", "sm_clock_GHz")]


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    names = {"gemm_tc_kernel<256, 1, 0": "fwd", "gemm_tc_kernel<16, 1, 0": "fwd_last", "gemm_tc_kernel<256, 0, 0": "wgrad",
             "gemm_tc_kernel<256, 1, 1": "bwd", "output_layer": "output", "skinny_bwd": "skinny_bwd (delta + gradient of the 10-unit layer)",
             "momentum_update": "update (per layer, second stream)", "mlp3_kernel": "mlp3 step", "mlp3_reduce_update": "mlp3 reduce+update"}
    lines, total, per_launch = [], 0.0, {}
    for r in rows[2:]:
        k = r[idx["Kernel Name"]]
        label = next((v for p, v in names.items() if p in k), k[:40])
        vals = []
        for m, _ in COLS:
            v = r[idx[m]] if m in idx else ""
            u = units[idx[m]] if m in idx else ""
            try:
                f = float(v)
                if u == "Gbyte": f *= 1000.0
                if u == "Kbyte": f /= 1000.0
                if u == "byte": f /= 1e6
                if u == "ms": f *= 1000.0
                if u == "ns": f /= 1000.0
                if u == "cycle/nsecond": pass
                if u == "cycle/usecond": f /= 1000.0
                vals.append(f"{f:.1f}" if m != "sm__cycles_elapsed.avg.per_second" else f"{f:.3f}")
            except ValueError:
                vals.append(v)
        total += float(vals[0])
        tag = f"{label}{len([x for x in lines if x.startswith(label)])}"
        per_launch[tag] = int((float(vals[3]) + float(vals[4])) * 1e6)
        lines.append(",".join([tag, '"' + k.replace('"', "'")[:60] + '"', r[idx["launch__grid_size"]]] + vals))
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    with open(out_csv, "w") as f:
        f.write(f"# ncu --set full --clock-control none, one training step, report {os.path.basename(rep)}, summarised at commit {commit}\n")
        f.write("launch,kernel,grid," + ",".join(c for _, c in COLS) + "\n")
        f.write("\n".join(lines) + "\n")
        f.write(f"# sum of the launches: {total:.1f} us (cold caches, serialised)\n")
    if out_json:
        tc = {k: v for k, v in per_launch.items() if k.startswith(("fwd", "wgrad", "bwd"))}
        sha = hashlib.sha256(open(os.path.join(ROOT, "cublas-ml_b200", "csrc", "gemm_tc.cu"), "rb").read()).hexdigest()[:16]
        json.dump({"source": f"{os.path.basename(out_csv)} (ncu --set full --clock-control none, one C3 training step)", "commit": commit,
                   "gemm_tc_cu_sha16": sha,
                   "gemm_tc_kernel": {"launches_per_step": len(tc), "dram_bytes_per_launch_mean": int(sum(tc.values()) / max(len(tc), 1)),
                                      "dram_bytes_per_step": int(sum(tc.values())), "per_launch": tc},
                   "other": {k: v for k, v in per_launch.items() if k not in tc}}, open(out_json, "w"), indent=1)
    print(open(out_csv).read())


if __name__ == "__main__":
    main()
