"""Turns an ncu CSV of `bench_micro/pass_once.py` (metrics below, kernel pass_kernel) into profiles/r02_fp64_counts.json,
stamped with the hash of clic_b200/csrc so that bench.py refuses stale counts.
  ncu --metrics <METRICS> --clock-control none --csv --log-file gpurun_out/fp64_counts.csv -k regex:pass_kernel \
      python bench_micro/pass_once.py
  python bench_micro/fp64_counts.py gpurun_out/fp64_counts.csv"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,"
           "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed_pipe_fp64.sum,"
           "smsp__inst_executed_pipe_tensor_subpipe_dmma.sum,smsp__inst_executed.sum,dram__bytes_read.sum,"
           "dram__bytes_write.sum,gpu__time_duration.sum,smsp__pipe_fp64_cycles_active.sum,"
           "smsp__pipe_tensor_subpipe_dmma_cycles_active.sum,smsp__pipe_shared_cycles_active.sum,"
           "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active,smsp__cycles_active.sum,"
           "smsp__cycles_elapsed.sum,sm__cycles_elapsed.avg.per_second,launch__registers_per_thread")


PASS_KERNEL_SOURCES = ("clic_math.cuh", "factor_eval.cuh", "kernels.cuh", "pass_fused.cuh")  # as in bench.py


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(i for i, r in enumerate(rows) if "Metric Name" in r)
    h = rows[hdr]
    ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    launches = {}
    for r in rows[hdr + 1:]:
        if len(r) <= vi or "pass_kernel" not in r[ki]:
            continue
        launches.setdefault(r[idi], {})[r[mi]] = float(r[vi].replace(",", ""))
    last = launches[sorted(launches, key=int)[-1]]  # the last captured launch (warm code, flushed data)
    dp = sum(last.get(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum", 0.0) for o in ("dfma", "dmul", "dadd"))
    pipe_warp = last.get("smsp__inst_executed_pipe_fp64.sum", 0.0)       # scalar FP64 warp instructions
    dmma = last.get("smsp__inst_executed_pipe_tensor_subpipe_dmma.sum", 0.0)  # DMMA warp instructions
    hsh = hashlib.sha256()
    csrc = os.path.join(ROOT, "clic_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if name in PASS_KERNEL_SOURCES:
            hsh.update(open(os.path.join(csrc, name), "rb").read())
    out = {
        "csrc_sha256": hsh.hexdigest(),
        "workload": "C5, one fused pass (pass_kernel<true>), 1 GPU",
        "dp_thread_inst": dp,
        "fp64_pipe_warp_inst": pipe_warp,
        "dmma_warp_inst": dmma,
        "dram_bytes": last.get("dram__bytes_read.sum", 0.0) + last.get("dram__bytes_write.sum", 0.0),
        "ncu_duration_ns": last.get("gpu__time_duration.sum"),
        "pipe_shared_active_pct": last.get("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"),
        "pipe_fp64_cycles_active": last.get("smsp__pipe_fp64_cycles_active.sum"),
        "pipe_dmma_cycles_active": last.get("smsp__pipe_tensor_subpipe_dmma_cycles_active.sum"),
        "pipe_shared_cycles_active": last.get("smsp__pipe_shared_cycles_active.sum"),
        "smsp_cycles_elapsed": last.get("smsp__cycles_elapsed.sum"),
        "n_smsp": 592, "sm_clock_hz": last.get("sm__cycles_elapsed.avg.per_second", 1.965e9),
        "fp64_pipe_peak": {"dfma_tflops": 36.3, "dmma_tflops": 37.1, "source": "profiles/r01_fp64_peaks.txt"},
        "raw": last,
    }
    with open(os.path.join(ROOT, "profiles", "r02_fp64_counts.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "raw"}, indent=1))


if __name__ == "__main__":
    main()
