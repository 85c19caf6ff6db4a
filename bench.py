#!/usr/bin/env python
"""bench.py -- the per-line VCF transform on the BASELINE.json workload (configs[1]: synthetic 1000G Phase 3
chr1-shape VCF, 6.2M variants x 2,504 phased diploid samples).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, libbvcf)
  python bench.py --impl reference --gpus N ...            the reference's CPU implementation of the path
                                                           (oracle port: Go cannot be built in this image)

A step = one pass of the whole pipeline over one rank's shard, input resident in HBM (`value`), and the
same metric through the C ABI with host buffers, H2D + D2H inside the timed region (`e2e`).
Multi-GPU: one process per GPU (torchrun), shards are independent line ranges, no data-path collective.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_LINES = int(os.environ.get("BVCF_BENCH_LINES", "6200000"))
N_SAMPLES = 2504
SEED = 20130502
METRIC = "variants/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe):
    one long-running `nvidia-smi -lms 100` started before the region and stopped after it."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._p = None

    def __enter__(self):
        try:
            self._p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                        "--format=csv,noheader,nounits", "-lms", "100"],
                                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.25)  # let the first sample land before the timed region starts
        except Exception:
            self._p = None
        return self

    def __exit__(self, *a):
        if self._p is None:
            return
        time.sleep(0.12)
        self._p.terminate()
        try:
            out, _ = self._p.communicate(timeout=5)
        except Exception:
            self._p.kill()
            out = ""
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                self.rows.append(f)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def scan_traffic_ratio():
    """DRAM bytes (read+write) per algorithmic byte of the scan kernel, from the committed ncu --set full
    capture (profiles/r01_scan_traffic.json); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "r01_scan_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["algorithmic_bytes"], d.get("source", p)
    except Exception:
        return None, None


def newline_cuts(view, total: int, chunk: int):
    """Newline-aligned chunk boundaries of a host buffer (numpy uint8 view)."""
    import numpy as np

    cuts = [0]
    while cuts[-1] < total:
        end = min(cuts[-1] + chunk, total)
        if end < total:
            lo = max(cuts[-1], end - (1 << 20))
            nl = np.flatnonzero(view[lo:end] == 10)
            end = lo + int(nl[-1]) + 1
        cuts.append(end)
    return cuts


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from bystro_vcf_b200 import Config, Transformer, synth
    from bystro_vcf_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local
    n_lines = args.lines
    first_line = rank * n_lines  # rank r owns lines [r*n, (r+1)*n): newline-aligned shards of one N*n-line file

    cfg = Config()
    cfg.allowedFilters = {"PASS": True, ".": True}
    cfg.device = dev
    tr = Transformer(cfg, eol_width=1, n_slots=3, max_chunk_bytes=args.chunk_mb << 20)
    tr.set_header(synth.chrom_line(SEED, N_SAMPLES))

    # ---- workload: generated on the device, straight into the resident input region ----
    _, need = synth.device_lines(SEED, N_SAMPLES, "chr1", first_line, n_lines, 0, 0, dev)
    out_cap = int(need * 0.12) + (64 << 20)
    d_in, d_out = tr.resident_alloc(need, out_cap)
    got, _ = synth.device_lines(SEED, N_SAMPLES, "chr1", first_line, n_lines, d_in, need, dev)
    assert got == need, "device generator failed"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident steps ----
    for _ in range(args.warmup):
        stats, times = tr.resident_run(need)
    assert stats["n_lines"] == n_lines and stats["n_records"] == n_lines, stats
    launches0 = tr.launches
    acc = {"scan_ms": 0.0, "compact_ms": 0.0, "stats_ms": 0.0, "rows_ms": 0.0, "names_ms": 0.0, "total_ms": 0.0}
    barrier()
    with ClockSampler(dev) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            stats, times = tr.resident_run(need)
            for k in acc:
                acc[k] += times[k]
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    gpu_launches = tr.launches - launches0
    dev_ms = acc["total_ms"]  # CUDA events on the launching stream, first to last launch of every step
    out_bytes = stats["out_bytes"]
    n_rows = stats["n_rows"]

    # ---- end to end through the C ABI: pinned host buffers, H2D + kernels + D2H per chunk ----
    e2e_lines = min(n_lines, args.e2e_lines)
    # the slice = first e2e_lines lines of this rank's shard, copied once from the device to pinned host memory
    probe = min(need, int(e2e_lines * 10400))
    host_ptr = C.c_void_p()
    _lib.check(_lib.lib().bvcf_host_alloc(C.byref(host_ptr), probe), None, "bvcf_host_alloc")
    tr.resident_peek(0, probe, host_ptr.value)
    hview = np.ctypeslib.as_array(C.cast(host_ptr, C.POINTER(C.c_uint8)), shape=(probe,))
    nl = np.flatnonzero(hview[max(0, probe - (1 << 20)):probe] == 10)
    e2e_bytes = max(0, probe - (1 << 20)) + int(nl[-1]) + 1
    cuts = newline_cuts(hview, e2e_bytes, args.chunk_mb << 20)
    # PCIe H2D bound, measured in the same run: pinned -> device copies of the slice
    tr.resident_upload(0, (host_ptr.value, min(e2e_bytes, 1 << 30)))  # warm
    tp0 = time.perf_counter()
    n_up = 3
    for _ in range(n_up):
        tr.resident_upload(0, (host_ptr.value, e2e_bytes))
    pcie_gbs = n_up * e2e_bytes / (time.perf_counter() - tp0) / 1e9
    e2e_ms = None
    e2e_nlines = 0
    e2e_out = 0
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            barrier()
            t0 = time.perf_counter()
            e2e_nlines = 0
            e2e_out = 0
        nchunks = len(cuts) - 1
        sub = col = 0
        L = _lib.lib()
        tsv, n = C.c_void_p(), C.c_size_t()
        st = _lib.CChunkStats()
        while col < nchunks:
            while sub < nchunks and sub - col < tr.n_slots:
                _lib.check(L.bvcf_submit(tr._ctx, sub, host_ptr.value + cuts[sub], cuts[sub + 1] - cuts[sub]), tr._ctx, "submit")
                sub += 1
            _lib.check(L.bvcf_collect(tr._ctx, col, C.byref(tsv), C.byref(n), None, None, None, C.byref(st)), tr._ctx, "collect")
            _lib.check(L.bvcf_release(tr._ctx, col), tr._ctx, "release")
            e2e_nlines += st.n_lines
            e2e_out += n.value
            col += 1
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e_lines_per_step = e2e_nlines // args.steps
    e2e_out_per_step = e2e_out // args.steps

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, e2e_ms = (float(x) for x in t.tolist())

    result = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        value = world * n_lines * args.steps / (dev_ms / 1e3)
        in_gbs = world * need * args.steps / (dev_ms / 1e3) / 1e9
        scan_ms = acc["scan_ms"] / args.steps
        alg_bytes = need + out_bytes
        # dominant kernel = the fused index+genotype scan: it must read every input byte once
        roof = {"bound": "hbm", "kernel": "bvcf_scan_genotype_kernel", "achieved": need / (scan_ms / 1e3) / 1e9,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": None,
                "algorithmic_bytes_per_step": need, "avg_ms_per_step": scan_ms}
        roof["frac"] = roof["achieved"] / peak
        # the kernel is launched once per 16 GiB sub-chunk of the resident input: per-launch figures beside the per-step ones
        n_launch = max(1, -(-need // (16 << 30)))
        roof["launches_per_step"] = n_launch
        roof["algorithmic_bytes_per_launch"] = need / n_launch
        roof["avg_ms_per_launch"] = scan_ms / n_launch
        ratio, src = scan_traffic_ratio()
        if ratio is not None:  # DRAM read+write bytes per launch, scaled from the committed ncu --set full capture
            roof["traffic"] = ratio * need / n_launch
            roof["traffic_per_step"] = ratio * need
            roof["traffic_source"] = src
        pipe = {"achieved": alg_bytes / (dev_ms / args.steps / 1e3) / 1e9, "unit": "GB/s",
                "algorithmic_bytes_per_step": alg_bytes, "bytes_per_variant": alg_bytes / n_lines}
        pipe["frac"] = pipe["achieved"] / peak
        e2e_val = world * e2e_lines_per_step * args.steps / (e2e_ms / 1e3)
        result = {
            "metric": METRIC, "value": value, "unit": "variants/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic 1000G Phase 3 chr1-shape VCF, %d variants x %d phased diploid "
                                   "samples per GPU (seed %d), default PASS/. filter" % (n_lines, N_SAMPLES, SEED),
                       "variants_per_gpu": n_lines, "samples": N_SAMPLES, "input_bytes_per_gpu": need,
                       "output_bytes_per_gpu": out_bytes, "rows_per_gpu": n_rows,
                       "l2": "input (%.1f GB) >> 126 MB L2, no flush needed" % (need / 1e9),
                       "parallelism": "%d independent newline-aligned shards, no collective" % world},
            "input_gb_per_s": in_gbs,
            "wall_ms_per_step": wall_ms / args.steps,
            "kernel_ms_per_step": {k: v / args.steps for k, v in acc.items()},
            "roofline": roof, "roofline_pipeline": pipe,
            "e2e": {"value": e2e_val, "unit": "variants/s", "h2d_bytes_per_step": e2e_bytes,
                    "d2h_bytes_per_step": e2e_out_per_step, "variants_per_step": e2e_lines_per_step,
                    "input_gb_per_s": world * e2e_bytes * args.steps / (e2e_ms / 1e3) / 1e9,
                    "pcie_h2d_gbs_measured": pcie_gbs,
                    "frac_of_pcie_h2d_bound": (e2e_bytes * args.steps / (e2e_ms / 1e3) / 1e9) / pcie_gbs,
                    "sample": "first %d variants of each rank's shard, pinned host memory, %d MiB chunks, 3 slots"
                              % (e2e_lines_per_step, args.chunk_mb)},
            "gpu_launches": int(gpu_launches),
            "clocks": clk.summary(),
        }
        # ---- CPU baseline beside it (N=1 only): the oracle port on a bounded sample ----
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O

            cb_lines = min(e2e_lines_per_step, args.cpu_lines)
            # newline-aligned prefix holding cb_lines lines
            idx = np.flatnonzero(hview[:min(e2e_bytes, int(cb_lines * 10400))] == 10)
            cb_lines = min(cb_lines, len(idx))
            cb_bytes = int(idx[cb_lines - 1]) + 1
            threads = os.cpu_count() or 1
            t0 = time.perf_counter()
            r = O.process_block(O.OracleConfig(), synth.chrom_line(SEED, N_SAMPLES), host_ptr.value, cb_bytes, 1, threads)
            dt = time.perf_counter() - t0
            result["cpu_baseline"] = {"value": r.n_lines / dt, "unit": "variants/s", "cores": threads, "kind": "port",
                                      "sample": "first %d variants of the workload, in RAM, output to memory; %.2f s"
                                                % (r.n_lines, dt),
                                      "note": "C restatement of main.go (oracle/), not the Go binary: no Go toolchain in this image"}
    _lib.lib().bvcf_host_free(host_ptr)
    tr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


def run_reference(args):
    """The reference's CPU implementation of the path on the host cores: the oracle port (cpu_baseline.kind
    "port"), all host threads, a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from bystro_vcf_b200 import synth
    from oracle import oracle as O

    threads = os.cpu_count() or 1
    n = min(args.lines, args.ref_lines)
    body = synth.host_lines(SEED, N_SAMPLES, "chr1", 0, n, threads)
    chrom = synth.chrom_line(SEED, N_SAMPLES)
    cfg = O.OracleConfig()
    for _ in range(args.warmup):
        O.process_block(cfg, chrom, body, threads=threads)
    t0 = time.perf_counter()
    rows = 0
    for _ in range(args.steps):
        r = O.process_block(cfg, chrom, body, threads=threads)
        rows = r.n_rows
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "first %d variants of the workload per step (%.1f MB), in RAM, rows to memory" % (n, len(body) / 1e6)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "variants/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic 1000G Phase 3 chr1-shape VCF, %d variants x %d phased diploid "
                               "samples per GPU (seed %d), default PASS/. filter" % (args.lines, N_SAMPLES, SEED),
                   "samples": N_SAMPLES},
        "cpu_baseline": {"value": value, "unit": "variants/s", "cores": threads, "kind": "port", "sample": sample,
                         "rows_per_step": rows,
                         "note": "oracle/ C restatement of main.go; the Go reference cannot be built here (no Go toolchain)"},
        "e2e": {"value": value, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lines", type=int, default=DEFAULT_LINES, help="variants per GPU (default: the full 6.2M)")
    ap.add_argument("--e2e-lines", type=int, default=400000, help="variants in the host-resident e2e slice")
    ap.add_argument("--cpu-lines", type=int, default=200000, help="variants in the CPU-baseline sample")
    ap.add_argument("--ref-lines", type=int, default=100000, help="variants per step of the reference arm")
    ap.add_argument("--chunk-mb", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
