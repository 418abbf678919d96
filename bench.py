#!/usr/bin/env python
"""bench.py — erased-cells per-cell compute path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[1], the config the metric is quoted on):
    CellBuffer::convert sweep over all 10x10 CellType pairs on 8192^2-cell buffers
    (reference src/buffer.rs:150-167): 59 pairs must return NarrowingError without launching, 10 are
    clones, 31 are cast kernels; algorithmic bytes per cell = size_of(S) + size_of(D).
One "step" = one pass of the whole sweep. At N GPUs every rank runs the sweep on its own row strip of
an (N*8192) x 8192 raster (weak scaling; the path shards with no data-path collective). The other
BASELINE configs are measured once each and reported under "configs" (not the headline value); the
reductions there finish with an NCCL all-reduce when N > 1.

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference ...                   # the reference's CPU algorithm (oracle port) on host cores

Prints ONE JSON line on rank 0.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIDE = 8192
CT_NAMES = ["UInt8", "UInt16", "UInt32", "UInt64", "Int8", "Int16", "Int32", "Int64", "Float32", "Float64"]
CT_SIZE = [1, 2, 4, 8, 1, 2, 4, 8, 4, 8]
METRIC = "Gcells/s (CellBuffer ops; HBM GB/s and % of 8 TB/s under roofline)"
NOMINAL_GBS = 8000.0
WORKLOAD = ("CellBuffer::convert sweep over all 10x10 CellType pairs (31 casts + 10 clones + 59 NarrowingError) "
            "on 8192^2-cell buffers, one row strip per GPU")


def workload_config(cells):
    """`config` of the JSON line — the same object from both arms (--impl ours / reference)."""
    return {"workload": WORKLOAD, "cells_per_buffer": cells, "legal_pairs": 41,
            "l2": "inputs + outputs of a step (23.3 GB) >> the 126 MB L2; dst-major order, a source is re-read after >= 1 GB of other traffic"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def legal_pairs(can_fit):
    """dst-major order: consecutive launches read different sources, so a source is re-read only after
    >= 1 GB of other traffic (inputs + outputs >> the 126 MB L2)."""
    return [(s, d) for d in range(10) for s in range(10) if can_fit(s, d)]


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled from a thread (the timed
    region is tens of ms, too short for nvidia-smi -lms), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.samples, self.reason_bits, self.run = [], 0, True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while self.run:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.001)

    def _stop_nvml(self):
        nv = self.nvml
        self.run = False
        self.thread.join(1.0)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}
        sm = [s[0] for s in self.samples]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "power_w_max": max([s[1] for s in self.samples], default=None),
                "samples": len(sm), "reasons": sorted(n for b, n in names.items() if self.reason_bits & b), "source": "nvml"}

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            return self._stop_nvml()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# =================================================================================================
# reference arm: the reference's CPU algorithm (oracle port: per-cell tagged dispatch) on host cores
# =================================================================================================
def run_reference(args, out):
    """The reference's CPU algorithm for the same workload: oracle port, faithful per-cell tagged path
    (reference src/buffer.rs:150-167 through src/value.rs:74-98), on all host threads as independent row strips
    (the reference itself is single-threaded: src/buffer.rs:324-329 has no threads). Each step converts a bounded sample
    of every source buffer (the path streams: Gcells/s does not depend on the buffer length); a 1-thread figure is
    printed beside it because that is what the unmodified crate would do."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    orc.lib()
    from erased_cells_b200 import synth  # host generator only (numpy)
    threads = os.cpu_count() or 1
    per_thread = 1 << 17  # cells per thread per pair: 41 pairs * 131072 cells ~ 0.3 s per step per core
    pairs = legal_pairs(orc.can_fit_into)
    srcs = [[synth.host(ct, per_thread, 0xEC10 + ct, index_offset=t * per_thread) for ct in range(10)] for t in range(threads)]

    def work(t):
        for s, d in pairs:
            orc.convert(srcs[t][s], d)
        for s in range(10):
            for d in range(10):
                if not orc.can_fit_into(s, d):
                    try:
                        orc.convert(srcs[t][s][:1], d)
                    except orc.NarrowingError:
                        pass

    def step(n_threads):
        ths = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
        [t.start() for t in ths]
        [t.join() for t in ths]

    for _ in range(args.warmup):
        step(threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(threads)
    dt = time.perf_counter() - t0
    cells = len(pairs) * per_thread * threads
    value = cells * args.steps / dt / 1e9
    t1 = time.perf_counter()
    step(1)
    one_core = len(pairs) * per_thread / (time.perf_counter() - t1) / 1e9
    sample = (f"{len(pairs)} legal pairs x {per_thread} cells x {threads} threads per step: the first cells of each thread's row strip of the "
              f"8192^2 buffers, same seeds as the GPU workload")
    out.emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gcells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8..f64 casts (integer/f32/f64)", "data": "synthetic",
        "config": workload_config(SIDE * SIDE),
        "cpu_baseline": {"value": value, "unit": "Gcells/s", "cores": threads, "kind": "port", "sample": sample,
                         "one_core_value": one_core, "sample_cells_per_pair_per_step": per_thread * threads},
        "e2e": {"value": value, "unit": "Gcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Rust (no toolchain in this image): timed the C++ restatement oracle/, faithful per-cell tagged path, "
                "one independent strip per host thread (the reference itself is single-threaded: one_core_value)",
    })


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so the pinned host buffers of the e2e leg are
    allocated on the GPU's NUMA node (torchrun does not bind ranks; cross-socket DMA halves PCIe throughput)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


# =================================================================================================
# this repo's arm
# =================================================================================================
def run_ours(args, out):
    import torch
    import torch.distributed as dist

    import erased_cells_b200 as ec
    from erased_cells_b200 import CellBuffer, CellType, Mask, MaskedCellBuffer, NoData, synth

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: erased_cells_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    L = ec.lib()
    ec._lib.check(L.ec_init(local))
    # One explicit (non-default) stream for torch events, NCCL collectives and this library's kernels.
    # (torch's default stream has handle 0, which ec_set_stream reads as "use the library's own stream".)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ec._lib.check(L.ec_set_stream(C.c_void_p(stream.cuda_stream)))
    assert L.ec_get_stream() == stream.cuda_stream
    bind_to_gpu_numa_node(local)
    L.ec_set_min_max_cache(0)  # every timed min_max below reads its raster from HBM (a buffer would otherwise remember the result)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version/debug lines go to stderr: stdout is the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    info = ec._lib.DeviceInfo()
    ec._lib.check(L.ec_device_info_get(C.byref(info)))
    cells = args.cells
    pairs = legal_pairs(lambda s, d: CellType(s).can_fit_into(CellType(d)))
    illegal = [(s, d) for d in range(10) for s in range(10) if not CellType(s).can_fit_into(CellType(d))]
    pair_bytes = {(s, d): (CT_SIZE[s] + CT_SIZE[d]) * cells for s, d in pairs}
    step_bytes = sum(pair_bytes.values())
    # row strip `rank` of the (world*8192) x 8192 raster: seeds per cell type, index offset per strip
    srcs = [synth.device(CellType(ct), cells, 0xEC10 + ct, index_offset=rank * cells) for ct in range(10)]

    def sweep(events=None):
        # events: len(pairs)+1 CUDA events, one at every launch boundary (launch i runs between events i and i+1)
        if events is not None:
            events[0].record()
        for i, (s, d) in enumerate(pairs):
            out = srcs[s].convert(CellType(d))
            if events is not None:
                events[i + 1].record()
            del out
        for s, d in illegal:
            try:
                srcs[s].convert(CellType(d))
                raise AssertionError("illegal convert did not fail")
            except ec.NarrowingError:
                pass

    # ---- parity of what is about to be timed: two 64 Ki-cell windows of every legal pair against the oracle -----------
    parity = {}
    from oracle import oracle as orc  # the checker (test infrastructure), never the thing measured
    orc.build()
    win = 1 << 16
    ok = True
    for s_, d_ in pairs:
        for w0 in (0, cells - win):
            got = srcs[s_].view(w0, win).convert(CellType(d_)).to_vec()
            want = orc.tight_convert(synth.host(CellType(s_), win, 0xEC10 + s_, index_offset=rank * cells + w0), d_)
            ok &= bool(got.dtype == want.dtype and np.array_equal(got.view(np.uint8), want.view(np.uint8)))
    parity["convert_sweep_sampled_windows_vs_oracle"] = ok

    # ---- device-resident timing ------------------------------------------------------------------
    # Region A (value): K steps bracketed by barrier + synchronize, nothing but the sweep inside.
    # Region B (roofline): K more steps with one CUDA event at every launch boundary on the launching stream
    # (a launch's duration = the gap between its two events, launch latency included); the events cost a few
    # percent of the step, which is why they are kept out of region A.
    for _ in range(max(args.warmup, 3)):
        sweep()
    barrier()
    launches0 = L.ec_kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for k in range(args.steps):
        sweep()
    stop.record()
    barrier()
    launches = L.ec_kernel_launches() - launches0
    total_ms = max_over_ranks(start.elapsed_time(stop))
    ms_per_step = total_ms / args.steps
    value = len(pairs) * cells * world / (ms_per_step * 1e-3) / 1e9

    # the same region with launch overlap (programmatic dependent launch) off: what the drain/ramp between the 41 launches costs
    overlap_was = L.ec_set_launch_overlap(0)
    sweep()
    barrier()
    start_o, stop_o = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start_o.record()
    for k in range(args.steps):
        sweep()
    stop_o.record()
    barrier()
    L.ec_set_launch_overlap(overlap_was)
    ms_per_step_serial = max_over_ranks(start_o.elapsed_time(stop_o)) / args.steps

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(pairs) + 1)] for _ in range(args.steps)]
    start_b, stop_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start_b.record()
    for k in range(args.steps):
        sweep(ev[k])
    stop_b.record()
    barrier()
    clocks = sampler.stop()
    ms_per_step_b = start_b.elapsed_time(stop_b) / args.steps

    per_pair, kern_ms_sum = [], 0.0
    for i, (s, d) in enumerate(pairs):
        ms = float(np.mean([ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(args.steps)]))
        kern_ms_sum += ms
        per_pair.append({"src": CT_NAMES[s], "dst": CT_NAMES[d], "bytes_per_cell": CT_SIZE[s] + CT_SIZE[d], "ms": round(ms, 4),
                         "GBps": round(pair_bytes[(s, d)] / (ms * 1e-3) / 1e9, 1), "Gcells_s": round(cells / (ms * 1e-3) / 1e9, 2)})
    peak, peak_src = measured_peak()
    # The step is nothing but this kernel family (41 launches, back to back on one stream), so the family's average
    # launch duration over the timed region is region A's time / launches — gaps between launches included. Region B's
    # per-launch event pairs serialise the launches and add the cost of the events: reported beside it.
    achieved = step_bytes / (ms_per_step * 1e-3) / 1e9
    achieved_ev = step_bytes / (kern_ms_sum * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per step from the committed ncu capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("convert_sweep_dram_bytes_per_step")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": (traffic // len(pairs)) if traffic else None, "traffic_per_step": traffic,
                "traffic_source": "profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the step's 41 launches under ncu, per launch = / 41",
                "peak_source": peak_src, "frac_of_nominal_8TBs": round(achieved / NOMINAL_GBS, 4),
                "kernel": "map1_kernel<CastF<S,D>> family: 31 casts + 10 clones per step (100% of the step's kernels)",
                "algorithmic_bytes_per_step": step_bytes, "launches_per_step": len(pairs),
                "avg_launch_ms": round(ms_per_step / len(pairs), 5), "algorithmic_bytes_per_launch": step_bytes // len(pairs),
                "timing": "CUDA events around the K timed steps on the launching stream (region A): step time / 41 launches, inter-launch gaps included",
                "per_launch_events": {"achieved": round(achieved_ev, 1), "frac": round(achieved_ev / peak, 4), "kernel_ms_per_step": round(kern_ms_sum, 4),
                                      "kernel_share_of_step": round(kern_ms_sum / ms_per_step_b, 4), "ms_per_step_with_events": round(ms_per_step_b, 4),
                                      "note": "one event pair per launch: serialises the launches (no overlap across an event) and adds the events' cost"},
                "launch_overlap": {"on_ms_per_step": round(ms_per_step, 4), "off_ms_per_step": round(ms_per_step_serial, 4),
                                   "enabled": bool(overlap_was), "what": "programmatic dependent launch: a grid's CTAs are scheduled while the previous grid drains"}}

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        def pinned(nbytes):
            p = C.c_void_p()
            ec._lib.check(L.ec_host_alloc(nbytes, C.byref(p)))
            return p, np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value))
        hsrc, keep = [], []
        for ct in range(10):
            p, raw = pinned(cells * CT_SIZE[ct])
            keep.append(p)
            a = raw.view(CellType(ct).dtype)
            srcs[ct].to_vec(out=a)  # same cells as the device-resident run
            hsrc.append(a)
        pout, rawout = pinned(cells * 8)
        h2d = sum(CT_SIZE[s] for s in range(10)) * cells
        d2h = sum(CT_SIZE[d] for _, d in pairs) * cells

        def e2e_step():
            chk = 0
            nxt = CellBuffer.from_vec(hsrc[0], wait=False)              # H2D on the upload stream
            for s in range(10):
                buf = nxt
                if s + 1 < 10:
                    nxt = CellBuffer.from_vec(hsrc[s + 1], wait=False)  # next source uploads while this one's results download
                for d in range(10):
                    if CellType(s).can_fit_into(CellType(d)):
                        out = buf.convert(CellType(d)).to_vec(out=rawout[: cells * CT_SIZE[d]].view(CellType(d).dtype))  # D2H
                        chk ^= int(out.view(np.uint8)[-1])
                    else:
                        try:
                            buf.convert(CellType(d))
                        except ec.NarrowingError:
                            pass
            return chk

        e2e_steps = max(2, min(args.steps, 4))
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        # what the host links can do at all, with every rank copying at once: bare pinned cudaMemcpyAsync, 256 MiB blocks,
        # D2H alone and D2H + H2D together (the e2e step is D2H-bound with the uploads hidden behind it)
        probe = pcie_probe(torch, barrier, max_over_ranks)
        d2h_rate = d2h / (e2e_ms * 1e-3) / 1e9
        e2e = {"value": len(pairs) * cells * world / (e2e_ms * 1e-3) / 1e9, "unit": "Gcells/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps,
               "pcie_GBps_each_way": [round(h2d / (e2e_ms * 1e-3) / 1e9, 2), round(d2h_rate, 2)],
               "roofline": {"bound": "pcie_d2h", "achieved": round(d2h_rate, 2), "peak": probe["d2h_GBps_per_rank_all_ranks_busy"], "unit": "GB/s per rank",
                            "frac": round(d2h_rate / probe["d2h_GBps_per_rank_all_ranks_busy"], 3), "probe": probe,
                            "note": "peak = this rank's bare D2H rate while all ranks copy at once (the host side is shared); "
                                    "the step moves d2h_bytes_per_step per rank"},
               "api": "CellBuffer.from_vec(pinned host, wait=False: next source prefetched) -> convert(ct) -> to_vec(pinned host), per rank"}
        if world == 1:
            # The crate's own callers hold Vec<T>: PAGEABLE memory. Same step from / into numpy arrays, once through the
            # library's staged path (8 MiB chunks via pinned staging, a pool of host threads) and once leaving the pageable
            # copies to the CUDA driver (ec_set_host_copy_threads(0)).
            pinned_src, pinned_out = hsrc, rawout
            hsrc = [np.array(a) for a in pinned_src]
            rawout = np.zeros(cells * 8, dtype=np.uint8)
            threads = L.ec_set_host_copy_threads(0)
            pageable = {}
            for label, t in (("driver", 0), ("staged", threads)):
                L.ec_set_host_copy_threads(t)
                e2e_step()
                t0 = time.perf_counter()
                chk = e2e_step()
                ms = (time.perf_counter() - t0) * 1e3
                pageable[label] = {"value": round(len(pairs) * cells / (ms * 1e-3) / 1e9, 3), "ms_per_step": round(ms, 1), "host_copy_threads": t,
                                   "d2h_GBps": round(d2h / (ms * 1e-3) / 1e9, 2), "checksum": chk}
            assert pageable["driver"]["checksum"] == pageable["staged"]["checksum"]
            pageable["what"] = ("same step with numpy (pageable) sources and destination, what a Vec<T> caller sees: 'staged' = the library's "
                                "chunked copy through pinned staging with a pool of host threads, 'driver' = cudaMemcpy on pageable memory")
            e2e["pageable"] = pageable
            hsrc, rawout = pinned_src, pinned_out
        for p in keep + [pout]:
            L.ec_host_free(p)
    del srcs

    # ---- strong scaling on the 32768^2 rasters (configs 4 and 5) with the cross-GPU finish, parity asserted ----------
    strong = None
    if not args.no_configs:
        strong = strong_scaling(ec, L, torch, dist, rank, world, local, barrier, max_over_ranks, peak, parity)

    # ---- the other BASELINE configs, once each (not the headline) -----------------------------------
    configs = {}
    if not args.no_configs:
        configs = other_configs(ec, L, torch, dist, rank, world, barrier, max_over_ranks, peak)
    parity_ok = all(parity.values())
    if world > 1:  # every rank must have seen parity
        t = torch.tensor([int(parity_ok)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        parity_ok = bool(int(t.item()))

    # ---- CPU baseline beside it (rank 0, N == 1): oracle port on a bounded sample ----------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        orc.build()
        n = 1 << 21
        hs = [synth.host(CellType(ct), n, 0xEC10 + ct) for ct in range(10)]
        t = 0.0
        for s, d in pairs:
            orc.convert(hs[s], d)
            t += orc.last_op_seconds()
        cpu = {"value": len(pairs) * n / t / 1e9, "unit": "Gcells/s", "cores": 1, "kind": "port",
               "sample": f"first {n} cells of each source buffer, all {len(pairs)} legal pairs, faithful per-cell tagged path "
                         f"(reference is single-threaded), {t:.1f} s of CPU work", "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gcells/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8..f64 casts (integer/f32/f64)", "data": "synthetic",
            "config": workload_config(cells),
            "device": {"name": info.name.decode(), "sm_count": info.sm_count, "l2_MB": info.l2_bytes >> 20},
            "parity_ok": parity_ok, "parity": parity, "strong_scaling": strong,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "GBps_algorithmic_step": round(step_bytes * world / (ms_per_step * 1e-3) / 1e9, 1),
            "per_pair": per_pair, "configs": configs,
        }
        out.emit(line)
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit("bench.py: a result differs from the oracle / between ranks: " + json.dumps(parity))


def pcie_probe(torch, barrier, max_over_ranks, block=256 << 20, reps=6):
    """Bare pinned-memory copies on every rank at once (torch copy_ = cudaMemcpyAsync): GB/s per rank, worst rank."""
    dev = torch.empty(block, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(block, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(block, dtype=torch.uint8).pin_memory()
    h_in = torch.empty(block, dtype=torch.uint8).pin_memory()
    side = torch.cuda.Stream()

    def timed(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        return reps * block / dt / 1e9

    def d2h():
        h_out.copy_(dev, non_blocking=True)

    def h2d():
        dev.copy_(h_in, non_blocking=True)

    def duplex():
        h_out.copy_(dev, non_blocking=True)
        with torch.cuda.stream(side):
            dev2.copy_(h_in, non_blocking=True)
    res = {"d2h_GBps_per_rank_all_ranks_busy": round(timed(d2h), 2), "h2d_GBps_per_rank_all_ranks_busy": round(timed(h2d), 2),
           "duplex_GBps_each_way_per_rank": round(timed(duplex), 2), "block_MiB": block >> 20,
           "how": "pinned host <-> HBM, cudaMemcpyAsync, every rank at the same time, slowest rank"}
    del dev, dev2, h_out, h_in
    return res


def strong_scaling(ec, L, torch, dist, rank, world, local, barrier, max_over_ranks, peak, parity):
    """BASELINE configs 4 and 5 as the north star states them: ONE 32768^2 raster in row strips over the N GPUs, the
    reduction finishing across the GPUs; 8 NDVI tiles dealt over the N GPUs. Fixed total work (strong scaling): at N > 1
    rank 0 also times the same work alone on its GPU in the same run (`n1_ms_same_run`), so the speed-up does not depend
    on another run or another box. Every result is asserted: against constants the oracle computed over the whole
    raster (tests/golden/bench_expected.json), against the oracle on sampled windows, and between the ranks."""
    from erased_cells_b200 import CellBuffer, CellType, sharding, synth
    from oracle import oracle as orc
    res = {"n_gpus": world, "scaling": "strong"}
    expected = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_expected.json")))

    def timed(fn, iters, warm=3):
        """device time per call (CUDA events on the launching stream, max over ranks) and host-visible wall clock per call"""
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.current_stream().synchronize()
        wall = (time.perf_counter() - t0) * 1e3 / iters
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / iters, max_over_ranks(wall)

    def solo(fn, iters, warm=3):
        """rank 0 alone (the other ranks wait at the barriers): the N = 1 time of the same work in the same run"""
        ms = 0.0
        barrier()
        if rank == 0:
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
        barrier()
        return max_over_ranks(ms)

    def entry(ms, wall_ms, nbytes, ncells, n1_ms=None, **kw):
        g = nbytes / (ms * 1e-3) / 1e9
        e = dict(ms=round(ms, 4), host_visible_ms=round(wall_ms, 4), GBps=round(g, 1), Gcells_s=round(ncells / (ms * 1e-3) / 1e9, 2),
                 frac_of_measured_peak_x_gpus=round(g / (peak * world), 3), **kw)
        if n1_ms is not None:
            e.update(n1_ms_same_run=round(n1_ms, 4), speedup_vs_n1_same_run=round(n1_ms / ms, 3), linear_speedup=world)
        return e

    # ---- config 4: f32 32768^2 min_max (+ statistics), row strips, cross-GPU finish -----------------------------------
    side = 32768
    n4 = side * side
    exp4 = expected["c4_f32_32768"]
    off, ln = sharding.row_strip(side, side, world, rank)
    strip = synth.device(CellType.Float32, ln, 0xEC40, index_offset=off, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
    keys = torch.empty(2, dtype=torch.int64, device="cuda")
    hkeys = torch.empty(2, dtype=torch.int64).pin_memory()

    def c4_nccl():
        ec._lib.check(L.ec_buf_min_max_keys(strip._h, None, C.c_void_p(keys.data_ptr())))
        if world > 1:
            dist.all_reduce(keys, op=dist.ReduceOp.MIN)
        hkeys.copy_(keys, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        mn, mx = ec._lib.Value(), ec._lib.Value()
        ec._lib.check(L.ec_min_max_from_keys(int(CellType.Float32), hkeys.numpy().ctypes.data_as(C.POINTER(C.c_int64)), C.byref(mn), C.byref(mx)))
        return hex(mn.bits), hex(mx.bits)

    comm = sharding.Comm.create() if world > 1 else None

    # the C ABI call itself, arguments bound once (what a compiled caller pays; the Python mirror adds ~5 us of object churn per call)
    vmn, vmx = ec._lib.Value(), ec._lib.Value()
    if comm is not None:
        c4_call, c4_args = L.ec_buf_min_max_sharded, (comm._h, strip._h, None, C.byref(vmn), C.byref(vmx))
    else:
        c4_call, c4_args = L.ec_buf_min_max, (strip._h, None, C.byref(vmn), C.byref(vmx))

    def c4_fused():
        ec._lib.check(c4_call(*c4_args))
        return hex(vmn.bits), hex(vmx.bits)

    def c4_fused_timed():
        c4_call(*c4_args)

    n1_ms = None
    if world > 1:
        whole = [None]
        if rank == 0:
            whole[0] = synth.device(CellType.Float32, n4, 0xEC40, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
        n1_ms = solo(lambda: whole[0].min_max(), 10)
        whole[0] = None
    want = (exp4["min_bits"], exp4["max_bits"])
    got_nccl, got_fused = c4_nccl(), c4_fused()
    parity["c4_min_max_equals_oracle_constant"] = got_nccl == want and got_fused == want
    # sampled windows of this rank's strip against the oracle (the strip really holds the raster the constant was computed from)
    okw = True
    for w0 in (0, (ln // 2) & ~127, ln - (1 << 16)):
        mn, mx = strip.view(w0, 1 << 16).min_max()
        omn, omx = orc.tight_min_max(synth.host(CellType.Float32, 1 << 16, 0xEC40, index_offset=off + w0, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4))
        okw &= (mn.bits, mx.bits) == (omn.bits, omx.bits)
    parity["c4_sampled_windows_vs_oracle"] = okw
    ms, wall = timed(c4_fused_timed, 40, 10)
    parity["c4_min_max_after_timing"] = c4_fused() == want
    res["c4_f32_32768_min_max"] = entry(ms, wall, 4.0 * n4, n4, n1_ms, shards=world, result_bits=list(got_fused), expected_bits=list(want),
                                        finish=("ec_buf_min_max_sharded: reduction + NVLink peer exchange + fold in ONE kernel per GPU, result polled from mapped pinned memory"
                                                if comm is not None and comm.peer_exchange else "ec_buf_min_max (one GPU)" if comm is None else "kernel + ncclAllReduce (ec_comm)"))
    # the same 40 calls issued by a C loop (tools/call_loop.c): a compiled caller, no interpreter between the calls
    loop_so = os.path.join(ROOT, "tools", "bin", "libec_call_loop.so")
    if os.path.exists(loop_so):
        LL = C.CDLL(loop_so)
        LL.ec_loop_min_max.restype = C.c_int
        LL.ec_loop_min_max.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        comm_h = comm._h if comm is not None else None

        def c_loop(iters):
            ec._lib.check(LL.ec_loop_min_max(comm_h, strip._h, iters, C.byref(vmn), C.byref(vmx)))

        c_loop(10)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        c_loop(40)
        b.record()
        torch.cuda.current_stream().synchronize()
        wall_c = max_over_ranks((time.perf_counter() - t0) * 1e3 / 40)
        barrier()
        ms_c = max_over_ranks(a.elapsed_time(b)) / 40
        parity["c4_min_max_after_c_loop"] = (hex(vmn.bits), hex(vmx.bits)) == want
        res["c4_f32_32768_min_max_c_loop"] = entry(ms_c, wall_c, 4.0 * n4, n4, n1_ms, shards=world, result_bits=[hex(vmn.bits), hex(vmx.bits)],
                                                   finish="same entry point as c4_f32_32768_min_max, the 40 timed calls issued back to back from C (tools/call_loop.c)")
    ms, wall = timed(c4_nccl, 20)
    res["c4_f32_32768_min_max_nccl"] = entry(ms, wall, 4.0 * n4, n4, n1_ms, shards=world, result_bits=list(got_nccl),
                                             finish="shard kernel + torch.distributed all-reduce(MIN, 2 x int64) + D2H of the result")
    st = comm.statistics(strip) if comm is not None else strip.statistics()
    parity["c4_statistics_count_min_max"] = (st.count, hex(st.min.bits), hex(st.max.bits)) == (exp4["count"], exp4["min_bits"], exp4["max_bits"])
    n1s = None
    if world > 1:
        whole = [None]
        if rank == 0:
            whole[0] = synth.device(CellType.Float32, n4, 0xEC40, kind=synth.REAL_RANGE, lo=-1e4, hi=1e4)
        n1s = solo(lambda: whole[0].statistics(), 5)
        if rank == 0:
            s1 = whole[0].statistics()
            parity["c4_statistics_bits_equal_one_gpu"] = np.array_equal(np.array([s1.mean, s1.stddev]).view(np.uint64), np.array([st.mean, st.stddev]).view(np.uint64))
        whole[0] = None
    ms, wall = timed((lambda: comm.statistics(strip)) if comm is not None else (lambda: strip.statistics()), 5)
    res["c4_f32_32768_statistics"] = entry(ms, wall, 4.0 * n4, n4, n1s, shards=world, result=[st.count, st.mean, st.stddev],
                                           note="extension (the reference has no statistics): parity unpinned, bit-identical for every N by construction; "
                                                "min_max pass + one pass of exact integer sums over the cells quantised to 2^-26 of their range; bytes counted once")
    if comm is None:  # when the raster's min_max is already known (ec_buf_min_max ran before, the buffer remembers it): ONE pass
        L.ec_set_min_max_cache(1)
        strip.min_max()
        s1 = strip.statistics()
        parity["c4_statistics_one_pass_equals_two_pass"] = np.array_equal(np.array([s1.mean, s1.stddev]).view(np.uint64), np.array([st.mean, st.stddev]).view(np.uint64))
        ms, wall = timed(lambda: strip.statistics(), 5)
        L.ec_set_min_max_cache(0)
        res["c4_f32_32768_statistics_min_max_known"] = entry(ms, wall, 4.0 * n4, n4, None, shards=world, note="one pass: the buffer remembers its min_max from an earlier call")
    del strip
    if comm is not None:
        comm.close()

    # ---- config 5: 8 NDVI tiles (u16 32768^2 each) over the N GPUs: fused (nir - red) / (nir + red) -> f64, per-tile min_max,
    #      then ONE all-reduce for the mosaic's min / max ------------------------------------------------------------------
    n5 = side * side
    tiles = [t for t in range(8) if t % world == rank]

    def make(t):
        return (synth.device(CellType.UInt16, n5, 0xEC50 + t, kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0),
                synth.device(CellType.UInt16, n5, 0xEC58 + t, kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0))
    bands = {t: make(t) for t in tiles}
    mosaic = torch.empty(2, dtype=torch.int64, device="cuda")

    def c5(which):
        k0 = k1 = None
        for t in which:
            nir, red = bands[t]
            ndvi = nir.normalized_difference(red)
            mn, mx = ndvi.min_max()
            k = sharding.keys_of(mn, mx)
            k0 = k[0] if k0 is None else min(k0, k[0])
            k1 = k[1] if k1 is None else min(k1, k[1])
        return k0, k1

    def c5_step():
        k0, k1 = c5(tiles)
        if world > 1:
            mosaic.copy_(torch.tensor([k0, k1], dtype=torch.int64))
            dist.all_reduce(mosaic, op=dist.ReduceOp.MIN)
            k0, k1 = [int(x) for x in mosaic.tolist()]
        return k0, k1

    # parity: sampled windows of every local tile's NDVI against the oracle
    ok5 = True
    for t in tiles:
        nir, red = bands[t]
        for w0 in (0, n5 - (1 << 16)):
            got = nir.view(w0, 1 << 16).normalized_difference(red.view(w0, 1 << 16)).to_vec()
            hn = synth.host(CellType.UInt16, 1 << 16, 0xEC50 + t, index_offset=w0, kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0)
            hr = synth.host(CellType.UInt16, 1 << 16, 0xEC58 + t, index_offset=w0, kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0)
            want5 = orc.tight_binary(orc.DIV, orc.tight_binary(orc.SUB, hn, hr), orc.tight_binary(orc.ADD, hn, hr))
            ok5 &= bool(np.array_equal(got.view(np.uint64), want5.view(np.uint64)))
    parity["c5_ndvi_sampled_windows_vs_oracle"] = ok5
    keys5 = c5_step()
    ms, wall = timed(c5_step, 3, 1)
    n1_5 = None
    if world > 1:
        if rank == 0:
            for t in range(8):
                if t not in bands:
                    bands[t] = make(t)
        n1_5 = solo(lambda: c5(range(8)), 2, 1)
        if rank == 0:
            parity["c5_mosaic_min_max_equals_one_gpu"] = tuple(int(x) for x in c5(range(8))) == tuple(int(x) for x in keys5)
    mnv, mxv = sharding.values_of(CellType.Float64, [keys5[0], keys5[1]])
    res["c5_ndvi_8_tiles_u16_32768"] = entry(ms, wall, 8 * (12.0 + 8.0) * n5, 8 * n5, n1_5, tiles_per_gpu=len(tiles), result_bits=[hex(mnv.bits), hex(mxv.bits)],
                                             note="per tile: fused NDVI (12 B/cell) + min_max of the f64 result (8 B/cell); one 16-byte all-reduce per step")
    bands.clear()
    ec._lib.check(L.ec_trim())

    # ---- the same raster driven by ONE process over all N GPUs through the plain C ABI (tools/shard_latency.c) ---------
    exe = os.path.join(ROOT, "tools", "bin", "shard_latency")
    if world > 1 and os.path.exists(exe):
        host_group = dist.new_group(backend="gloo")
        barrier()
        torch.cuda.empty_cache()
        single = None
        if rank == 0:
            try:
                env = dict(os.environ)
                for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "EC_DEVICE", "EC_DEVICES"):
                    env.pop(k, None)
                r = subprocess.run([exe, str(side), ",".join(str(i) for i in range(world)), "30"], capture_output=True, text=True, timeout=300, env=env)
                single = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1][len("RESULT "):])
                parity["c4_single_process_sharded_equals_oracle_constant"] = all(
                    [hex(int(b, 16)) for b in single[k]] == list(want) for k in single if k.startswith("bits_"))
            except Exception as e:  # the probe is extra evidence, not the bench
                single = {"error": repr(e)[:300]}
        res["c4_single_process_all_gpus"] = single
        # the other ranks must not touch their GPUs meanwhile: they wait on the host (gloo); an NCCL barrier would spin a kernel there
        dist.barrier(group=host_group)
        barrier()
    return res


def other_configs(ec, L, torch, dist, rank, world, barrier, max_over_ranks, peak):
    """BASELINE configs 1, 3, 4, 5: a few timed iterations each, CUDA events, buffers >> L2 or rotated."""
    from erased_cells_b200 import CellBuffer, CellType, MaskedCellBuffer, NoData, synth
    res = {}

    def timed(fn, iters=5, warm=2):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / iters

    def entry(ms, nbytes, ncells, **kw):
        g = nbytes / (ms * 1e-3) / 1e9
        return dict(ms=round(ms, 4), GBps=round(g, 1), Gcells_s=round(ncells / (ms * 1e-3) / 1e9, 2), frac_of_measured=round(g / peak, 3),
                    frac_of_8TBs=round(g / NOMINAL_GBS, 3), **kw)

    # config 1: u8 4096^2 / u16 4096^2 * 0.5 -> f64; 8 rotating buffer sets so inputs are not L2-resident
    n1 = 4096 * 4096
    sets = [(synth.device(CellType.UInt8, n1, 0xEC01 + 16 * i, kind=synth.INT_RANGE, lo=0, hi=255),
             synth.device(CellType.UInt16, n1, 0xEC02 + 16 * i, kind=synth.INT_RANGE, lo=0, hi=65535)) for i in range(8)]
    it = [0]

    def c1_unfused():
        a, b = sets[it[0] % 8]; it[0] += 1
        return a / b * 0.5

    def c1_div():
        a, b = sets[it[0] % 8]; it[0] += 1
        return a / b

    def c1_fused():
        a, b = sets[it[0] % 8]; it[0] += 1
        return a.binary_scalar(ec.DIV, b, ec.MUL, 0.5)
    def c1_lazy():  # the README expression itself, deferred: `a / b * 0.5` runs as one fused kernel on first access
        a, b = sets[it[0] % 8]; it[0] += 1
        with ec.lazy():
            r = a / b * 0.5
            r.device_ptr()
        return r
    res["c1_readme_4096"] = {"div_u8_u16": entry(timed(c1_div, 16), 11 * n1, n1), "div_then_mul_unfused": entry(timed(c1_unfused, 16), 27 * n1, n1),
                             "div_mul_fused": entry(timed(c1_fused, 16), 11 * n1, n1), "div_mul_lazy_operators": entry(timed(c1_lazy, 16), 11 * n1, n1),
                             "scaling": "weak (same buffers on every rank)"}
    del sets

    # config 3: masked i16 16384^2 with NoData: (a - b) * s, min_max, counts
    n3 = 16384 * 16384
    a = synth.device(CellType.Int16, n3, 0xEC31, index_offset=rank * n3, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=50, sentinel=-32768)
    b = synth.device(CellType.Int16, n3, 0xEC32, index_offset=rank * n3, kind=synth.INT_RANGE, lo=-32768, hi=32767, period=50, sentinel=-32768)
    nd = NoData.default(CellType.Int16)
    c3 = {"from_nodata_i16": entry(timed(lambda: MaskedCellBuffer.from_buffer_with_nodata(a, nd)), 2.125 * n3, n3)}
    ma, mb = MaskedCellBuffer.from_buffer_with_nodata(a, nd), MaskedCellBuffer.from_buffer_with_nodata(b, nd)
    c3["masked_sub_i16_i16"] = entry(timed(lambda: ma - mb), 12.375 * n3, n3)
    r = ma - mb
    c3["masked_mul_scalar_f64"] = entry(timed(lambda: r * 0.0001), 16.0 * n3, n3, note="the result shares the operand's mask words (refcounted): 16 B/cell, no mask traffic")
    def c3_lazy():  # `(&ma - &mb) * 0.0001` through the operators, deferred: one fused data pass + the mask AND
        with ec.lazy():
            x = (ma - mb) * 0.0001
            x.buffer().device_ptr()
        return x
    c3["masked_sub_then_scale_lazy_1_pass"] = entry(timed(c3_lazy), 12.0 * n3, n3, note="data kernel 12 B/cell; mask AND + clone ride along (5/8 B/cell)")
    rs = r * 0.0001
    c3["masked_min_max_f64"] = entry(timed(lambda: rs.min_max()), 8.125 * n3, n3, note="includes the 16-byte D2H + stream sync of the result")
    c3["counts"] = entry(timed(lambda: rs.counts()), 0.125 * n3, n3, note="the kernel that produced the mask counted its set bits: no launch, the cached value")
    fresh = ~(~rs.mask())
    c3["counts_popcount_pass"] = entry(timed(lambda: (rs.mask() & fresh).counts()), 0.375 * n3, n3, note="mask AND (3/8 B/cell) with the count produced by the same kernel, polled from pinned memory")
    # statistics extension (count/min/max/mean/stddev; the reference has none): integer cells of <= 32 bits are ONE pass
    c3["statistics_masked_i16_one_pass"] = entry(timed(lambda: ma.statistics()), 2.125 * n3, n3, note="extension: min, max, count, sum x, sum x^2 in one read; host finish + sync included")
    c3["statistics_masked_f64_two_passes"] = entry(timed(lambda: rs.statistics()), 8.125 * n3, n3, note="extension: min_max pass + FP64 window moments pass; bytes counted once")
    res["c3_masked_i16_16384"] = c3
    del a, b, ma, mb, r, rs

    # config 5: NDVI (nir - red) / (nir + red), u16 32768^2 -> f64, one tile per GPU (weak)
    n5 = 32768 * 32768
    nir = synth.device(CellType.UInt16, n5, 0xEC50 + rank, kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0)
    red = synth.device(CellType.UInt16, n5, 0xEC58 + rank, kind=synth.INT_RANGE, lo=5000, hi=40000, period=1000, sentinel=0)
    res["c5_ndvi_u16_32768_per_gpu_tile"] = {
        "unfused_3_ops": entry(timed(lambda: (nir - red) / (nir + red), 3, 1), 48.0 * n5, n5),
        "fused_1_pass": entry(timed(lambda: nir.normalized_difference(red), 3, 1), 12.0 * n5, n5),
        "scaling": "weak (one tile per GPU)"}
    def c5_lazy():  # `(&nir - &red) / (&nir + &red)` through the operators, deferred -> one fused pass
        with ec.lazy():
            r = (nir - red) / (nir + red)
            r.device_ptr()
        return r
    res["c5_ndvi_u16_32768_per_gpu_tile"]["lazy_operators_1_pass"] = entry(timed(c5_lazy, 3, 1), 12.0 * n5, n5)
    nd5 = nir.normalized_difference(red)
    res["c5_ndvi_u16_32768_per_gpu_tile"]["min_max_f64"] = entry(timed(lambda: nd5.min_max(), 3, 1), 8.0 * n5, n5)

    # beyond the BASELINE configs: a longer band-math chain through the operators, EVI = 2.5*(nir-red)/(nir+6*red-7.5*blue+1),
    # 8 ops over three u16 bands — op by op vs. one kernel specialised at run time in lazy mode
    del nir, red, nd5
    ne = 16384 * 16384
    b_nir, b_red, b_blue = [synth.device(CellType.UInt16, ne, 0xEC60 + i, kind=synth.INT_RANGE, lo=100, hi=40000) for i in range(3)]

    def evi():
        return ((b_nir - b_red) * 2.5) / (((b_nir + b_red * 6.0) - b_blue * 7.5) + 1.0)

    def evi_jit():  # the same operators, chain compiled at run time into one streaming kernel (NVRTC build outside the timed calls)
        with ec.lazy(jit=True):
            r = evi()
            r.device_ptr()
        return r
    unfused_bytes = (2 + 2 + 8) + 16 + (2 + 8) + (2 + 8 + 8) + (2 + 8) + 24 + 16 + 24  # per cell, op by op
    res["extra_evi_u16_16384"] = {"unfused_8_ops": entry(timed(evi, 3, 1), float(unfused_bytes) * ne, ne),
                                  "lazy_specialised_kernel_1_pass": entry(timed(evi_jit, 3, 1), 14.0 * ne, ne, kernel=L.ec_last_kernel().decode()),
                                  "note": "not a BASELINE config; opt-in ec_set_lazy(3) = kernel specialised at run time with NVRTC (falls back to op by op without libnvrtc); "
                                          "GB/s of the 1-pass line = 14 B/cell (3 x u16 in, f64 out) / time"}
    return res


class OneLineStdout:
    """The contract is ONE JSON line on stdout. Libraries underneath (NCCL's version banner, for one) write to fd 1
    on their own, so for the duration of the run fd 1 is pointed at stderr and the line is written to the saved fd."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, obj):
        os.write(self.saved, (json.dumps(obj) + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=SIDE * SIDE)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    with OneLineStdout() as out:
        if args.impl == "reference":
            run_reference(args, out)
        else:
            run_ours(args, out)


if __name__ == "__main__":
    main()
