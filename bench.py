#!/usr/bin/env python
"""bench.py -- codewords/s of the LLR-domain SCL polar decoder (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c4|c1|c2|c3|c5] [--mode strict|fp32]

A "step" is one pass of the hot path (decode_scl_llr) over one batch of synthetic BPSK/AWGN
codewords. Default workload = BASELINE.json configs[3], the one the metric is quoted on:
N=2048, K=1024, 16 parity ("CRC") bits, list 32, Eb/N0 swept 1.0-2.5 dB (the sweep points are
interleaved over the batch), 65536 codewords per GPU; the batch shards over ranks with no
data-path collective (weak scaling) and only the block-error counters are all-reduced.

One JSON line on rank 0. `value` = whole-job codewords/s with LLRs resident in HBM (CUDA events on
the launch stream, max over ranks); `e2e` = the same through the C ABI's host entry point
(pinned host LLRs -> H2D -> decode -> D2H of the packed bits, every step); `roofline` = the
decode kernel against the measured HBM copy bandwidth with SURVEY.md section 8(d)'s algorithmic bytes
(4N + ceil(K/8) per codeword); `cpu_baseline` = the unmodified reference (oracle/_ref) or its port
timed on this box's host cores on a bounded sample of the same workload.

`--impl reference` times that CPU implementation instead (rank 0 only, all host threads, each
step a bounded sample of the same workload).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n, K, crc, L, sweep of Eb/N0 in dB, BASELINE.json index)
    "c1": (9, 256, 0, 1, (2.0,), 0),
    "c2": (11, 1024, 0, 1, (2.0,), 1),
    "c3": (11, 1024, 16, 4, (2.0,), 2),
    "c4": (11, 1024, 16, 32, (1.0, 1.25, 1.5, 1.75, 2.0, 2.25, 2.5), 3),
    "c5": (9, 256, 0, 32, (2.0,), 4),
}
SEED = 0x5EED0000


def workload_name(cfg, batch):
    n, K, crc, L, sweep, idx = CONFIGS[cfg]
    return "configs[%d]: N=%d K=%d crc=%d L=%d SCL-LLR, Eb/N0 %s dB, batch=%d codewords/GPU" % (
        idx, 1 << n, K, crc, L, "%.2f-%.2f sweep" % (sweep[0], sweep[-1]) if len(sweep) > 1 else "%.2f" % sweep[0], batch)


def recorded_traffic(cfg, batch):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/ncu_traffic.json,
    written by tools/ncu_traffic.py), scaled from the captured batch to this one; None if absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[cfg]
        return d["dram_bytes_per_codeword"] * batch, d
    except Exception:
        return None, None


def wilson(err, n, z=1.96):
    """95 % Wilson interval of a binomial proportion"""
    if n == 0:
        return [None, None]
    p = err / n
    den = 1 + z * z / n
    c = (p + z * z / (2 * n)) / den
    h = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / den
    return [float(max(0.0, c - h)), float(min(1.0, c + h))]


def issue_roofline(src, cw_per_s_per_gpu, sm_count, clocks):
    try:
        ipc = float(src["warp_instructions_per_codeword"])
        mhz = float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz"))
        peak = sm_count * 4 * mhz * 1e6
        return {"bound": "issue", "achieved": cw_per_s_per_gpu * ipc, "peak": peak, "unit": "warp-instr/s",
                "frac": cw_per_s_per_gpu * ipc / peak, "warp_instr_per_codeword": ipc, "source": src.get("source")}
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe). Started before the warm-up so that
    the tool is already streaming when the timed region begins; samples are time-stamped and only those
    inside [mark_begin, mark_end] are used (all samples under load if the region was shorter than one tick)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                f = [x.strip() for x in r.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, reasons
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e30) + 0.06]
        window = "timed region"
        if not inside:
            inside, window = self.rows[-8:], "warm-up + timed region (timed region shorter than one sampling tick)"
        sm, mx, reasons = parse(inside)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def cpu_decoder(n, K, crc):
    """(kind, object with decode_batch(llr, L, nthreads)) -- the compiled reference when it was prebuilt."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    if oracle_lib.have_ref():
        return "reference", oracle_lib.Ref(n, K, 0.32, crc)
    return "port", oracle_lib.Port(n, K, 0.32, crc)


def cpu_decoder_o3(n, K, crc):
    """the same reference sources at -O3 -march=native, or None where the build is
    absent or this host lacks an ISA extension it was compiled for"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    if oracle_lib.have_ref_o3():
        return oracle_lib.Ref(n, K, 0.32, crc, so=oracle_lib.REF_O3_SO)
    return None


def cpu_table(dec, dec_o3, llr, L, threads, seconds):
    """SURVEY.md section 8(d)'s baseline table: 1 thread and all host threads, -O2 and -O3 -march=native, each on a
    bounded sample of the same batch (at least 200 / 500 / 2000 codewords per thread for lists 32 / 4 / 1 when the
    time budget allows)."""
    rows = []
    for name, d in (("-O2", dec), ("-O3 -march=native", dec_o3)):
        if d is None:
            rows.append({"build": name, "unavailable": "not prebuilt, or this host lacks an ISA extension it needs"})
            continue
        for th in sorted({1, threads}):
            v, S = time_cpu(d, llr, L, th, seconds)
            rows.append({"build": "g++ " + name, "threads": th, "value": v, "unit": "codewords/s", "sample_codewords": S})
    return rows


def time_cpu(dec, llr, L, threads, budget_s, want_out=False):
    """decode a bounded sample sized for ~budget_s seconds; returns (cw/s, sample size[, decoded bits])."""
    probe = min(len(llr), max(threads, 8))
    t0 = time.perf_counter(); dec.decode_batch(llr[:probe], L, threads); t = time.perf_counter() - t0
    rate = probe / max(t, 1e-6)
    S = int(min(len(llr), max(threads, rate * budget_s)))
    S = max(threads, (S // threads) * threads)
    t0 = time.perf_counter(); out = dec.decode_batch(llr[:S], L, threads); t = time.perf_counter() - t0
    return (S / t, S, out) if want_out else (S / t, S)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from polar_b200 import synth
    n, K, crc, L, sweep, _ = CONFIGS[args.config]
    kind, dec = cpu_decoder(n, K, crc)
    threads = os.cpu_count() or 1
    # sample sized so that warmup + steps finish in about two minutes
    per_step_s = max(1.0, 120.0 / (args.steps + args.warmup))
    info, llr = synth.make_shard(dec, SEED, 0, 4 * synth.BLOCK, sweep=sweep)
    probe = max(threads, 8)
    t0 = time.perf_counter(); dec.decode_batch(llr[:probe], L, threads); rate = probe / (time.perf_counter() - t0)
    S = int(min(len(llr), max(threads, rate * per_step_s)))
    S = max(threads, (S // threads) * threads)
    for _ in range(args.warmup):
        dec.decode_batch(llr[:S], L, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = dec.decode_batch(llr[:S], L, threads)
    dt = time.perf_counter() - t0
    value = S * args.steps / dt
    sample = "%d codewords/step of the same workload, %d host threads, %s" % (
        S, threads, "unmodified PolarC/PolarCode.cpp -O2 (oracle/_ref)" if kind == "reference" else "oracle port -O2")
    print(json.dumps({
        "impl": "reference", "metric": "codewords/sec", "value": value, "unit": "codewords/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.batch), "sample_per_step": S},
        "cpu_baseline": {"value": value, "unit": "codewords/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "codewords/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "bler": float((out != info[:S]).any(1).mean()),
    }), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from polar_b200 import PolarCode, bler, synth, pack_bits, unpack_bits
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- polar_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, K, crc, L, sweep, _ = CONFIGS[args.config]
    N, B = 1 << n, args.batch
    mode = args.mode                                  # arithmetic of the headline arm (default: strict)
    other = "fp32" if mode in ("strict", "f64") else "strict"
    code = PolarCode(n, K, 0.32, crc, device=local, mode=mode)
    KW = code.KW

    # this rank's shard of the global batch (weak scaling: B codewords per GPU), pinned on the host
    host_buf = None
    if args.host_alloc == "wc":
        # write-combined page-locked input buffer from the library (polar_b200_host_alloc): read by the GPU without
        # snooping the CPU caches
        from polar_b200 import HostBuffer
        host_buf = HostBuffer((B, N), np.float32, write_combined=True)
        h_llr = torch.from_numpy(host_buf.array)
    else:
        h_llr = torch.empty((B, N), dtype=torch.float32, pin_memory=True)
    h_out = torch.empty((B, KW), dtype=torch.int32, pin_memory=True)
    info, _ = synth.make_shard(code, SEED, rank * B, B, sweep=sweep, out_llr=h_llr.numpy())
    d_truth = torch.from_numpy(pack_bits(info).view(np.int32)).to(dev)
    d_llr = torch.empty((B, N), dtype=torch.float32, device=dev)
    d_llr.copy_(h_llr)
    d_out = torch.empty((B, KW), dtype=torch.int32, device=dev)
    d_out2 = torch.empty((B, KW), dtype=torch.int32, device=dev)
    d_berr = torch.zeros(B, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(m, steps, out):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            code.decode_device(d_llr, L, out=out, mode=m)
        e1.record(stream)
        return e0, e1

    # ---- device-resident arm (headline mode) ----
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        code.decode_device(d_llr, L, out=d_out)
    launches0 = code.kernel_launches
    barrier()
    sampler.mark_begin()
    e0, e1 = timed(mode, args.steps, d_out)
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = code.kernel_launches - launches0
    kernel_kind = code.info(6)             # first-pass kernel of the timed region
    flagged = code.last_flagged if mode == "strict" else 0
    code.count_errors(d_out, d_truth, d_berr, None)
    torch.cuda.synchronize(dev)

    # ---- the other arithmetic mode, for the record (same batch, same timing rules, fewer steps) ----
    o_steps = max(1, min(args.steps, 3))
    for _ in range(3):
        code.decode_device(d_llr, L, out=d_out2, mode=other)
    barrier()
    f0, f1 = timed(other, o_steps, d_out2)
    barrier()
    ms_other = f0.elapsed_time(f1) / o_steps
    flagged_other = code.last_flagged if other == "strict" else 0
    differs_between_modes = int((d_out != d_out2).any(1).sum().item())

    # ---- plain H2D copy of the step's input, all ranks at once (names the end-to-end limiter) ----
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream); d_llr.copy_(h_llr, non_blocking=True); c1.record(stream)
    barrier()
    h2d_ms = c0.elapsed_time(c1)

    # ---- end-to-end arm: pinned host LLRs in, packed bits out, through the C ABI host entry ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    code.decode_batch(h_llr, L, packed=True, out=h_out)      # warm (allocates the staging buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        code.decode_batch(h_llr, L, packed=True, out=h_out)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_ok = bool(np.array_equal(h_out.numpy(), d_out.cpu().numpy()))

    # per-Eb/N0 block-error counters of this rank's shard; global codeword index g uses sweep[g % len(sweep)]
    ns = len(sweep)
    be = d_berr.to(torch.int64)
    be2 = (d_out2 != d_truth).any(dim=1).to(torch.int64)
    counts = np.zeros((2, ns, 2), np.int64)          # [headline mode, other mode][Eb/N0 point][num_err, num_run]
    for e in range(ns):
        off = (e - rank * B) % ns
        counts[0, e, 0] = int(be[off::ns].sum().item())
        counts[1, e, 0] = int(be2[off::ns].sum().item())
        counts[:, e, 1] = int(be[off::ns].numel())
    # ---- the BLER workload end to end on the device (polar_b200_bler_sweep): synthesis, decode and comparison stay on
    # the GPU, per step only (seed, index range) go in and the counters come out ----
    sweep_steps = max(1, min(args.steps, args.e2e_steps))
    sw_counts = code.bler_sweep_device(sweep, [L], B, SEED + 1, first_index=rank * B)       # warm
    barrier()
    t0 = time.perf_counter()
    for i in range(sweep_steps):
        sw_counts = code.bler_sweep_device(sweep, [L], B, SEED + 1, first_index=(world * (i + 1) + rank) * B)
    sweep_s = time.perf_counter() - t0
    barrier()

    t = torch.tensor([ms, e2e_s, ms_other, h2d_ms, sweep_s], dtype=torch.float64, device=dev)
    fl = torch.tensor([flagged, flagged_other, differs_between_modes], dtype=torch.int64, device=dev)
    collective = "none (one GPU)"
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
        # the path's one real collective -- the (num_err, num_run) counters -- goes through the library's own NCCL
        # communicator (polar_b200_comm_*: ncclAllReduce issued from C++); torch.distributed only carries the 128-byte id
        def exchange(b):
            tt = torch.tensor(list(b), dtype=torch.uint8, device=dev)
            dist.broadcast(tt, 0)
            return bytes(tt.cpu().tolist())
        comm = bler.Comm(local, world, rank, exchange)
        counts = comm.all_reduce(counts)
        sw_counts = comm.all_reduce(sw_counts)
        comm.close()
        collective = "ncclAllReduce(int64 x %d) from C++ (polar_b200_comm_allreduce_i64), ranks = %d" % (counts.size, world)
    ms, e2e_s, ms_other, h2d_ms, sweep_s = (float(x) for x in t.tolist())
    flagged, flagged_other, differs_between_modes = (int(x) for x in fl.tolist())

    if rank == 0:
        ms_step = ms / args.steps
        value = world * B * args.steps / (ms * 1e-3)
        bytes_cw = 4 * N + (K + 7) // 8
        peak, peak_src = measured_peak()
        traffic, traffic_src = recorded_traffic(args.config, B)
        sm_count = code.info(1)
        # the dominant kernel (scl_fast_kernel) alone: in strict / f64 mode the step also holds the second pass, so its
        # launch duration is the fp32 arm's step (the same kernel, one launch per step, timed with CUDA events above)
        # (list size 1 in strict mode runs its own first-pass kernel, sc_ssc_kernel: its duration is the strict step itself,
        # the second pass behind it being a near-empty launch)
        ssc = kernel_kind == 500
        kernel_ms = ms_other if (mode in ("strict", "f64") and not ssc) else ms_step
        achieved = B * bytes_cw / (kernel_ms * 1e-3) / 1e9   # per GPU: one launch decodes this rank's B codewords
        strict_flagged = flagged if mode == "strict" else flagged_other
        out = {
            "metric": "codewords/sec", "value": value, "unit": "codewords/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if mode in ("fp32", "minsum") else ("f64" if mode == "f64" else "f32+f64"),
            "data": "synthetic",
            "config": {"workload": workload_name(args.config, B), "global_batch": world * B,
                       "arithmetic": {"strict": "fp32 kernels; codewords decided on a margin below tau decoded again in double "
                                                "(bit-exact against the double reference, see `parity`)",
                                      "fp32": "fp32 kernels alone", "f64": "everything in double",
                                      "minsum": "OPT-IN, NOT THE REFERENCE'S ARITHMETIC: min-sum check nodes + hardware-friendly "
                                                "metric update (SURVEY.md section 8(f)4); compare `modes` and the per-point BLER "
                                                "of both modes"}[mode],
                       "l2": "input %d MiB per GPU per step > 126 MB L2, no flush needed" % (B * N * 4 >> 20)
                             if B * N * 4 > 126e6 else "input %d MiB per GPU per step fits L2 (small plumbing config)" % (B * N * 4 >> 20),
                       "sharding": "contiguous codeword blocks per rank, no data-path collective; counters all-reduced"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_codeword": bytes_cw,
                         "algorithmic_bytes_per_launch": B * bytes_cw,
                         "traffic_source": (traffic_src or {}).get("source"),
                         "traffic_note": "dram bytes/codeword of the committed ncu --set full capture (batch %s) x this batch"
                                         % (traffic_src or {}).get("captured_batch", "16384"),
                         "kernel": "sc_ssc_kernel" if ssc else ("scl_fast_kernel" if kernel_kind > 0 else "scl_decode_kernel"),
                         "kernel_kind": kernel_kind, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms_step},
            "modes": {mode: value, other: world * B / (ms_other * 1e-3), "unit": "codewords/s (device-resident)",
                      "strict_flagged_per_step": strict_flagged, "strict_flag_rate": strict_flagged / float(world * B),
                      "codewords_differing_between_modes": differs_between_modes},
            "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": "codewords/s", "h2d_bytes_per_step": B * N * 4,
                    "d2h_bytes_per_step": B * KW * 4, "steps": e2e_steps, "matches_device_arm": e2e_ok,
                    "pipelined_chunks": code.info(7), "mode": mode, "host_buffer": args.host_alloc,
                    "h2d_copy_gbs_per_gpu_all_ranks_at_once": B * N * 4 / (h2d_ms * 1e-3) / 1e9,
                    "h2d_needed_gbs_per_gpu_at_device_rate": B * N * 4 / (ms_step * 1e-3) / 1e9},
            "e2e_sweep": {"value": world * B * sweep_steps / sweep_s, "unit": "codewords/s",
                          "what": "polar_b200_bler_sweep: Philox info bits + encoder + AWGN (double) + decode + block-error count on "
                                  "the device, fresh codewords every step; per step only the index range goes in and the counters come out",
                          "h2d_bytes_per_step": 8 * len(sweep) + 64, "d2h_bytes_per_step": 16 * len(sweep), "steps": sweep_steps,
                          "bler_last_step": float(sw_counts[0, :, 0].sum() / max(1, sw_counts[0, :, 1].sum())),
                          "collective": collective},
            "gpu_launches": int(launches),
            # informational, beside the contract's HBM roofline: the decoder is bound by instruction issue, so the same
            # rate is also stated against the issue-slot ceiling (148 SMs x 4 schedulers x SM clock), with the warp
            # instructions per codeword taken from the committed ncu capture (profiles/ncu_traffic.json)
            "issue_roofline": issue_roofline(traffic_src, value / world, sm_count, clocks),
            "clocks": clocks,
        }
        per_point = [{"ebno_db": float(sweep[e]), "n": int(counts[0, e, 1]), "err_gpu": int(counts[0, e, 0]),
                      "bler_gpu": float(counts[0, e, 0] / counts[0, e, 1]), "ci95": wilson(int(counts[0, e, 0]), int(counts[0, e, 1])),
                      "err_gpu_%s_mode" % other: int(counts[1, e, 0])}
                     for e in range(ns)]
        out["bler"] = float(counts[0, :, 0].sum() / counts[0, :, 1].sum())
        if world == 1 and not args.no_cpu:
            kind, dec = cpu_decoder(n, K, crc)
            threads = os.cpu_count() or 1
            v, S, want = time_cpu(dec, h_llr.numpy(), L, threads, args.cpu_seconds, want_out=True)
            out["cpu_baseline"] = {
                "value": v, "unit": "codewords/s", "cores": threads, "kind": kind,
                "sample": "first %d codewords of the same batch, %s, g++ -O2, one decoder object per thread" % (
                    S, "unmodified PolarC/PolarCode.cpp (oracle/_ref)" if kind == "reference" else "oracle port"),
                "table": cpu_table(dec, cpu_decoder_o3(n, K, crc) if kind == "reference" else None, h_llr.numpy(), L, threads,
                                   min(4.0, args.cpu_seconds))}
            # parity: the bits the CPU reference just decoded against the GPU's bits for the same codewords
            got = unpack_bits(d_out[:S].cpu().numpy().view(np.uint32), K)
            got_other = unpack_bits(d_out2[:S].cpu().numpy().view(np.uint32), K)
            mm, mmo = (got != want).any(1), (got_other != want).any(1)
            err_ref, err_gpu = (want != info[:S]).any(1), (got != info[:S]).any(1)
            out["parity"] = {"against": kind, "compared": int(S), "mismatches": int(mm.sum()), "mode": mode,
                             "mismatches_that_are_block_errors_in_both": int((mm & err_ref & err_gpu).sum()),
                             "mismatches_in_%s_mode" % other: int(mmo.sum())}
            for e in range(ns):
                sel = (np.arange(S) % ns) == e
                per_point[e].update({"n_ref_sample": int(sel.sum()), "err_ref_sample": int(err_ref[sel].sum()),
                                     "err_gpu_same_sample": int(err_gpu[sel].sum())})
        out["bler_per_ebno"] = per_point
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="codewords per GPU per step (default 65536; 4096 for c1)")
    ap.add_argument("--mode", default=os.environ.get("POLAR_B200_MODE", "strict"), choices=["strict", "fp32", "f64", "minsum"],
                    help="arithmetic of the headline arm (the other of strict / fp32 is reported in `modes`)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--host-alloc", default="pinned", choices=["pinned", "wc"],
                    help="host LLR buffer of the end-to-end arm: torch pinned memory, or write-combined from the library")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.batch is None:
        args.batch = 4096 if args.config == "c1" else 65536
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
