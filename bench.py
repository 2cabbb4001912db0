#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native streaming-FIR hot path.

Metric (BASELINE.json): Msamples/s of the 128-tap decimate-by-8 FIR (real taps on complex float32 data,
fastDecimatorC, reference hs_sources/SDR/Filter.hs:352-356 -> c_sources/decimate.c:105) over 8192-sample buffers.

Workload: a 2^28-sample synthetic white-noise IQ stream = 32768 buffers of 8192 samples (configs[1] shape streamed;
it is also configs[4], the multi-GPU stream).  One STEP = one pass of the decimator over the whole stream.
  * `value`   : device-resident pass (input already in HBM, output to HBM), whole job over all N GPUs.  With N > 1 the
                stream is sharded in overlapping chunks and the T-D boundary samples travel by NCCL (strong scaling).
  * `e2e`     : the same stream fed from pinned HOST memory through the reference-facing Pipes boundary
                (sdr_pipe_run over firDecimator: 8192-sample input vectors in, 8192-sample output vectors out),
                host->device and device->host copies inside the timed region.
  * `roofline`: algorithmic HBM bytes (9 B per input sample: 8 read + 8/8 written) / measured duration of the
                decimator launch, against the measured copy bandwidth in MEASURED_PEAKS.json.
  * `cpu_baseline`: the reference's own AVX C path (oracle/_ref, compiled unmodified) timed on this box's host cores
                exactly as firDecimator issues it, on a bounded sample.
`--impl reference` times that CPU path alone with every host thread.

Only this file's cpu_baseline / --impl reference legs touch oracle/.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "Msamples/sec FIR decimate-by-8 (128 taps, 8192-buf); %HBM roofline"
UNIT = "Msamples/s"
LOG2_STREAM = 28
TAPS, FACTOR, BUF = 128, 8, 8192
ALGO_BYTES_PER_SAMPLE = 8.0 + 8.0 / FACTOR   # each input read once, each output written once


def design_taps():
    """the package's tap designer (sdr_b200/filterdesign.py: the Hamming-windowed sinc of SDR.FilterDesign), loaded by path so
    that the reference arm does not need the native library"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_sdr_b200_filterdesign", os.path.join(ROOT, "sdr_b200", "filterdesign.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.windowed_sinc_taps(TAPS, 1.0 / (2 * FACTOR))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """per-launch DRAM bytes of the dominant kernel from the committed ncu capture, scaled to this run's launch"""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled from a thread every millisecond (the
    timed region of the default run is only ~8 ms long, an `nvidia-smi -lms 100` loop would see it once or twice);
    nvidia-smi is the fallback when the NVML binding is unusable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
                    (0x80, "hw_power_brake_slowdown"))

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.stop_flag, self.extra = [], None, None, False, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    self.extra.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_MEM)),
                                       n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0, time.time()))
                except Exception:
                    pass
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.time(), sm, mask))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        """what the poll saw between two wall-clock instants (the sampler keeps running)"""
        if not self.nvml:
            return None
        rows = [r for r in list(self.rows) if t0 <= r[0] <= t1]
        if not rows:
            return None
        mask = 0
        for r in rows:
            mask |= r[2]
        res = {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": self.max_mhz, "sm_mhz_min": min(r[1] for r in rows),
               "reasons": sorted(name for bit, name in self.NVML_REASONS if mask & bit), "samples": len(rows), "source": "nvml, 1 ms poll"}
        ex = [e for e in list(self.extra) if t0 <= e[2] <= t1]
        if ex:
            res["mem_mhz"] = statistics.median(e[0] for e in ex)
            res["power_w_median"] = statistics.median(e[1] for e in ex)
        return res

    def stop(self, t0, t1):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
            if not rows:
                return None
            mask = 0
            for r in rows:
                mask |= r[2]
            res = {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": self.max_mhz, "sm_mhz_min": min(r[1] for r in rows),
                   "reasons": sorted(name for bit, name in self.NVML_REASONS if mask & bit), "samples": len(rows), "source": "nvml, 1 ms poll"}
            ex = [e for e in self.extra if t0 <= e[2] <= t1]
            if ex:
                res["mem_mhz"] = statistics.median(e[0] for e in ex)
                res["power_w_median"] = statistics.median(e[1] for e in ex)
                res["power_w_max"] = max(e[1] for e in ex)
            return res
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (the only users of oracle/)
# ---------------------------------------------------------------------------------------------------------------------
_CPU_STREAM = {}


def cpu_stream(n_vectors):
    """the CPU legs' bounded sample of the stream: n_vectors x 8192 complex samples of the same counter-based white
    noise, generated once.  4096 vectors = 256 MiB: like the 2 GiB workload (and unlike a few-MB loop) it does not
    stay in the host's last-level cache, so the CPU path streams from DRAM as firDecimator would."""
    import synth
    if n_vectors not in _CPU_STREAM:
        parts = [synth.noise(2 * BUF * 256, first=2 * BUF * v) for v in range(0, n_vectors, 256)]
        _CPU_STREAM[n_vectors] = np.concatenate(parts)[:2 * BUF * n_vectors]
    return _CPU_STREAM[n_vectors]


def cpu_run(threads, seconds, n_vectors=4096):
    import oracle
    port = oracle.port()
    ref = oracle.ref()
    lib = port.lib
    lib.o_bench_fir_decimator.restype = C.c_double
    lib.o_bench_fir_decimator.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int,
                                          C.c_int, C.c_double, C.POINTER(C.c_long), C.POINTER(C.c_double)]
    if ref is not None:
        fn, kind = C.cast(ref.lib.decimateAVXRC, C.c_void_p), "reference"
    else:
        fn, kind = C.cast(lib.o_port_decimateAVXRC, C.c_void_p), "port"
    taps = design_taps()
    dup = np.repeat(taps, 2).astype(np.float32)
    n_vectors = max(n_vectors, threads * 8)
    x = cpu_stream(n_vectors)
    done, secs = C.c_long(), C.c_double()
    rate = lib.o_bench_fir_decimator(fn, FACTOR, TAPS, dup.ctypes.data, taps.ctypes.data, x.ctypes.data, n_vectors, BUF,
                                     threads, seconds, C.byref(done), C.byref(secs))
    return {"value": rate / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{done.value} input samples ({done.value // BUF} x {BUF}-sample vectors from a {n_vectors}-vector "
                      f"({n_vectors * BUF * 8 >> 20} MiB, DRAM-resident) white-noise stream, firDecimator call pattern: decimateAVXRC 1009 outputs + 15 crossover outputs per vector) in "
                      f"{secs.value:.1f} s"}


def cpu_run_u8(threads, seconds, n_vectors=16384):
    """the reference's own CPU path for the u8-fed chain: convertCAVX (convert.c:37) + decimateAVXRC (decimate.c:105)
    per 16384-byte vector, as `P.map interleavedIQUnsignedByteToFloatFast >-> firDecimator` issues them"""
    import oracle
    import synth
    port = oracle.port()
    ref = oracle.ref()
    lib = port.lib
    lib.o_bench_u8_fir_decimator.restype = C.c_double
    lib.o_bench_u8_fir_decimator.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
                                             C.c_int, C.c_int, C.c_double, C.POINTER(C.c_long), C.POINTER(C.c_double)]
    if ref is not None:
        conv, fn, kind = C.cast(ref.lib.convertCAVX, C.c_void_p), C.cast(ref.lib.decimateAVXRC, C.c_void_p), "reference"
    else:
        conv, fn, kind = C.cast(lib.o_convertC, C.c_void_p), C.cast(lib.o_port_decimateAVXRC, C.c_void_p), "port"
    taps = design_taps()
    dup = np.repeat(taps, 2).astype(np.float32)
    n_vectors = max(n_vectors, threads * 8)
    key = ("u8", n_vectors)
    if key not in _CPU_STREAM:   # 256 MiB of bytes (DRAM-resident like the f32 sample): a 16 MiB piece of the counter stream, tiled
        piece = synth.rand_bytes(1 << 24)
        _CPU_STREAM[key] = np.tile(piece, (2 * BUF * n_vectors + len(piece) - 1) // len(piece))[:2 * BUF * n_vectors]
    x = _CPU_STREAM[key]
    done, secs = C.c_long(), C.c_double()
    rate = lib.o_bench_u8_fir_decimator(conv, fn, FACTOR, TAPS, dup.ctypes.data, taps.ctypes.data, x.ctypes.data, n_vectors, BUF,
                                        threads, seconds, C.byref(done), C.byref(secs))
    return {"value": rate / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{done.value} input samples ({done.value // BUF} x {2 * BUF}-byte u8 IQ vectors from a {n_vectors}-vector "
                      f"({n_vectors * BUF * 2 >> 20} MiB) stream; per vector convertCAVX + decimateAVXRC 1009 outputs + 15 crossover outputs) in "
                      f"{secs.value:.1f} s"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = 2.0
    for _ in range(args.warmup):
        cpu_run(threads, 0.3)
    vals = [cpu_run(threads, per_step) for _ in range(max(1, args.steps))]
    best = max(vals, key=lambda v: v["value"])
    mean = sum(v["value"] for v in vals) / len(vals)
    # second record: the u8-fed chain (convertCAVX + decimateAVXRC), same threads, a bounded sample
    u8 = None
    try:
        cpu_run_u8(threads, 0.3)
        u8v = [cpu_run_u8(threads, per_step) for _ in range(min(3, max(1, args.steps)))]
        u8 = dict(max(u8v, key=lambda v: v["value"]), value=sum(v["value"] for v in u8v) / len(u8v),
                  h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    except Exception as e:   # an old prebuilt liboracle.so without the u8 driver
        u8 = {"unavailable": str(e)}
    line = {"impl": "reference", "metric": METRIC, "value": mean, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "taps": TAPS, "decimation": FACTOR, "buffer": BUF,
                       "note": "each step is a bounded 2 s sample of the stream on all host threads"},
            "cpu_baseline": dict(best, value=mean),
            "e2e": {"value": mean, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "e2e_u8": u8}
    print(json.dumps(line), flush=True)


def workload_name():
    return (f"fastDecimatorC decimate-by-{FACTOR}, {TAPS} real taps on complex f32, 2^{LOG2_STREAM}-sample white-noise IQ stream as "
            f"{(1 << LOG2_STREAM) // BUF} x {BUF}-sample buffers")


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def timed_passes(ctx, comm, dist, world, fn, n_calls, sdr_b200, marks=None, every=0):
    """device time of n_calls back-to-back calls of fn on the library's stream.  All ranks are lined up IN-STREAM first
    (a 4-byte ncclAllReduce on the same stream, sdr_comm_barrier): no rank's span contains another rank's host start-up."""
    ctx.sync()
    if world > 1:
        dist.barrier()
        comm.barrier()
    e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    e0.record()
    evs = []
    for i in range(n_calls):
        fn()
        if every and (i + 1) % every == 0 and i + 1 < n_calls:
            ev = sdr_b200.Event(ctx); ev.record(); evs.append(ev)
    e1.record()
    ms = e0.elapsed_ms(e1)
    if marks is not None:   # device time of every step inside the region (shows clock / power drift over the region)
        prev = e0
        for ev in evs + [e1]:
            marks.append(prev.elapsed_ms(ev)); prev = ev
    for ev in evs:
        ev.destroy()
    e0.destroy(); e1.destroy()
    return ms


def gather_max(dist, world, ms):
    """(max over ranks, per-rank list)"""
    if world == 1:
        return ms, [ms]
    import torch
    parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.tensor([ms], dtype=torch.float64))
    per = [float(p.item()) for p in parts]
    return max(per), per


def run_gpu(args, rank, world, local_rank, dist):
    import sdr_b200
    from sdr_b200 import _lib as L
    from sdr_b200 import multigpu

    n = 1 << args.log2n
    K = max(1, args.passes if args.passes > 0 else 3 * world)
    warmup = max(args.warmup, 3)
    pin_rank_to_gpu_numa(local_rank)
    # NVML is initialised and polling on EVERY rank before anything is timed (round 1 started it on rank 0 only, between
    # the barrier and the first event: the other ranks' spans then contained rank 0's NVML start-up)
    sampler = ClockSampler(local_rank)
    ctx = sdr_b200.Context(local_rank)
    taps = design_taps()
    dec = sdr_b200.cudaDecimatorC(FACTOR, taps, ctx=ctx, sizeMultiple=4)
    plan = multigpu.shard_plan(n, TAPS, FACTOR, world, rank)
    comm = None
    if world > 1:
        import torch
        uid = multigpu.unique_id() if rank == 0 else bytes(L.COMM_ID_BYTES)
        t = torch.tensor(list(uid), dtype=torch.uint8)
        dist.broadcast(t, 0)
        comm = multigpu.Comm(ctx, bytes(t.tolist()), world, rank)

    # this rank's chunk of the stream, generated in place from the counter RNG (keyed on the global sample index)
    d_in = ctx.alloc(8 * plan.in_count + 256)
    d_out = ctx.alloc(8 * max(plan.out_count, 1) + 256)
    ctx.synth_noise(d_in, 2 * plan.in_count, first_float=2 * plan.in_begin)
    ctx.sync()

    def step_pass():
        multigpu.decimate_sharded(dec, comm, plan, d_in.ptr, d_out.ptr)

    halo = "none"
    if world > 1:
        halo = "nccl send/recv per pass (side stream) + boundary launch"
        if args.halo == "peer":
            dist.barrier()   # every chunk complete before a neighbour may read it
            try:
                comm.share_chunks(d_in.ptr)
                halo = ("peer memory: the ring kernel's edge fills read the T-D halo samples in place from the right neighbour's "
                        "HBM over NVLink (CUDA IPC mapping, TMA bulk copies); one launch per pass, no rendezvous")
            except sdr_b200.SdrError as e:
                if rank == 0:
                    print(f"peer-memory halo unavailable ({e}); using NCCL", file=sys.stderr)

    def all_clocks(c):
        """rank 0's sample, with every rank's median clock and the union of the reasons"""
        if world == 1:
            return c
        allc = [None] * world
        dist.all_gather_object(allc, c)
        if c is not None:
            c["per_rank_sm_mhz"] = [x.get("sm_mhz") if x else None for x in allc]
            rs = set()
            for x in allc:
                rs |= set((x or {}).get("reasons", []))
            c["reasons"] = sorted(rs)
        return c

    # ---- the timed region: `steps` steps of K passes right after `warmup` steps, from a cool start (see `regime`) ----
    for _ in range(warmup * K):
        step_pass()
    launches0 = ctx.launches
    t_wall0 = time.time()
    step_ms = []
    ms = timed_passes(ctx, comm, dist, world, step_pass, args.steps * K, sdr_b200, marks=step_ms, every=K)
    ctx.sync()
    t_wall1 = time.time()
    launches = ctx.launches - launches0
    kernel_name = dec.last_kernel()
    ms_max, ms_ranks = gather_max(dist, world, ms)
    clocks = all_clocks(sampler.window(t_wall0, t_wall1) or sampler.window(t_wall0 - 0.01, t_wall1 + 0.01))
    if world > 1:
        import torch
        tl = torch.tensor([launches], dtype=torch.int64)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches_total = int(tl.item())
    else:
        launches_total = launches

    # ---- the same measurement in the SUSTAINED regime: ~150 ms of back-to-back passes first, so that the board's power
    # management has settled (this kernel keeps HBM and the FP32 pipe busy at once and runs into the power cap after
    # ~50 ms: profiles/r02_power_regime.txt), then the same number of passes again ----
    sustained = None
    if not args.no_sustained:
        pre = int(150.0 / max(1e-3, ms_max / (args.steps * K))) + 1
        for _ in range(pre):
            step_pass()
        t0s = time.time()
        ms_s = timed_passes(ctx, comm, dist, world, step_pass, args.steps * K, sdr_b200)
        ctx.sync()
        t1s = time.time()
        ms_s_max, _ = gather_max(dist, world, ms_s)
        sustained = {"ms_per_pass": ms_s_max / (args.steps * K), "preheat_passes": pre,
                     "clocks": all_clocks(sampler.window(t0s, t1s) or sampler.window(t0s - 0.01, t1s + 0.01))}

    # ---- secondary: the NCCL transport north_star names (send/recv on a side stream + a boundary launch), same regime as
    # `sustained`, alternated with the peer transport ----
    nccl_ms_per_pass, peer_ms_again = None, None
    if world > 1 and args.halo == "peer" and comm.peer_halo_active(d_in.ptr):
        a, b = [], []
        for _ in range(2):
            comm.share_chunks(None)
            for _ in range(3):
                step_pass()
            a.append(gather_max(dist, world, timed_passes(ctx, comm, dist, world, step_pass, 2 * K, sdr_b200))[0] / (2 * K))
            ctx.sync(); dist.barrier()
            comm.share_chunks(d_in.ptr)
            b.append(gather_max(dist, world, timed_passes(ctx, comm, dist, world, step_pass, 2 * K, sdr_b200))[0] / (2 * K))
        nccl_ms_per_pass, peer_ms_again = min(a), min(b)
    ms_per_step = ms_max / args.steps
    ms_per_pass = ms_per_step / K
    value = n * K / (ms_per_step * 1e-3) / 1e6

    # the same pass WITHOUT the exchange (interior outputs of the resident chunk only), alternated with the sharded pass in
    # the same clock regime: what the halo costs
    halo_ms, interior_ms = None, None
    if world > 1:
        def interior_pass():
            L.check(L.lib.sdr_decimate_stream(dec.handle, d_in.ptr, plan.in_count, d_out.ptr, plan.out_interior))
        a, b = [], []
        for _ in range(3):
            a.append(gather_max(dist, world, timed_passes(ctx, comm, dist, world, step_pass, 2 * K, sdr_b200))[0] / (2 * K))
            b.append(gather_max(dist, world, timed_passes(ctx, comm, dist, world, interior_pass, 2 * K, sdr_b200))[0] / (2 * K))
        interior_ms = statistics.median(b)
        halo_ms = statistics.median(a) - interior_ms
        for _ in range(2):   # leave d_out holding the sharded result for the checksum
            step_pass()

    # this box's own D2D copy rate over a region as long as the timed one (MEASURED_PEAKS.json's figure is a best-of-10 burst)
    copy_gbs = None
    if rank == 0:
        nb = min(8 * plan.in_count // 2, 1 << 30)
        reps = max(4, int(ms / max(1e-3, 2.0 * nb / 6.5e9 * 1e3)))
        def copy_pass():
            L.check(L.lib.sdr_memcpy_d2d(ctx.h, d_in.at(nb), d_in.ptr, nb))
        for _ in range(3):
            copy_pass()
        ctx.sync()
        e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
        e0.record()
        for _ in range(reps):
            copy_pass()
        e1.record()
        copy_gbs = 2.0 * nb * reps / (e0.elapsed_ms(e1) * 1e-3) / 1e9
        ctx.synth_noise(d_in, 2 * plan.in_count, first_float=2 * plan.in_begin)   # the copy overwrote the chunk's upper half
        ctx.sync()
    if world > 1:
        dist.barrier()

    # size-independent parity property at full size: checksum of all shards' outputs == rank-independent value
    csum = ctx.checksum32(d_out, 2 * plan.out_count, first_word=2 * plan.out_begin)
    if world > 1:
        import torch
        lo = torch.tensor([csum & 0xffffffff, csum >> 32], dtype=torch.int64)
        parts = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(parts, lo)
        csum = sum(int(p[0]) + (int(p[1]) << 32) for p in parts) & 0xffffffffffffffff

    # ---- end to end through the Pipes boundary with host buffers (this rank's share of the stream + its halo) ----
    e2e, e2e_spot, e2e_u8 = None, None, None
    if not args.no_e2e:
        r = run_e2e(args, ctx, dec, plan, rank, world, dist, sdr_b200, L)
        if r is not None:
            e2e, e2e_spot = r
        e2e_u8 = run_e2e_u8(args, ctx, dec, plan, rank, world, dist, sdr_b200, L)

    # ---- roofline of the dominant kernel, from the MAX-over-ranks time and the largest chunk ----
    peak, peak_src = measured_peak()
    in_max = multigpu.shard_plan(n, TAPS, FACTOR, world, 0).in_count
    achieved = ALGO_BYTES_PER_SAMPLE * in_max / (ms_per_pass * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ((traffic or {}).get("dram_bytes_per_input_sample") or 0) * in_max or None,
                "kernel": kernel_name, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * in_max,
                "launches_per_pass": launches / (args.steps * K),
                "copy_GBps_this_box_same_duration": copy_gbs,
                "note": "per GPU: 9 B x the largest rank's chunk / (max-over-ranks device time per pass); the pass is ONE launch "
                        "(ragged end and, sharded, the boundary windows are computed by the ring kernel itself)"}
    if traffic:
        roofline["traffic_source"] = traffic.get("source")
    if sustained:
        sustained["value"] = n / (sustained["ms_per_pass"] * 1e-3) / 1e6
        sustained["unit"] = UNIT
        sustained["roofline_frac"] = ALGO_BYTES_PER_SAMPLE * in_max / (sustained["ms_per_pass"] * 1e-3) / 1e9 / peak

    cpu, configs, pipes_mode = None, None, None
    if rank == 0 and world == 1:
        if not args.no_configs:
            d_in.free(); d_out.free()
            configs = run_configs(args, ctx, dec, sdr_b200, L, peak)
            pipes_mode = run_pipes_mode(args, ctx, dec, sdr_b200, L)
        if not args.no_cpu:
            cpu = cpu_run(1, args.cpu_seconds)
            try:
                if e2e is not None:
                    e2e["spot_parity_vs_reference_avx"] = e2e_spot_parity(e2e_spot)
            except Exception:
                pass
    elif world > 1 and e2e is not None:
        # every rank checks 600 outputs of ITS OWN end-to-end run against the reference C; the line carries the AND
        try:
            ok = e2e_spot_parity(e2e_spot) if not args.no_cpu else None
        except Exception:
            ok = None
        oks = [None] * world
        dist.all_gather_object(oks, ok)
        e2e["spot_parity_vs_reference_avx"] = None if any(o is None for o in oks) else bool(all(oks))

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": workload_name() if args.log2n == LOG2_STREAM else f"reduced 2^{args.log2n}-sample stream (not the headline size)",
                           "taps": TAPS, "decimation": FACTOR, "buffer": BUF, "samples": n,
                           "passes_per_step": K,
                           "step": f"{K} back-to-back passes of the decimator over the whole 2^{args.log2n}-sample stream (3 x n_gpus: a step "
                                   f"lasts ~1.1-1.4 ms at every N; at 8 GPUs one pass is only ~50 us); value = samples x passes / step time",
                           "regime": "warm-up + timed region total ~30 ms of load from a cool start at every N: the kernel is timed the way "
                                     "MEASURED_PEAKS.json's burst copy figure was (this kernel keeps HBM and the FP32 pipe busy together and runs "
                                     "into the board power cap after ~50 ms of back-to-back passes; `sustained` is the same measurement after "
                                     "a 150 ms pre-heat)",
                           "l2": "inputs exceed L2 (per-GPU chunk %.0f MiB in + %.0f MiB out vs 126 MB L2)" % (
                               8 * plan.in_count / 2 ** 20, 8 * plan.out_count / 2 ** 20),
                           "sharding": "single GPU" if world == 1 else f"{world} overlapping chunks, halo of {TAPS - FACTOR} samples per boundary: {halo}",
                           "timing": "CUDA events on the library's stream after an in-stream ncclAllReduce lines the ranks up; max over ranks",
                           "arithmetic": "fp32 FMA, taps in increasing order; parity vs reference AVX path <= 1e-5 of output scale (tests/test_gpu_parity.py)",
                           "output_checksum": "%016x" % csum},
                "ms_per_pass": ms_per_pass, "ms_per_pass_by_rank": [m / (args.steps * K) for m in ms_ranks],
                "sustained": sustained,
                "ms_by_step_rank0": [round(m, 4) for m in step_ms],
                "halo_ms_per_pass": halo_ms, "interior_only_ms_per_pass": interior_ms,
                "peer_halo_ms_per_pass_same_regime": peer_ms_again, "nccl_halo_ms_per_pass": nccl_ms_per_pass,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_u8": e2e_u8, "configs": configs, "pipes_mode": pipes_mode,
                "gpu_launches": launches_total, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if comm:
        comm.close()


def run_e2e(args, ctx, dec, plan, rank, world, dist, sdr_b200, L):
    import synth
    n_local = plan.in_count + plan.halo           # host slice this rank feeds: its chunk plus the halo samples
    n_vecs = n_local // BUF                       # whole 8192-sample vectors (the ragged end cannot complete a window set)
    if n_vecs == 0:
        return None
    hin = sdr_b200.PinnedArray(np.float32, 2 * n_vecs * BUF)
    out_cap = (n_vecs * BUF // FACTOR // BUF + 1) * BUF
    hout = sdr_b200.PinnedArray(np.float32, 2 * out_cap)
    # fill the pinned input from the device generator (same counter RNG as the resident run)
    tmp = ctx.alloc(hin.array.nbytes)
    ctx.synth_noise(tmp, 2 * n_vecs * BUF, first_float=2 * plan.in_begin)
    L.check(L.lib.sdr_memcpy_d2h(ctx.h, hin.p, tmp.ptr, hin.array.nbytes))
    ctx.sync()
    tmp.free()
    n_out = C.c_longlong()

    # the stage is long-lived, as in a running receiver: successive passes continue one stream through the same pipe
    pipe = sdr_b200.pipeFirDecimator(dec, BUF)
    L.check(L.lib.sdr_pipe_set_batch(pipe.h, args.e2e_batch_vectors * BUF))

    def one_pass():
        L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, hin.p, BUF, n_vecs, L.SDR_HOST_PINNED, hout.p, out_cap, L.SDR_HOST_PINNED,
                                   C.byref(n_out)))
        return n_out.value

    # the PCIe ceiling of this box, measured the plain way: one pinned host-to-device copy of the whole input
    dtmp = ctx.alloc(hin.array.nbytes)
    ev0, ev1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    L.check(L.lib.sdr_memcpy_h2d(ctx.h, dtmp.ptr, hin.p, hin.array.nbytes))
    ev0.record()
    L.check(L.lib.sdr_memcpy_h2d(ctx.h, dtmp.ptr, hin.p, hin.array.nbytes))
    ev1.record()
    h2d_gbs = hin.array.nbytes / (ev0.elapsed_ms(ev1) * 1e-3) / 1e9
    dtmp.free()

    steps = max(1, min(args.steps, args.e2e_steps))
    got = one_pass()
    # keep a slice of the popped host vectors of the first pass (output m is window m of this rank's chunk): the CPU leg
    # checks it against the reference C (the only place this file touches oracle/)
    spot = None
    try:
        if got >= 4096:
            spot = {"first_output": 1000, "in_begin": int(plan.in_begin),
                    "y": np.array(hout.array[2 * 1000:2 * 1600], dtype=np.float32, copy=True)}
    except Exception:
        spot = None
    one_pass()
    ctx.sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    e0.record()
    for _ in range(steps):
        got = one_pass()
    e1.record()
    ms = e0.elapsed_ms(e1)
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = max(ms, wall_ms)   # the host loop is part of the end-to-end path
    if world > 1:
        import torch
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tn = torch.tensor([n_vecs * BUF], dtype=torch.int64)
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        total = int(tn.item())
    else:
        total = n_vecs * BUF
    res = {"value": total / (ms / steps * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(8 * n_vecs * BUF),
           "d2h_bytes_per_step": int(8 * got), "steps": steps,
           "api": "sdr_pipe_run(firDecimator, 8192-sample pinned host vectors in, 8192-sample host vectors out), "
                  f"sdr_pipe_set_batch = {args.e2e_batch_vectors} output vectors per launch",
           "spot_parity_vs_reference_avx": None,
           "pcie_h2d_GBps_plain_memcpy": h2d_gbs,
           "h2d_GBps_achieved": 8.0 * (n_vecs * BUF) / (ms / steps * 1e-3) / 1e9,
           "ceiling": {"value": h2d_gbs * 1e9 / 8.0 / 1e6 * (world if world > 1 else 1), "unit": UNIT,
                       "what": "plain pinned cudaMemcpy host-to-device rate of this rank's link / 8 B per complex f32 sample"
                               + (f", x {world} links" if world > 1 else "")
                               + ": no kernel can make a host-fed f32 stream faster than this"},
           "bound": "PCIe host-to-device: 8 B per input sample must cross the link"}
    pipe.close()
    hin.free()
    hout.free()
    return res, spot


def run_e2e_u8(args, ctx, dec, plan, rank, world, dist, sdr_b200, L):
    """the same decimator fed with what an SDR front end delivers: u8 IQ host vectors (2 B per sample over the link)
    through `P.map interleavedIQUnsignedByteToFloat >-> firDecimator` (examples/fm/fm.hs:34-36) as ONE fused stage
    (sdr_pipe_u8_decimator), complex f32 host vectors out"""
    n_local = plan.in_count + plan.halo
    n_vecs = n_local // BUF
    if n_vecs == 0:
        return None
    hin = sdr_b200.PinnedArray(np.uint8, 2 * n_vecs * BUF)
    out_cap = (n_vecs * BUF // FACTOR // BUF + 1) * BUF
    hout = sdr_b200.PinnedArray(np.float32, 2 * out_cap)
    tmp = ctx.alloc(hin.array.nbytes)
    ctx.synth_bytes(tmp, 2 * n_vecs * BUF, first_byte=2 * plan.in_begin)
    L.check(L.lib.sdr_memcpy_d2h(ctx.h, hin.p, tmp.ptr, hin.array.nbytes))
    ctx.sync()
    tmp.free()
    n_out = C.c_longlong()
    pipe = sdr_b200.pipeU8Decimator(dec, BUF)
    L.check(L.lib.sdr_pipe_set_batch(pipe.h, args.e2e_batch_vectors * BUF))

    def one_pass():
        L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, hin.p, 2 * BUF, n_vecs, L.SDR_HOST_PINNED, hout.p, out_cap, L.SDR_HOST_PINNED,
                                   C.byref(n_out)))
        return n_out.value

    steps = max(1, min(args.steps, args.e2e_steps))
    got = one_pass()
    spot = None
    if got >= 4096:
        spot = {"first_output": 1000, "in_begin": int(plan.in_begin),
                "y": np.array(hout.array[2 * 1000:2 * 1600], dtype=np.float32, copy=True),
                "bytes": np.array(hin.array[2 * 1000 * FACTOR:2 * (1600 * FACTOR + TAPS)], dtype=np.uint8, copy=True)}
    kernel = L.lib.sdr_pipe_last_kernel(pipe.h).decode()
    one_pass()
    ctx.sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    e0.record()
    for _ in range(steps):
        got = one_pass()
    e1.record()
    ms = max(e0.elapsed_ms(e1), (time.perf_counter() - t0) * 1e3)
    if world > 1:
        import torch
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tn = torch.tensor([n_vecs * BUF], dtype=torch.int64)
        dist.all_reduce(tn, op=dist.ReduceOp.SUM)
        total = int(tn.item())
    else:
        total = n_vecs * BUF
    ok = None
    if rank == 0 and not args.no_cpu and spot is not None:
        try:
            ok = e2e_u8_spot_parity(spot)
        except Exception:
            ok = None
    res = {"value": total / (ms / steps * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(2 * n_vecs * BUF),
           "d2h_bytes_per_step": int(8 * got), "steps": steps, "kernel": kernel,
           "api": "sdr_pipe_run(sdr_pipe_u8_decimator = P.map interleavedIQUnsignedByteToFloat >-> firDecimator fused; 16384-byte "
                  "pinned host u8 IQ vectors in, 8192-sample complex f32 host vectors out), "
                  f"sdr_pipe_set_batch = {args.e2e_batch_vectors} output vectors per launch",
           "workload": f"the headline stream as u8 IQ (what the RTL-SDR source yields, RTLSDRStream.hs:54-57): {n_vecs} x {BUF}-sample vectors per rank",
           "spot_parity_vs_reference_convert_plus_avx": ok,
           "link_GBps_achieved": (2.0 + 8.0 / FACTOR) * (n_vecs * BUF) / (ms / steps * 1e-3) / 1e9}
    pipe.close()
    hin.free()
    hout.free()
    return res


def e2e_u8_spot_parity(spot):
    """CPU leg: 600 outputs of the u8-fed end-to-end run against convertCAVX + decimateAVXRC of the same bytes"""
    import oracle
    ref = oracle.ref()
    if ref is None:
        return None
    xs = ref.convert_u8("convertCAVX", spot["bytes"]).view(np.complex64)
    want = ref.decimate("decimateAVXRC", 600, FACTOR, np.repeat(design_taps(), 2), xs)
    got = spot["y"].view(np.complex64)
    scale = np.maximum(np.abs(want), np.sqrt(np.mean(np.abs(want) ** 2)))
    return bool(np.all(np.abs(got - want) <= 1e-5 * scale))


# FP32 pipe peak measured on this pool's B200 (profiles/r01_mb_fma.txt, tools/mb_fma.cu: FFMA2 with a broadcast operand,
# 67.1 TFLOP/s = 33.5 T lane-FMA/s); the secondary roofline of the FIR kernels
FP32_PEAK_TFMA = 33.5


def run_configs(args, ctx, dec, sdr_b200, L, peak):
    """the other BASELINE.json configs and the element-wise stages, device resident, each with its own roofline"""
    log2 = min(args.log2n, 27)
    n = 1 << log2                                   # complex samples in the buffers (2n floats for the real kernels)
    x = ctx.alloc(8 * n + 256)
    y = ctx.alloc(8 * n + 256)
    ctx.synth_noise(x, 2 * n)
    out = {}

    def timed(fn, steps=8, warm=3):
        for _ in range(warm):
            fn()
        ctx.sync()
        e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        ms = e0.elapsed_ms(e1) / steps
        e0.destroy(); e1.destroy()
        return ms

    def put(key, what, ms, samples, bytes_per, fma_per, kernel, unit_name="input samples"):
        rate = samples / (ms * 1e-3)
        gbs = rate * bytes_per / 1e9
        e = {"what": what, "ms": ms, "value": rate / 1e6, "unit": "Msamples/s", "samples_per_launch": samples, "kernel": kernel,
             "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                          "algorithmic_bytes_per_sample": bytes_per}}
        if fma_per:
            e["roofline"]["fp32_fma_per_sample"] = fma_per
            e["roofline"]["fp32_frac"] = rate * fma_per / 1e12 / FP32_PEAK_TFMA
            if e["roofline"]["fp32_frac"] > e["roofline"]["frac"]:
                e["roofline"]["bound"] = "fp32 pipe (see fp32_frac; measured FFMA2 peak 33.5 TFMA/s)"
        out[key] = e

    nr = 2 * n
    # cfg1: fastFilterSymR, 64 taps (32 half taps), real (Filter.hs:258-261 -> filter.c:60)
    half = sdr_b200.windowed_sinc_taps(64, 1 / 4)[:32]
    f = sdr_b200.cudaFilterSymR(half, ctx=ctx)
    num = nr - 64 + 1
    ms = timed(lambda: L.check(L.lib.sdr_filter_stream(f.handle, x.ptr, nr, y.ptr, num)))
    put("cfg1", "fastFilterSymR 64-tap symmetric real FIR, 2^%d-sample real stream" % (log2 + 1), ms, nr, 8.0, 64, f.last_kernel())
    # cfg3: fastResamplerR 3/10, 90 taps, real (Filter.hs:468-473 -> resample.c:70)
    t_res = sdr_b200.windowed_sinc_taps(90, 1 / 20, gain=3.0)
    r = sdr_b200.cudaResamplerR(3, 10, t_res, ctx=ctx, sizeMultiple=8)
    num = (nr * 3 - r.numCoeffsR) // 10 + 1
    ms = timed(lambda: L.check(L.lib.sdr_resample_stream(r.handle, x.ptr, nr, y.ptr, num)))
    put("cfg3", "fastResamplerR 3/10 rational polyphase, 90 taps, 2^%d-sample real stream" % (log2 + 1), ms, nr, 5.2, 9, r.last_kernel())
    # element-wise stages
    nbytes = 2 * n
    bbuf = ctx.alloc(nbytes + 256)
    ctx.synth_bytes(bbuf, nbytes)
    ms = timed(lambda: L.check(L.lib.sdr_dev_convert_u8(ctx.h, bbuf.ptr, y.ptr, nbytes)))
    put("convert", "interleavedIQUnsignedByteToFloat (convert.c:37), per IQ pair", ms, n, 10.0, 0, "k_convert_u8_vec")
    ms = timed(lambda: L.check(L.lib.sdr_dev_fm_demod(ctx.h, 0.0, 0.0, x.ptr, y.ptr, n)))
    put("fmDemod", "fmDemod (Demod.hs:40), per complex sample", ms, n, 12.0, 0, "k_fm_demod4")
    ms = timed(lambda: L.check(L.lib.sdr_dev_scale(ctx.h, 0.2, x.ptr, y.ptr, nr)))
    put("scale", "scale (scale.c:30), per float", ms, nr, 8.0, 0, "k_scale_vec")
    d_fin = ctx.alloc(8)
    ms = timed(lambda: ctx.dc_blocker(x.ptr, y.ptr, nr, d_fin.ptr), steps=4, warm=2)
    _, par = ctx.dc_stats()
    put("dcBlocker", "dcBlocker (filter.c:152), chunk-parallel speculation, bit-exact, per float", ms, nr, 8.0, 0,
        "k_dc_spec_tiles + k_dc_repair" if par else "k_dc_blocker")
    d_fin.free()
    # cfg4: the whole FM chain, u8 IQ in -> audio out, 32 MiB pushes (16 Mi IQ pairs each) of device-resident vectors read in
    # place (SDR_DEVICE_HELD): two fused stages = front end and low-rate end.  The stream is 1 GiB (32 pushes) per pass so
    # that the first launch's latency and the closing synchronisation of sdr_pipe_run weigh what they weigh in a running
    # receiver; `x` is reused as the byte buffer.
    fil = sdr_b200.cudaFilterSymR(half, ctx=ctx)
    n_out = C.c_longlong()
    push = 1 << 25
    cbytes = 8 * n
    nc = cbytes // 2                                # IQ pairs per pass
    ctx.synth_bytes(x, cbytes)

    def time_chain(stages, label, key, thresholds=None):
        for a, b in zip(stages, stages[1:]):
            a.connect(b)
        for i, p in enumerate(stages):
            try:
                L.check(L.lib.sdr_pipe_set_batch(p.h, thresholds[i] if thresholds else 1 << 21))
            except sdr_b200.SdrError:
                pass

        def chain():
            L.check(L.lib.sdr_pipe_run(stages[0].h, stages[-1].h, x.ptr, push, cbytes // push, L.SDR_DEVICE_HELD, y.ptr, 2 * n, L.SDR_DEVICE,
                                       C.byref(n_out)))
        for _ in range(2):
            chain()
        ctx.sync()
        l0 = ctx.launches
        ms = timed(chain, steps=4, warm=0)
        put(key, label, ms, nc, 2.15, 32 + 9 / 8 + 64 * 3 / 80, " + ".join(
            k for k in (L.lib.sdr_pipe_last_kernel(st.h).decode() for st in stages) if k != "none"))
        out[key]["launches_per_push"] = (ctx.launches - l0) / 4 / (cbytes // push)
        out[key]["pushes_per_pass"] = cbytes // push
        out[key]["audio_samples_out"] = int(n_out.value)
        for p in stages:
            p.close()

    time_chain([sdr_b200.pipeFmFrontEnd(dec, BUF), sdr_b200.pipeFmLowRate(r, BUF, fil, BUF, 0.2)],
               "full FM pipe: u8 IQ -> [convert + decimate-by-8 (128 taps) + fmDemod] -> [resample 3/10 (90 taps) + 64-tap filter + x0.2], two fused "
               "stages, 32 MiB device pushes read in place, the front end launches at every push; per input IQ sample", "cfg4_chain")
    time_chain([sdr_b200.pipeFmFrontEnd(dec, BUF), sdr_b200.pipeFmLowRate(r, BUF, fil, BUF, 0.2)],
               "the same two stages and 32 MiB pushes with launch thresholds sized for throughput (sdr_pipe_set_batch: front end every 4 pushes, "
               "low-rate end every ~7): what a receiver that can afford 30 ms of latency would set", "cfg4_chain_batched",
               thresholds=[1 << 23, 1 << 22])
    time_chain([sdr_b200.pipeFmFrontEnd(dec, BUF), sdr_b200.pipeFirResampler(r, BUF), sdr_b200.pipeFirFilter(fil, BUF), sdr_b200.pipeScale(0.2, ctx)],
               "the same chain with the low-rate end as three separate stages (round 1's form)", "cfg4_chain_unfused_lowrate")
    # the fused front end alone (u8 IQ -> phase)
    fe = sdr_b200.pipeFmFrontEnd(dec, BUF)
    L.check(L.lib.sdr_pipe_set_batch(fe.h, 1 << 23))

    def front():
        L.check(L.lib.sdr_pipe_run(fe.h, fe.h, bbuf.ptr, nbytes, 1, L.SDR_DEVICE_HELD, y.ptr, 2 * n, L.SDR_DEVICE, C.byref(n_out)))
    ms = timed(front, steps=4, warm=2)
    put("cfg4_front", "fused front end alone: u8 IQ -> convert -> decimate-by-8 -> fmDemod, one push", ms, n, 2.5, 32, L.lib.sdr_pipe_last_kernel(fe.h).decode())
    fe.close()
    for b in (bbuf, x, y):
        b.free()
    return out


def run_pipes_mode(args, ctx, dec, sdr_b200, L):
    """device-resident Pipes mode, vector by vector: 8192-sample device vectors pushed through firDecimator by the native
    loop (sdr_pipe_run), launch threshold = 1 / 8 / 256 output vectors (sdr_pipe_set_batch)"""
    n = 1 << min(args.log2n, 26)
    x = ctx.alloc(8 * n + 256)
    y = ctx.alloc(n + 8 * BUF + 256)
    ctx.synth_noise(x, 2 * n)
    res = {"vector": BUF, "samples": n, "unit": "Msamples/s"}
    n_out = C.c_longlong()
    for mem, tag in ((L.SDR_DEVICE_HELD, ""), (L.SDR_DEVICE, "_copied")):
        for batch in (0, 8, 256):
            pipe = sdr_b200.pipeFirDecimator(dec, BUF)
            L.check(L.lib.sdr_pipe_set_batch(pipe.h, batch * BUF))

            def run():
                L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, x.ptr, BUF, n // BUF, mem, y.ptr, n // FACTOR + BUF, L.SDR_DEVICE, C.byref(n_out)))
            for _ in range(2):
                run()
            ctx.sync()
            l0 = ctx.launches
            t0 = time.perf_counter()
            e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
            e0.record()
            for _ in range(3):
                run()
            e1.record()
            ms = max(e0.elapsed_ms(e1), (time.perf_counter() - t0) * 1e3) / 3
            res[f"batch_{batch}{tag}"] = {"value": n / (ms * 1e-3) / 1e6, "ms": ms, "launches_per_pass": (ctx.launches - l0) / 3,
                                          "input_vectors_per_launch": (n // BUF) / max(1.0, (ctx.launches - l0) / 3)}
            pipe.close()
    # the persistent consumer: no launches at all, a push is one store the resident kernel polls.  A longer stream (2^28
    # samples = 32768 vectors) so that opening / closing the session (~0.1 ms) is amortised as it is in a running receiver.
    x.free(); y.free()
    n2 = 1 << min(args.log2n, 28)
    x = ctx.alloc(8 * n2 + 256)
    y = ctx.alloc(n2 + 8 * BUF + 256)
    ctx.synth_noise(x, 2 * n2)
    pipe = sdr_b200.pipeFirDecimator(dec, BUF)
    L.check(L.lib.sdr_pipe_set_persistent(pipe.h, n2))

    def run_p():
        L.check(L.lib.sdr_pipe_run(pipe.h, pipe.h, x.ptr, BUF, n2 // BUF, L.SDR_DEVICE_HELD, y.ptr, n2 // FACTOR + BUF, L.SDR_DEVICE, C.byref(n_out)))
    for _ in range(2):
        run_p()
    ctx.sync()
    l0 = ctx.launches
    t0 = time.perf_counter()
    for _ in range(3):
        run_p()
    ctx.sync()
    ms = (time.perf_counter() - t0) * 1e3 / 3
    res["persistent"] = {"value": n2 / (ms * 1e-3) / 1e6, "ms": ms, "samples": n2, "vectors_per_pass": n2 // BUF,
                         "launches_per_pass": (ctx.launches - l0) / 3, "kernel": "dec_c_ring_persist<128,8,8,32>",
                         "note": "wall clock per pass of 8192-sample SDR_DEVICE_HELD pushes, one vector per push (sdr_pipe_set_persistent): ONE "
                                 "resident kernel per pass consumes them as they are published; includes opening and closing the session, "
                                 "the ordinary launch for what it leaves and the device-to-device copies of the yielded vectors"}
    pipe.close()
    res["note"] = ("batch_N: launch threshold of N output vectors (N = 0: as soon as one 8192-sample output vector completes = every 8 input "
                   "vectors); default rows push SDR_DEVICE_HELD vectors (read in place, zero-copy), *_copied rows push SDR_DEVICE vectors "
                   "(one device-to-device copy per vector into the stage, round 1's only mode)")
    x.free(); y.free()
    return res


def pin_rank_to_gpu_numa(local_rank):
    """bind this process to the CPUs NVML reports as local to its GPU BEFORE any page-locked memory is allocated (first
    touch then places the pinned buffers on the GPU's NUMA node).  Best effort: silently a no-op where the topology is
    not exposed (VMs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def e2e_spot_parity(spot):
    """CPU leg: 600 outputs of the end-to-end run against decimateAVXRC on the CPU-regenerated stream"""
    import oracle
    import synth
    ref = oracle.ref()
    if ref is None or spot is None:
        return None
    m0 = spot["first_output"]
    xs = synth.noise_complex(600 * FACTOR + TAPS, first=spot["in_begin"] + m0 * FACTOR)
    want = ref.decimate("decimateAVXRC", 600, FACTOR, np.repeat(design_taps(), 2), xs)
    got = spot["y"].view(np.complex64)
    scale = np.maximum(np.abs(want), np.sqrt(np.mean(np.abs(want) ** 2)))
    return bool(np.all(np.abs(got - want) <= 1e-5 * scale))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2_STREAM, help="stream length (default 2^28, the headline workload)")
    ap.add_argument("--halo", default="peer", choices=["nccl", "peer"],
                    help="multi-GPU halo transport: in-kernel peer-memory reads over NVLink (default; NCCL is timed beside it) "
                         "or NCCL send/recv per pass")
    ap.add_argument("--passes", type=int, default=0,
                    help="back-to-back passes over the stream per step; default 3 x n_gpus, so that a step lasts ~1.1-1.4 ms at every N "
                         "(one pass is ~0.36 ms on 1 GPU, ~0.05 ms on 8) and every N is measured in the same power/clock regime")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config / per-stage device-resident measurements")
    ap.add_argument("--no-sustained", action="store_true", help="skip the second measurement after a 150 ms pre-heat")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-batch-vectors", type=int, default=0,
                    help="sdr_pipe_set_batch: output vectors per launch in the end-to-end run; default 256 / n_gpus (256 x 8192 outputs = "
                         "128 MiB of input per launch on one GPU; smaller batches on the smaller per-rank streams keep the copy pipeline deep)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.e2e_batch_vectors <= 0:
        args.e2e_batch_vectors = max(16, 256 // max(1, world))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dist = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import torch.distributed as dist_mod
        dist_mod.init_process_group("gloo", rank=rank, world_size=world)   # host-side plumbing only (id exchange, max)
        dist = dist_mod
    try:
        run_gpu(args, rank, world, local_rank, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
