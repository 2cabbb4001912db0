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
    import synth
    return synth.windowed_sinc_taps(TAPS, 1.0 / (2 * FACTOR))


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
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
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
            return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(name for bit, name in self.NVML_REASONS if mask & bit), "samples": len(rows), "source": "nvml, 1 ms poll"}
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
    line = {"impl": "reference", "metric": METRIC, "value": mean, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(), "taps": TAPS, "decimation": FACTOR, "buffer": BUF,
                       "note": "each step is a bounded 2 s sample of the stream on all host threads"},
            "cpu_baseline": dict(best, value=mean),
            "e2e": {"value": mean, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name():
    return (f"fastDecimatorC decimate-by-{FACTOR}, {TAPS} real taps on complex f32, 2^{LOG2_STREAM}-sample white-noise IQ stream as "
            f"{(1 << LOG2_STREAM) // BUF} x {BUF}-sample buffers")


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank, dist):
    import sdr_b200
    from sdr_b200 import _lib as L
    from sdr_b200 import multigpu

    n = 1 << args.log2n
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
    halo = "none"
    if world > 1:
        halo = "nccl send/recv per pass"
        if args.halo == "peer":
            dist.barrier()   # every chunk complete before a neighbour may read it
            try:
                comm.share_chunks(d_in.ptr)
                halo = "peer memory (CUDA IPC over NVLink), read in place by the boundary launch"
            except sdr_b200.SdrError as e:
                if rank == 0:
                    print(f"peer-memory halo unavailable ({e}); using NCCL", file=sys.stderr)

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
            ctx.sync()

    def step():
        multigpu.decimate_sharded(dec, comm, plan, d_in.ptr, d_out.ptr)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = sdr_b200.Event(ctx), sdr_b200.Event(ctx)
    launches0 = ctx.launches
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    ms = e0.elapsed_ms(e1)
    barrier()
    t_wall1 = time.time()
    launches = ctx.launches - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    kernel_name = dec.last_kernel()

    # max over ranks (device time)
    if world > 1:
        import torch
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        tl = torch.tensor([launches], dtype=torch.int64)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches_total = int(tl.item())
    else:
        ms_max, launches_total = ms, launches
    ms_per_step = ms_max / args.steps
    value = n / (ms_per_step * 1e-3) / 1e6

    # size-independent parity property at full size: checksum of all shards' outputs == rank-independent value
    csum = ctx.checksum32(d_out, 2 * plan.out_count, first_word=2 * plan.out_begin)
    if world > 1:
        import torch
        lo = torch.tensor([csum & 0xffffffff, csum >> 32], dtype=torch.int64)
        parts = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(parts, lo)
        csum = sum(int(p[0]) + (int(p[1]) << 32) for p in parts) & 0xffffffffffffffff

    # ---- end to end through the Pipes boundary with host buffers (this rank's share of the stream + its halo) ----
    e2e, e2e_spot = None, None
    if not args.no_e2e:
        r = run_e2e(args, ctx, dec, plan, rank, world, dist, sdr_b200, L)
        if r is not None:
            e2e, e2e_spot = r

    # ---- roofline of the dominant kernel (rank 0's launch: its chunk / its time) ----
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_PER_SAMPLE * plan.in_count / ((ms / args.steps) * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ((traffic or {}).get("dram_bytes_per_input_sample") or 0) * plan.in_count or None,
                "kernel": kernel_name, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * plan.in_count,
                "note": "duration = CUDA events on the library's stream over the timed region / steps (ring kernel + ragged-tail launch)"}
    if traffic:
        roofline["traffic_source"] = traffic.get("source")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_run(1, args.cpu_seconds)
        try:
            if e2e is not None:
                e2e["spot_parity_vs_reference_avx"] = e2e_spot_parity(e2e_spot)
        except Exception:
            pass

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": workload_name() if args.log2n == LOG2_STREAM else f"reduced 2^{args.log2n}-sample stream (not the headline size)",
                           "taps": TAPS, "decimation": FACTOR, "buffer": BUF, "samples": n,
                           "l2": "inputs exceed L2 (per-GPU chunk %.0f MiB in + %.0f MiB out vs 126 MB L2)" % (
                               8 * plan.in_count / 2 ** 20, 8 * plan.out_count / 2 ** 20),
                           "sharding": "single GPU" if world == 1 else f"{world} overlapping chunks, halo of {TAPS - FACTOR} samples per boundary via {halo}",
                           "arithmetic": "fp32 FMA, taps in increasing order; parity vs reference AVX path <= 1e-5 of output scale (tests/test_gpu_parity.py)",
                           "output_checksum": "%016x" % csum},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_total, "clocks": clocks}
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
           "bound": "PCIe host-to-device: 8 B per input sample must cross the link"}
    pipe.close()
    hin.free()
    hout.free()
    return res, spot


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
    ap.add_argument("--halo", default="nccl", choices=["nccl", "peer"],
                    help="multi-GPU halo transport: NCCL send/recv per pass (default) or in-place peer-memory reads")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-batch-vectors", type=int, default=256,
                    help="sdr_pipe_set_batch: output vectors per launch in the end-to-end run (256 x 8192 outputs = 128 MiB of input)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
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
