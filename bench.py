#!/usr/bin/env python
"""Benchmark of the destripe hot path (BASELINE.json metric: destriped Mpixel/s of uint16 planes).

    python bench.py --gpus N --steps K --warmup W            # B200 engine (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port)

One "step" = one pass of the hot path over one synthetic 128 x 2048 x 2048 uint16 chunk
(BASELINE.json configs[1]; `--workload c3` = dual-config dispatch + dark/flat epilogue,
configs[2]).  Every rank owns its own chunk (weak scaling, Z-slab sharding, no collective).
`value` is measured with the chunk resident in HBM (CUDA events on the engine's compute
stream); `e2e` goes through the public API with pinned HOST buffers (H2D + D2H inside the
timed region).  The chunk (1.07 GB in, 1.07 GB out) is far larger than the 126 MB L2.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NO_CELLS = {"wavelet": "db3", "level": None, "sigma": 128, "max_threshold": 12}  # run_capsule.py:377-382
CELLS = {"wavelet": "db3", "level": None, "sigma": 64, "max_threshold": 3}  # run_capsule.py:383-388
HIGH_INT = 2500  # zarr_destriper.py:326
METRIC = "destriped Mpixel/s (uint16 planes)"
ALGO_BYTES_PER_PX = 4.0  # uint16 in + uint16 out


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--stream-planes", type=int, default=2000, help="planes of the streamed tile (c4 / c5)")
    ap.add_argument("--pool-planes", type=int, default=64, help="distinct planes recycled through the streamed tile (c4 / c5)")
    ap.add_argument("--planes", type=int, default=128)
    ap.add_argument("--height", type=int, default=2048)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--batch-planes", type=int, default=0, help="planes per kernel launch (0 = engine default)")
    ap.add_argument("--unique-planes", type=int, default=0, help="distinct synthetic planes in the chunk (0 = all planes distinct, seeds 0..Z-1)")
    ap.add_argument("--subchunk", type=int, default=0, help="planes per H2D/compute/D2H pipeline stage in the e2e leg (0 = engine default)")
    ap.add_argument("--cpu-planes", type=int, default=0, help="planes in the CPU-baseline sample (0 = 2 x cores)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--no-tma", action="store_true", help="level-1 analysis with the register-streaming kernel instead of the TMA ring")
    ap.add_argument("--no-overlap", action="store_true", help="issue every kernel on one stream in stage order")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.c3 (dual-config) resident timing")
    ap.add_argument("--umma", action="store_true", help="row filter on the tcgen05 kernel (opt-in: measured slower)")
    ap.add_argument("--row-filter", type=int, default=-1, help="0: FMA row filter, 1: mma.sync row filter with 8 rows per block, 2: with 4 rows (default: engine default = 1)")
    ap.add_argument("--e2e-sync", action="store_true", help="wait for every step's result before submitting the next")
    return ap.parse_args()


def workload_name(args):
    if args.workload in ("c4", "c5"):
        what = ("one SmartSPIM tile, Z-slab sharded over the ranks" if args.workload == "c4"
                else "one tile per rank (multi-tile channel)")
        return (f"{what}: {args.stream_planes}x1600x2000 uint16 streamed from host memory in 64-plane chunks through "
                "destripe_volume (dual-config dispatch + dark/flat epilogue), H2D / kernels / D2H overlapped")
    kind = (
        "log-space filter (no_cells config: db3, level None, sigma 128, max_threshold 12)"
        if args.workload == "c2"
        else "dual-config dispatch (cells sigma 64/thr 3, no_cells sigma 128/thr 12) + dark/flat epilogue"
    )
    return f"synthetic Zarr chunk {args.planes}x{args.height}x{args.width} uint16, {kind}"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(workload):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fp:
            return json.load(fp).get(workload)
    except Exception:
        return None


class ClockSampler:
    """SM clock / throttle-reason sampling during the timed region: an NVML polling thread
    (about 1 ms per sample; the recipe's ``nvidia-smi --query-gpu=clocks.sm,...`` line is the
    fallback when NVML cannot be loaded)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device):
        self.device = device
        self.samples = []  # (sm_mhz, reasons bitmask-as-names)
        self.sm_max = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None
        self._nv = None

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:
            import torch

            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.device).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(self.device))

    def _poll_nvml(self, nv, h):
        bits = [(nv.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                (nv.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (nv.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")]
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(mhz), [nm for bit, nm in bits if mask & bit]))
            except Exception:
                pass
            time.sleep(0.002)

    def _poll_smi(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                self.samples.append((float(parts[0]),
                                     [nm for nm, v in zip(self.NAMES, parts[2:6]) if v.lower().startswith("active")]))
                self.sm_max = float(parts[1])
            except Exception:
                time.sleep(0.01)

    def sample_now(self):
        """One synchronous sample from the calling thread (NVML only)."""
        if self._nv is None:
            return
        nv, h = self._nv
        try:
            mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            bits = [(nv.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                    (nv.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                    (nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                    (nv.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")]
            self.samples.append((float(mhz), [nm for bit, nm in bits if mask & bit]))
        except Exception:
            pass

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self._nv = (nv, h)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
        except Exception:
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._poll_smi, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["sampler not started"]}
        self.stop_flag.set()
        self.thread.join(timeout=6)
        sm = [s[0] for s in self.samples]
        reasons = sorted({r for s in self.samples for r in s[1]})
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": self.sm_max,
            "samples": len(sm),
            "source": self.source,
            "reasons": reasons,
        }


def cpu_reference_run(args, n_planes, cores, seed=1000):
    """Time the oracle port of filter_stripes / log_space_fft_filtering on the host cores with the
    reference's scheduling shape (N processes, whole blocks of planes each)."""
    from aind_smartspim_destripe_b200 import synthetic as S
    from oracle import worker as OW

    stack = S.synthetic_stack(n_planes, args.height, args.width, base_seed=seed,
                              cells_every=2 if args.workload == "c3" else 0, workers=cores)
    shadow = None
    cells = NO_CELLS
    if args.workload == "c3":
        flat, dark = S.synthetic_flat_dark(args.height, args.width)
        shadow = dict(retrospective=True, flatfield=flat, darkfield=dark, tile_config=None)
        cells = CELLS
    t0 = time.perf_counter()
    OW.run_planes_multiprocess(stack, NO_CELLS, cells, shadow, cores)
    dt = time.perf_counter() - t0
    return n_planes * args.height * args.width / dt / 1e6, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  pywt / skimage cannot be
    installed here (no network), so the timed code is the oracle port (kind = "port")."""
    from aind_smartspim_destripe_b200 import distributed as D
    from oracle import worker as OW

    rank, world, _ = D.env_rank()
    if rank != 0:
        return
    cores = OW.get_cpu_limit()
    # one step = the whole chunk of the workload (same config as the B200 arm); --cpu-planes bounds it
    n_planes = args.cpu_planes or args.planes
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(args, max(1, cores // 2), cores)
    vals, times = [], []
    for k in range(args.steps):
        v, dt = cpu_reference_run(args, n_planes, cores, seed=0)
        vals.append(v)
        times.append(dt)
    total_px = n_planes * args.height * args.width * args.steps
    value = total_px / sum(times) / 1e6
    sample = (f"{n_planes} planes of {args.height}x{args.width} per step on {cores} processes"
              + ("" if n_planes == args.planes else " (bounded sample of the workload)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "planes_timed_per_step": n_planes,
                   "note": "CPU path: numpy/scipy port of the reference (oracle/); pywt and skimage are installable neither "
                           "here nor on the GPU box.  pywt's C DWT is about 2-3x faster than the port's numpy DWT on "
                           "75-85 % of the time (SURVEY.md Appendix C): read the GPU/CPU ratio as an upper bound of "
                           "roughly 2x on what the real stack would give."},
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_b200(args):
    import torch

    from aind_smartspim_destripe_b200 import distributed as D
    from aind_smartspim_destripe_b200 import engine as E
    from aind_smartspim_destripe_b200 import synthetic as S

    rank, world, local = D.init()
    device = local if world > 1 else int(os.environ.get("DSTR_DEVICE", "0"))
    torch.cuda.set_device(device)
    numa = None if args.no_numa_bind else D.bind_to_gpu_numa(device)
    Z, H, W = args.planes, args.height, args.width
    px_per_step = Z * H * W

    # ---- synthetic chunk (seeded; ranks get different planes) ---------------------------------
    gen_workers = max(1, (os.cpu_count() or 8) // max(world, 1))
    stack = S.synthetic_stack(Z, H, W, base_seed=10_000 * rank, n_unique=args.unique_planes,
                              cells_every=2 if args.workload == "c3" else 0, workers=gen_workers)
    batch = args.batch_planes or min(Z, 128)
    eng = E.DestripeEngine(H, W, max_planes=batch, device=device)
    if args.no_overlap:
        eng.set_overlap(False)
    if args.no_tma:
        eng.set_tma(False)
    if args.umma:
        eng.set_umma(True)
    if args.row_filter >= 0:
        eng.set_row_filter(args.row_filter)
    pn, pc = E.make_params(NO_CELLS), None
    mode, flags = E.MODE_LOGSPACE, 0
    if args.workload == "c3":
        flat, dark = S.synthetic_flat_dark(H, W)
        eng.set_flat_dark(flat, dark.astype(np.float32))
        pc, mode, flags = E.make_params(CELLS), E.MODE_DISPATCH, E.FLAG_SHADOW

    d_in = E.DeviceBuffer(eng, stack.nbytes)
    d_out = E.DeviceBuffer(eng, stack.nbytes)
    d_in.upload(stack)
    stream = torch.cuda.ExternalStream(eng.compute_stream(), device=device)

    def step_resident():
        eng.filter_chunk_ptr(d_in.ptr, E.DSTR_U16, d_out.ptr, E.DSTR_U16, Z, pc, pn, HIGH_INT, mode,
                             flags | E.FLAG_NO_SYNC)

    # ---- HBM-resident timing (value) ------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    eng.synchronize()
    D.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    eng.reset_timers()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
    ev1.record(stream)
    if rank == 0:
        sampler.sample_now()  # the steps are enqueued and still running: at least one sample under load
    eng.synchronize()
    torch.cuda.synchronize()
    D.barrier()
    ms_local = ev0.elapsed_time(ev1)
    _, launches = eng.timers()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = D.max_over_ranks(ms_local)
    value = world * px_per_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- per-stage device times (CUDA events on the compute stream, same steps) -----------------
    eng.set_profiling(True)
    eng.reset_timers()
    for _ in range(args.steps):
        eng.filter_chunk_ptr(d_in.ptr, E.DSTR_U16, d_out.ptr, E.DSTR_U16, Z, pc, pn, HIGH_INT, mode, flags)
    stage_ms, _ = eng.timers()
    eng.set_profiling(False)
    stage_ms = {k: v / args.steps for k, v in stage_ms.items()}
    peak, peak_src = measured_peak()
    algo_bytes = ALGO_BYTES_PER_PX * px_per_step
    step_ms = ms_total / args.steps
    achieved = algo_bytes / (step_ms * 1e-3) / 1e9
    # per-kernel entries: each stage against ITS OWN algorithmic bytes (what it has to read and write once),
    # its device time (CUDA events on the compute stream, stage-ordered pass of the same steps) and, where an
    # ncu capture is committed under profiles/, the measured DRAM bytes per launch group
    lv = [(H, W)]
    for l in range(1, eng.max_level + 1):
        lv.append(E.level_shape(H, W, l))
    band = [float(h * w) for h, w in lv]  # coefficients per plane and level (level 0 = pixels)
    deep = sum(band[2:])
    own_bytes = {
        "analysis_l1": Z * (2.0 * band[0] + 8.0 * band[1]),                      # u16 in, cA1 + cH1 out
        "analysis_deep": Z * (4.0 * sum(band[1:-1]) + 8.0 * deep),               # cA_{l-1} in, cA_l + cH_l out
        "histogram": Z * 4.0 * sum(band[1:]),                                    # cH_l in
        "row_filter": Z * 8.0 * sum(band[1:]),                                   # cH_l in, dH_l out
        "row_filter_level1": Z * 8.0 * band[1],
        "synthesis_deep": Z * (8.0 * deep + 4.0 * sum(band[1:-1])),              # dA_l + dH_l in, dA_{l-1} out
        "final_synthesis_epilogue": Z * (8.0 * band[1] + 4.0 * band[0]),         # dA1 + dH1 + u16 in, u16 out
    }
    measured = recorded_traffic(args.workload)
    if not isinstance(measured, dict):
        measured = {}
    kernels = []
    for name, ob in own_bytes.items():
        ms = stage_ms.get(name, 0.0)
        if ms <= 0.0:
            continue
        gbs = ob / (ms * 1e-3) / 1e9
        kernels.append({"stage": name, "ms": ms, "share_of_stage_sum": None, "algorithmic_bytes": ob, "achieved_GBps": gbs,
                        "frac": gbs / peak, "dram_bytes_ncu": (measured.get("stages") or {}).get(name)})
    ssum = sum(k["ms"] for k in kernels if k["stage"] != "row_filter_level1")
    for k in kernels:
        k["share_of_stage_sum"] = k["ms"] / ssum if ssum > 0 else None
    dom = max((k for k in kernels if k["stage"] != "row_filter"), key=lambda k: k["ms"])["stage"] if kernels else None
    roofline = {
        "bound": "hbm", "kernel": "whole pipeline (all kernels of one chunk pass)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": measured.get("dram_bytes_per_step"), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes,
        "note": "headline = 4 B/px (u16 in + u16 out) x pixels of the chunk / device time of the whole step; kernels[] lists "
                "every stage against its own algorithmic bytes and the DRAM bytes ncu measured for it (profiles/traffic.json); "
                "analysis and final synthesis run at 0.6 of the HBM peak, the row filter is issue-slot bound (median "
                "selection, operand construction) with its contractions on the tensor pipe",
        "dominant_stage": dom,
        "row_filter_path": ("tcgen05 (kind::f16 hi/lo, TMEM accumulators)" if args.umma else
                            "fma (register-tiled FIR)" if args.row_filter == 0 else "mma.sync m16n8k16 fp16 hi/lo"),
        "kernels": kernels,
        "stage_ms_per_step": stage_ms,
    }

    # ---- extra: the dual-config dispatch + dark/flat epilogue (BASELINE configs[2]) on the same engine ----
    extra = None
    if args.workload == "c2" and not args.no_extra:
        stack3 = S.synthetic_stack(Z, H, W, base_seed=10_000 * rank + 5000, n_unique=args.unique_planes, cells_every=2,
                                   workers=gen_workers)
        flat, dark = S.synthetic_flat_dark(H, W)
        eng.set_flat_dark(flat, dark.astype(np.float32))
        d_in.upload(stack3)
        pc3 = E.make_params(CELLS)

        def step_c3(extra_flags=E.FLAG_NO_SYNC):
            eng.filter_chunk_ptr(d_in.ptr, E.DSTR_U16, d_out.ptr, E.DSTR_U16, Z, pc3, pn, HIGH_INT, E.MODE_DISPATCH,
                                 E.FLAG_SHADOW | extra_flags)

        for _ in range(3):
            step_c3()
        eng.synchronize()
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_c3()
        e1.record(stream)
        eng.synchronize()
        torch.cuda.synchronize()
        ms3 = D.max_over_ranks(e0.elapsed_time(e1))
        eng.set_profiling(True)
        eng.reset_timers()
        for _ in range(args.steps):
            step_c3(0)
        st3, _ = eng.timers()
        eng.set_profiling(False)
        v3 = world * px_per_step * args.steps / (ms3 * 1e-3) / 1e6
        extra = {"c3": {"workload": "dual-config dispatch (cells sigma 64/thr 3, no_cells sigma 128/thr 12) + dark/flat epilogue, "
                                    "every second plane with dense bright cells",
                        "value": v3, "unit": "Mpixel/s", "ms_per_step": ms3 / args.steps,
                        "whole_pipeline_frac": (algo_bytes / (ms3 / args.steps * 1e-3) / 1e9) / peak,
                        "stage_ms_per_step": {k: v / args.steps for k, v in st3.items()}}}
        # classic dual-band mode (pystripe filter_streaks) on the same planes: two notch-only
        # sub-band passes + clamp + blend; thresholds (Otsu of every plane) are derived once, outside the timed steps
        hist = eng.histogram_u16(stack3)
        from aind_smartspim_destripe_b200 import filtering as FL
        thr = np.array([FL.otsu_from_counts(hist[z]) for z in range(Z)], dtype=np.float32)
        del hist

        def step_db():
            eng.dual_band_chunk_ptr(d_in.ptr, E.DSTR_U16, d_out.ptr, Z, 256.0, 64.0, thr, level=-1, crossover=10.0)

        for _ in range(2):
            step_db()
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_db()
        e1.record(stream)
        eng.synchronize()
        torch.cuda.synchronize()
        msd = D.max_over_ranks(e0.elapsed_time(e1))
        extra["dual_band"] = {"workload": "classic dual-band (sigma fg 256 / bg 64, per-plane Otsu threshold, crossover 10): two "
                                          "notch-only sub-band passes + clamp + sigmoid blend, device-resident chunk",
                              "value": world * px_per_step * args.steps / (msd * 1e-3) / 1e6, "unit": "Mpixel/s",
                              "ms_per_step": msd / args.steps}
        d_in.upload(stack)
        del stack3

    # ---- end to end through the public API with pinned host buffers ----------------------------
    e2e = None
    if not args.no_e2e:
        pin_in = E.PinnedBuffer((Z, H, W), np.uint16)
        pin_out = E.PinnedBuffer((Z, H, W), np.uint16)
        pin_out2 = E.PinnedBuffer((Z, H, W), np.uint16)
        pin_in.array[...] = stack
        if args.subchunk:
            eng.set_subchunk(args.subchunk)
        for _ in range(2):
            eng.filter_chunk(pin_in.array, pn, cells=pc, out=pin_out.array, high_int=HIGH_INT, mode=mode, flags=flags)
        # Every step copies its chunk host -> device and its result device -> host through the public call.
        # Steps are submitted without waiting (DSTR_FLAG_NO_SYNC on pinned host buffers, results alternating
        # between two host buffers), so the upload of step k+1 overlaps the download of step k, as a caller
        # streaming chunks does; the clock stops after the last result is on the host.
        outs = [pin_out.array, pin_out2.array]
        D.barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            eng.filter_chunk(pin_in.array, pn, cells=pc, out=outs[k & 1], high_int=HIGH_INT, mode=mode,
                             flags=flags | (0 if args.e2e_sync else E.FLAG_NO_SYNC))
        eng.synchronize()
        dt = time.perf_counter() - t0
        dt = D.max_over_ranks(dt)
        if args.steps > 1 and not np.array_equal(outs[0][::17, ::64, ::64], outs[1][::17, ::64, ::64]):
            raise RuntimeError("e2e: the two result buffers differ")
        # result check-sum read back on the host (the step's result is consumed)
        checksum = int(pin_out.array[:: max(1, Z // 4), ::64, ::64].astype(np.int64).sum())
        e2e = {"value": world * px_per_step * args.steps / dt / 1e6, "unit": "Mpixel/s",
               "h2d_bytes_per_step": int(stack.nbytes), "d2h_bytes_per_step": int(stack.nbytes),
               "ms_per_step": 1e3 * dt / args.steps, "checksum": checksum,
               "submission": "synchronous" if args.e2e_sync else "asynchronous (DSTR_FLAG_NO_SYNC, 2 result buffers)"}
        # the ceiling of this figure: concurrent pinned H2D + D2H copies of the same buffers, nothing else
        if True:  # every rank probes at the same time: the ceiling under the same contention as the e2e leg
            try:
                D.barrier()
                t_in = torch.from_numpy(pin_in.array.reshape(-1).view(np.uint8))
                t_out = torch.from_numpy(pin_out.array.reshape(-1).view(np.uint8))
                g_in = torch.empty(t_in.numel(), dtype=torch.uint8, device=f"cuda:{device}")
                g_out = torch.empty_like(g_in)
                s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

                def both():
                    with torch.cuda.stream(s1):
                        g_in.copy_(t_in, non_blocking=True)
                    with torch.cuda.stream(s2):
                        t_out.copy_(g_out, non_blocking=True)

                for _ in range(2):
                    both()
                torch.cuda.synchronize()
                gbps = 0.0
                for _ in range(3):  # best of three bursts of four copy pairs
                    tc = time.perf_counter()
                    for _ in range(4):
                        both()
                    torch.cuda.synchronize()
                    gbps = max(gbps, 4 * t_in.numel() / (time.perf_counter() - tc) / 1e9)
                gmin = -D.max_over_ranks(-gbps)
                gsum = D.sum_over_ranks(gbps)
                e2e["pcie_bidirectional_GBps_each"] = gmin
                e2e["pcie_bidirectional_GBps_each_sum_over_ranks"] = gsum
                e2e["pcie_ceiling_Mpixel_per_s"] = gsum * 1e9 / 2 / 1e6
                e2e["frac_of_pcie_ceiling"] = e2e["value"] / e2e["pcie_ceiling_Mpixel_per_s"]
                e2e["pcie_probe"] = f"concurrent pinned H2D + D2H of the e2e buffers on all {world} rank(s) at the same time"
                del g_in, g_out
            except Exception as exc:  # the probe is informative only
                e2e["pcie_probe_error"] = str(exc)
        pin_in.free()
        pin_out.free()
        pin_out2.free()

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import worker as OW

        cores = OW.get_cpu_limit()
        n_planes = args.cpu_planes or 2 * cores
        v, dt = cpu_reference_run(args, n_planes, cores)
        cpu = {"value": v, "unit": "Mpixel/s", "cores": cores, "kind": "port",
               "sample": f"{n_planes} planes of {H}x{W} on {cores} processes, {dt:.1f} s wall "
                         "(oracle port; pywt/skimage not installable)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "planes_per_launch": batch,
                       "distinct_planes": Z if args.unique_planes <= 0 else min(Z, args.unique_planes),
                       "l2_policy": "inputs (1.07 GB/chunk) larger than L2; no flush needed",
                       "sharding": f"one chunk per rank, {world} rank(s), no collective",
                       "numa_node_rank0": numa},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "extra": extra, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit(line)
    d_in.free()
    d_out.free()
    eng.close()
    D.shutdown()


class CyclicVolume:
    """(Z, H, W) uint16 array-like that recycles a pool of distinct planes (a 2000-plane tile is 12.8 GB and
    eight of them exceed the host memory: SURVEY.md section 8d)."""

    def __init__(self, pool, Z):
        self.pool, self.shape, self.dtype = pool, (Z,) + pool.shape[1:], pool.dtype

    def __getitem__(self, key):
        a, b, _ = key.indices(self.shape[0])
        return self.pool[np.arange(a, b) % self.pool.shape[0]]

    def read_into(self, dst, a, b):  # destripe_volume's zero-copy source protocol
        P, z = self.pool.shape[0], a
        while z < b:  # contiguous runs of the pool: plain memcpy
            p0 = z % P
            n = min(b - z, P - p0)
            dst[z - a : z - a + n] = self.pool[p0 : p0 + n]
            z += n


class RecycledSink:
    """(Z, H, W) sink: every result lands in a host ring (the write stage of the pipeline), a check-sum is kept."""

    def __init__(self, shape, ring=128):
        self.shape, self.ring, self.checksum = shape, np.zeros((ring,) + shape[1:], np.uint16), 0

    def __setitem__(self, key, value):
        a, b, _ = key.indices(self.shape[0])
        for z0 in range(a, b, self.ring.shape[0]):
            z1 = min(b, z0 + self.ring.shape[0])
            self.ring[: z1 - z0] = value[z0 - a : z1 - a]
        self.checksum += int(value[0, ::64, ::64].astype(np.int64).sum())

    def write_from(self, src, a, b):  # destripe_volume's zero-copy sink protocol
        self[a:b] = src


def run_stream(args):
    """c4 / c5: the streamed-tile configurations (BASELINE.json configs[3], configs[4]) through the public
    chunk scheduler `zarr_destriper.destripe_volume` with host buffers on both sides."""
    import torch

    from aind_smartspim_destripe_b200 import distributed as D
    from aind_smartspim_destripe_b200 import synthetic as S
    from aind_smartspim_destripe_b200 import zarr_destriper as zd

    rank, world, local = D.init()
    device = local if world > 1 else int(os.environ.get("DSTR_DEVICE", "0"))
    torch.cuda.set_device(device)
    numa = None if args.no_numa_bind else D.bind_to_gpu_numa(device)
    H, W, Z = 1600, 2000, args.stream_planes
    pool = S.synthetic_stack(args.pool_planes, H, W, base_seed=20_000 + 1000 * rank, cells_every=2,
                             workers=max(1, (os.cpu_count() or 8) // max(world, 1)))
    flat, dark = S.synthetic_flat_dark(H, W)
    shadow = dict(retrospective=True, flatfield=flat, darkfield=dark, tile_config=None)
    if args.workload == "c4":
        z0, z1 = zd.z_slab(Z, rank, world, 64)  # strong scaling: the tile is split
    else:
        z0, z1 = 0, Z  # weak scaling: a whole tile per rank
    vol, sink = CyclicVolume(pool, Z), RecycledSink((Z, H, W))
    sampler = ClockSampler(device)
    steps_t, last = [], None
    for k in range(max(args.warmup, 1) + args.steps):
        timed = k >= max(args.warmup, 1)
        D.barrier()
        if timed and rank == 0 and not steps_t:
            sampler.start()
        t = zd.destripe_volume(vol, sink, NO_CELLS, CELLS, shadow, chunk_planes=64, z_range=(z0, z1), device=device,
                               microscope_high_int=HIGH_INT, io_threads=max(2, 16 // max(world, 1)))
        D.barrier()
        if timed:
            # whole call (engine contexts and pinned buffers are kept from the warm-up call: setup_s ~ 0)
            steps_t.append(D.max_over_ranks(t["wall_s"]))
            last = t
    clocks = sampler.stop() if rank == 0 else None
    px_step = (Z if args.workload == "c4" else world * Z) * H * W
    value = px_step * args.steps / sum(steps_t) / 1e6
    split = {k: D.max_over_ranks(float(last[k])) for k in ("read_s", "device_s", "write_s", "setup_s", "stream_s", "teardown_s", "wall_s")}
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = ALGO_BYTES_PER_PX * value * 1e6 / 1e9
        bytes_step = int(2 * px_step)
        emit({
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * sum(steps_t) / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "planes_per_chunk": 64, "distinct_planes": int(pool.shape[0]),
                       "planes_per_rank": int(z1 - z0), "l2_policy": "streamed chunks (410 MB each) larger than L2",
                       "sharding": ("Z-slabs aligned to the 64-plane output chunk" if args.workload == "c4" else "one tile per rank")
                                   + f", {world} rank(s), no collective", "numa_node_rank0": numa},
            "roofline": {"bound": "hbm", "kernel": "whole pipeline (streamed: bounded by the host <-> device copies)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * px_step},
            "cpu_baseline": None,
            "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": bytes_step, "d2h_bytes_per_step": bytes_step,
                    "note": "this workload IS the end-to-end path: host source -> pinned buffers -> GPU -> pinned buffers -> host sink"},
            "stage_seconds_last_step_max_over_ranks": split,
            "gpu_launches": None, "clocks": clocks,
        })
    D.shutdown()


def main():
    args = parse_args()
    # Exactly one JSON line may reach stdout: libraries (e.g. the NCCL version banner) write to
    # file descriptor 1 behind Python's back, so everything but the final line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("c4", "c5"):
        run_stream(args)
    else:
        run_b200(args)


_REAL_STDOUT = None


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


if __name__ == "__main__":
    main()
