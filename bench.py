#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (voice-samples/s at 48 kHz offline render; achieved HBM
GB/s vs peak) on BASELINE configs[1] (cfg2: saw Oscillator -> Moog Filter -> ADSR -> VCA,
4096 detuned voices, 48 kHz x 1 s) per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one 1-second render (48000 samples) of every voice.  `value`: outputs (stems + mix)
stay in HBM, per-voice parameters already resident.  `e2e`: the same render through the C ABI
with HOST buffers -- per-voice parameters re-sent from host memory and the stereo mix read back
every step (plus an `e2e_stems` line where all per-voice stems cross PCIe too).  Multi-GPU: weak
scaling, 4096 voices per rank, contiguous voice ranges, one NCCL sum of the mix per step.
`--impl reference` times the reference's CPU algorithm (the C++ oracle: the Rust reference cannot
be built in this image) on the box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 48000
BYTES_PER_VOICE_SAMPLE = {"cfg1": 8, "cfg2": 8, "cfg3": 8, "cfg3b": 16, "cfg4": 8}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(config):
    """dram bytes per launch of the voice kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(config)
    except Exception:
        return None


def ncu_utilization(config):
    """issue / pipe utilisation of the same capture (SURVEY.md §8d asks for it next to the HBM fraction)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(config + "_utilization")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons while the timed region runs: NVML (nvidia_ml_py) every 20 ms,
    `nvidia-smi --query-gpu` every 200 ms as the fallback (the B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index, uuid=None):
        super().__init__(daemon=True)
        self.gpu, self.uuid, self.rows, self.stop_flag = gpu_index, uuid, [], threading.Event()
        self.sm_max, self.source = None, "nvidia-smi"

    def _nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = None
        if self.uuid:
            for u in (self.uuid, "GPU-" + self.uuid):
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
                    break
                except Exception:
                    h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = [(nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else nv.nvmlClocksThrottleReasonHwSlowdown),
                (nv.nvmlClocksEventReasonHwThermalSlowdown if hasattr(nv, "nvmlClocksEventReasonHwThermalSlowdown") else nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                (nv.nvmlClocksEventReasonSwThermalSlowdown if hasattr(nv, "nvmlClocksEventReasonSwThermalSlowdown") else nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                (nv.nvmlClocksEventReasonSwPowerCap if hasattr(nv, "nvmlClocksEventReasonSwPowerCap") else nv.nvmlClocksThrottleReasonSwPowerCap)]
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        get(h)  # raises here if unsupported -> fallback
        self.source = "nvml"
        while not self.stop_flag.is_set():
            mask = get(h)
            self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(self.sm_max)] +
                             ["Active" if mask & b else "Not Active" for b in bits])
            self.stop_flag.wait(0.02)

    def _smi(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def run(self):
        try:
            self._nvml()
        except Exception:
            self._smi()

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        reasons = sorted({n for r in self.rows for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


def cpu_reference_run(config, steps, warmup, sample_voices=None):
    """The reference's CPU algorithm (oracle port) on all host cores, bounded sample of the workload."""
    import srack_b200 as srk  # patch descriptions only
    from oracle import orc

    cores = os.cpu_count() or 1
    builders = srk.patches.CFG5_GRAPHS if config == "cfg5" else [srk.patches.CONFIGS[config][0]]
    V = sample_voices or (64 * cores) // len(builders)  # cfg5: the sample is spread over its 8 graphs
    banks = []
    for b in builders:
        p = orc.OraclePatch(48000, 1024, 2)
        b(p, V)
        p.render(V, 0, stems=False, mix=False, n_threads=cores)  # build the voice bank outside the timed region
        banks.append(p)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in banks:
            p.render(V, N_SAMPLES, stems=False, mix=True, n_threads=cores)  # state carries over, like execute()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    V = V * len(builders)
    return dict(value=V * N_SAMPLES / per_step, ms_per_step=per_step * 1e3, cores=cores, voices=V,
                sample=f"{V} voices x {N_SAMPLES} samples per step ({V} of the workload's voices; "
                       f"block-based execute(), buffer_size 1024, one patch instance per voice, {cores} threads)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(BYTES_PER_VOICE_SAMPLE) + ["cfg5"])
    ap.add_argument("--voices-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import srack_b200 as srk

    desc = srk.patches.CONFIGS[args.config][2] if args.config in srk.patches.CONFIGS else "mixed batch: 8 distinct patch graphs x 32768 voices each"
    metric = "voice-samples/sec @48 kHz offline render"
    unit = "voice-samples/s"

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(args.config, args.steps, args.warmup)
        line = {
            "impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (+f64 oscillator phase)", "data": "synthetic",
            "config": {"workload": f"{args.config}: {desc}", "sample_rate": 48000, "n_samples": N_SAMPLES,
                       "buffer_size": 1024, "note": "CPU oracle = C++ restatement of the Rust reference "
                       "(rustc/cargo absent), timed on host cores"},
            "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: srack_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries the one JSON line: NCCL's banner / debug log (whatever level the box sets) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    if args.config == "cfg5":
        return run_cfg5(args, srk, torch, dist, rank, local_rank, world, dev, metric, unit)
    V_gpu = args.voices_per_gpu or (min(srk.patches.CONFIGS[args.config][1], 65536) if world == 1
                                    else {"cfg4": 32768}.get(args.config, min(srk.patches.CONFIGS[args.config][1], 65536)))
    V_total = V_gpu * world
    off, cnt = srk.shard.voice_range(V_total, rank, world)
    C = 2

    patch = srk.Patch(device=local_rank)
    srk.patches.CONFIGS[args.config][0](patch, V_total)
    patch.plan()
    info = patch.program_info(cnt)
    stems = torch.empty((C, N_SAMPLES, cnt), dtype=torch.float32, device=dev)
    mix = torch.empty((C, N_SAMPLES), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_resident():
        patch.render_into(cnt, N_SAMPLES, off, stems.data_ptr(), mix.data_ptr(), device_out=True, async_=True,
                          stream=stream)
        if world > 1:
            srk.shard.reduce_mix(mix)  # one NCCL sum of the [2][48000] mix

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, collect_kernel_ms=False):
        for _ in range(warmup):
            fn()
        barrier()
        launches0 = patch.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ms = []
        e0.record()
        for _ in range(steps):
            fn()
            if collect_kernel_ms:
                kernel_ms.append(None)  # placeholder; kernel events are read after the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, patch.launch_count() - launches0

    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step, launches = timed(step_resident, args.steps, args.warmup)
    # voice-kernel device time (CUDA events on the render stream, inside the library), averaged live
    k_ms = []
    for _ in range(max(3, min(args.steps, 10))):
        step_resident()
        torch.cuda.synchronize()
        k_ms.append(patch.last_render_ms()[0])
    kernel_ms = sum(k_ms) / len(k_ms)
    value = V_total * N_SAMPLES / (ms_step * 1e-3)

    # ---- e2e: host buffers through the C ABI (per-voice params H2D + mix D2H every step)
    pv_params = list(patch.per_voice.items())  # ((module, param id), host array) set by the patch description
    mix_host = torch.empty((C, N_SAMPLES), dtype=torch.float32).pin_memory()
    h2d_bytes = info["param_words"] * cnt * 4
    d2h_bytes = C * N_SAMPLES * 4

    def step_e2e():
        for (m, pid), arr in pv_params:
            m.set_param_per_voice(pid, arr)  # marks the table dirty -> rebuilt in pinned staging + H2D
        patch.render_into(cnt, N_SAMPLES, off, None, mix_host.data_ptr(), device_out=False)
        return float(mix_host[0, -1])

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = V_total * N_SAMPLES / e2e_s
    clocks = sampler.summary() if sampler else None  # covers the device-timed and the e2e-timed regions

    # ---- e2e with all stems to pinned host memory as well (PCIe bound), N = 1 only, bounded memory
    e2e_stems = None
    if world == 1 and cnt * N_SAMPLES * C * 4 <= 4 << 30:
        stems_host = torch.empty((C, N_SAMPLES, cnt), dtype=torch.float32).pin_memory()

        def step_e2e_stems():
            for (m, pid), arr in pv_params:
                m.set_param_per_voice(pid, arr)
            patch.render_into(cnt, N_SAMPLES, off, stems_host.data_ptr(), mix_host.data_ptr(), device_out=False)

        step_e2e_stems()
        t0 = time.perf_counter()
        n_rep = max(2, min(args.steps, 5))
        for _ in range(n_rep):
            step_e2e_stems()
        dt = (time.perf_counter() - t0) / n_rep
        e2e_stems = {"value": V_total * N_SAMPLES / dt, "unit": unit, "h2d_bytes_per_step": h2d_bytes,
                     "d2h_bytes_per_step": d2h_bytes + C * N_SAMPLES * cnt * 4}
        del stems_host

    # ---- the same patch at the voice count where the chip is full (one-warp schedule): a second, throughput-shaped
    #      data point next to the headline (which BASELINE.json quotes at 4096 voices, a latency-shaped launch)
    full_chip = None
    if world == 1 and args.voices_per_gpu == 0 and args.config == "cfg2":
        Vf = 65536
        del stems
        torch.cuda.empty_cache()
        pf = srk.Patch(device=local_rank)
        srk.patches.CONFIGS[args.config][0](pf, Vf)
        pf.plan()
        info_f = pf.program_info(Vf)
        stems_f = torch.empty((C, N_SAMPLES, Vf), dtype=torch.float32, device=dev)
        kf = []
        for i in range(2 + 4):
            pf.render_into(Vf, N_SAMPLES, 0, stems_f.data_ptr(), mix.data_ptr(), device_out=True, stream=stream)
            torch.cuda.synchronize()
            if i >= 2:
                kf.append(pf.last_render_ms())
        k_f = sum(k for k, _ in kf) / len(kf)
        t_f = sum(t for _, t in kf) / len(kf)
        peak_f, _ = measured_peak_hbm()
        full_chip = {"workload": f"{args.config} @ {Vf} voices x {N_SAMPLES} samples, stems + mix in HBM", "value": Vf * N_SAMPLES / (t_f * 1e-3),
                     "unit": unit, "ms_per_step": t_f, "kernel_ms": k_f, "steps": len(kf), "warmup": 2,
                     "roofline": {"bound": "hbm", "achieved": Vf * N_SAMPLES * BYTES_PER_VOICE_SAMPLE[args.config] / (k_f * 1e-3) / 1e9,
                                  "peak": peak_f, "unit": "GB/s",
                                  "frac": Vf * N_SAMPLES * BYTES_PER_VOICE_SAMPLE[args.config] / (k_f * 1e-3) / 1e9 / peak_f},
                     "block_threads": info_f["block_threads"], "step_samples": info_f["step_samples"],
                     "voice_groups_per_block": info_f["groups_per_block"], "smem_bytes": info_f["smem_bytes"],
                     "gpu_launches": int(pf.launch_count()),
                     "timing": "CUDA events on the render stream inside the library (whole call and voice kernel), 25 GB of stems per step (> L2)"}
        del stems_f, pf

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        bpvs = BYTES_PER_VOICE_SAMPLE[args.config]
        algo_bytes = cnt * N_SAMPLES * bpvs  # per launch (one rank's voices)
        achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(args.config, steps=3, warmup=1)
            cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (+f64 oscillator phase)", "data": "synthetic",
            "config": {"workload": f"{args.config}: {desc}", "voices_per_gpu": V_gpu, "voices_total": V_total,
                       "n_samples": N_SAMPLES, "sample_rate": 48000, "parallelism": f"voice-shard x{world}",
                       "outputs": "stems [2][48000][V] + mix [2][48000] in HBM",
                       "l2": f"each step writes {C * N_SAMPLES * cnt * 4 / 1e6:.0f} MB of stems (> 126 MB L2), no flush needed",
                       "block_threads": info["block_threads"], "step_samples": info["step_samples"],
                       "smem_bytes": info["smem_bytes"], "warps_per_voice_group": info["n_warps"],
                       "pipeline_stages": info["n_stages"], "voice_groups_per_block": info["groups_per_block"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.config), "peak_source": peak_src,
                         "algorithmic_bytes_per_voice_sample": bpvs, "kernel": "render_voices_kernel",
                         "kernel_ms": kernel_ms, "ncu_utilization": ncu_utilization(args.config),
                         "note": "latency/issue-bound DSP recurrences: HBM fraction is small by construction (SURVEY.md §8d)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_s * 1e3, "what": "per-voice params from host + render + mix to pinned host"},
            "e2e_stems": e2e_stems,
            "full_chip": full_chip,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_cfg5(args, srk, torch, dist, rank, local_rank, world, dev, metric, unit):
    """BASELINE configs[4]: 8 distinct patch graphs x 32768 voices each, one graph per GPU at N = 8; with fewer
    ranks every rank renders its share of the graphs one after the other (the job stays the same: strong
    scaling; shares balanced by measured kernel time).  One NCCL sum of the mix per step.  Stems stay in HBM (one graph's 12.6 GB buffer, reused)."""
    graphs = srk.patches.CFG5_GRAPHS
    V = args.voices_per_gpu or srk.patches.CFG5_VOICES
    # graphs -> ranks: longest-processing-time first, with the kernel times measured at 32768 voices
    # (profiles/r03s_bench_cfg5_1gpu.json) as weights; every rank computes the same assignment
    weight = {"cfg2": 20.3, "cfg3": 29.2, "cfg3b": 41.7, "cfg4": 57.4, "cfg5_bandpass": 20.4, "cfg5_two_osc": 31.5,
              "cfg5_no_noise": 54.8, "cfg5_gated_sine": 19.6}
    load, owner = [0.0] * world, {}
    for g in sorted(range(len(graphs)), key=lambda i: (-weight.get(graphs[i].__name__, 30.0), i)):
        r = min(range(world), key=lambda k: (load[k], k))
        owner[g] = r
        load[r] += weight.get(graphs[g].__name__, 30.0)
    mine = [g for g in range(len(graphs)) if owner[g] == rank]
    C = 2
    patches = []
    for g in mine:
        p = srk.Patch(device=local_rank)
        graphs[g](p, V)
        p.plan()
        patches.append((g, p, p.program_info(V)))
    stems = torch.empty((C, N_SAMPLES, V), dtype=torch.float32, device=dev)
    mixes = [torch.empty((C, N_SAMPLES), dtype=torch.float32, device=dev) for _ in mine]
    total = torch.zeros((C, N_SAMPLES), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        for (g, p, _), mx in zip(patches, mixes):
            p.render_into(V, N_SAMPLES, 0, stems.data_ptr(), mx.data_ptr(), device_out=True, async_=True, stream=stream)
        torch.sum(torch.stack(mixes), dim=0, out=total)
        if world > 1:
            srk.shard.reduce_mix(total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = sum(p.launch_count() for _, p, _ in patches)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = sum(p.launch_count() for _, p, _ in patches) - launches0
    per_graph = {}
    for g, p, info in patches:  # kernel time per graph (CUDA events inside the library)
        p.render_into(V, N_SAMPLES, 0, stems.data_ptr(), mixes[0].data_ptr(), device_out=True, stream=stream)
        torch.cuda.synchronize()
        per_graph[graphs[g].__name__] = {"kernel_ms": p.last_render_ms()[0], "warps_per_voice_group": info["n_warps"],
                                         "step_samples": info["step_samples"], "voice_groups_per_block": info["groups_per_block"]}
    # e2e: host per-voice parameters in, mix out to pinned host memory, every step
    mix_host = torch.empty((C, N_SAMPLES), dtype=torch.float32).pin_memory()
    pv = [(p, list(p.per_voice.items())) for _, p, _ in patches]
    h2d = sum(info["param_words"] for _, _, info in patches) * V * 4

    def step_e2e():
        for (p, items), mx in zip(pv, mixes):
            for (m, pid), arr in items:
                m.set_param_per_voice(pid, arr)
            p.render_into(V, N_SAMPLES, 0, None, mx.data_ptr(), device_out=True, async_=True, stream=stream)
        torch.sum(torch.stack(mixes), dim=0, out=total)
        if world > 1:
            srk.shard.reduce_mix(total)
        mix_host.copy_(total, non_blocking=False)
        return float(mix_host[0, -1])

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(2, min(args.steps, 5))
    for _ in range(n_e2e):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0].item()), float(t[1].item()) * 1e-3
        gathered = [None] * world
        dist.all_gather_object(gathered, per_graph)
        per_graph = {k: v for d in gathered for k, v in d.items()}
    if rank == 0:
        n_vs = len(graphs) * V * N_SAMPLES
        peak, peak_src = measured_peak_hbm()
        k_sum = sum(v["kernel_ms"] for v in per_graph.values())
        algo = sum((16 if name == "cfg3b" else 8) * V * N_SAMPLES for name in per_graph)
        print(json.dumps({
            "metric": metric, "value": n_vs / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (+f64 oscillator phase)", "data": "synthetic",
            "config": {"workload": "cfg5: mixed batch, 8 distinct patch graphs x %d voices each, 48 kHz x 1 s" % V,
                       "graphs": [g.__name__ for g in graphs], "graphs_per_gpu": len(graphs) / world, "voices_total": len(graphs) * V,
                       "n_samples": N_SAMPLES, "parallelism": f"graph-shard x{world}",
                       "l2": "every graph writes 12583 MB of stems per step (> 126 MB L2), no flush needed"},
            "roofline": {"bound": "hbm", "achieved": algo / (k_sum * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / (k_sum * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "render_voices_kernel (sum over the 8 graphs' launches)", "per_graph": per_graph},
            "cpu_baseline": None,
            "e2e": {"value": n_vs / e2e_s, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": C * N_SAMPLES * 4,
                    "ms_per_step": e2e_s * 1e3, "what": "per-voice params from host + render + mix to pinned host"},
            "gpu_launches": int(launches)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
