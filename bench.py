#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (voice-samples/s at 48 kHz offline render; achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline (the top-level keys of the one JSON line): BASELINE configs[1] -- cfg2, saw Oscillator -> Moog Filter ->
ADSR -> VCA, 4096 detuned voices per GPU, 48 kHz x 1 s; weak scaling over GPUs (contiguous voice ranges, one NCCL sum
of the stereo mix per step).  A step = one 1-second render (48000 samples) of every voice.
  value : outputs (stems + mix) stay in HBM, per-voice parameters already resident; CUDA events, max over ranks.
  e2e   : the same render through the C ABI with HOST buffers -- per-voice parameters re-sent from host memory, the
          stereo MIX (not the stems) read back to pinned host memory every step; at N > 1 the NCCL sum is inside.
Every other BASELINE config rides along in the same line as a sub-record with its own ms_per_step / kernel_ms /
roofline, so the driver's records carry them:
  N = 1 : `configs` = cfg2 @ 65536 voices (the full chip), cfg3 and cfg3b @ 65536 (configs[2]), cfg4 @ 32768 (one GPU's
          share of configs[3]); `cfg5` (configs[4], all eight graphs on the one GPU: the N = 1 point of its strong-scaling
          curve); `block_cadence` (srk_render per 1024-sample block); the CPU baselines.
  N > 1 : `cfg4_shard` (configs[3]: 32768 voices per GPU, 262144 at N = 8), `cfg5` (configs[4]: 8 graphs x 32768 voices
          dealt to the ranks, one graph per GPU at N = 8) and `cfg5_balanced` (the same job dealt as (graph, voice range)
          pieces: graphs longer than a rank's fair share are cut in two).
`mix_check` verifies the result in the same run: at N = 1 the mix against the f64 sum of the stems; at N > 1 the
NCCL-reduced mix against the f64 sum of the gathered per-rank mixes, and rank 0 re-renders another rank's voice range
(voice_offset) and compares it with what that rank produced, bit for bit.
`--impl reference` times the reference's CPU algorithm (oracle/: the C++ restatement -- the Rust reference cannot be
built in this image) on the box's host cores; that arm never imports the product package.
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
C = 2
BYTES_PER_VOICE_SAMPLE = {"cfg1": 8, "cfg2": 8, "cfg3": 8, "cfg3b": 16, "cfg4": 8}
METRIC = "voice-samples/sec @48 kHz offline render"
UNIT = "voice-samples/s"
DTYPE = "f32 (+f64 oscillator phase)"
# dependent-instruction chain of one cfg2 voice-sample: the ladder filter's 20 dependent f32 operations, ~4.4 cycles each
# (88 cycles/sample): no schedule can render 48000 samples of one voice faster than this
LADDER_CHAIN_CYCLES = 88.0
N_SM, SCHEDULERS_PER_SM = 148, 4


def bytes_per_voice_sample(name):
    return 16 if name == "cfg3b" else 8


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def sm_max_mhz():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["sm_max_mhz"])
    except Exception:
        return 1965.0


def ncu_capture(kernel_id):
    """The committed ncu capture of exactly this kernel image (profiles/traffic.json is keyed by srk_kernel_id: the hash
    of the kernel's source), or None -- a capture of another kernel is never quoted."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("kernels", {}).get(kernel_id)
    except Exception:
        return None


def roofline(name, V, kernel_ms, kernel_id, sm_mhz=None):
    """HBM roofline of one voice-kernel launch over V voices (algorithmic bytes) + the issue roofline
    when this kernel image has a committed ncu capture."""
    peak, peak_src = measured_peak_hbm()
    bpvs = bytes_per_voice_sample(name)
    achieved = V * N_SAMPLES * bpvs / (kernel_ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
         "peak_source": peak_src, "algorithmic_bytes_per_voice_sample": bpvs, "kernel": kernel_id, "kernel_ms": kernel_ms}
    cap = ncu_capture(kernel_id)
    if cap and cap.get("voices") == V:
        r["traffic"] = cap.get("dram_bytes_per_launch")
        r["ncu"] = {k: cap[k] for k in ("issue_active_pct", "warps_active_pct", "fp64_pipe_pct", "fma_pipe_pct", "alu_pipe_pct",
                                       "dram_throughput_pct", "source", "commit") if k in cap}
        wi = cap.get("warp_instructions_per_launch")
        if wi:
            clock = (sm_mhz or sm_max_mhz()) * 1e6
            r["roofline_issue"] = {"warp_instructions_per_32_voice_samples": wi / (V * N_SAMPLES / 32.0),
                                   "issue_slots_per_s": N_SM * SCHEDULERS_PER_SM * clock, "sm_mhz": clock / 1e6,
                                   "frac": wi / (kernel_ms * 1e-3) / (N_SM * SCHEDULERS_PER_SM * clock),
                                   "what": "warp instructions of this launch (ncu smsp__inst_executed.sum) / live kernel time / "
                                           "(148 SMs x 4 schedulers x SM clock)"}
    else:
        r["traffic_note"] = "no committed ncu capture of this kernel image at this voice count (profiles/traffic.json)"
    return r


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


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the oracle (test infrastructure) as the CPU implementation; never touches the product package
# ---------------------------------------------------------------------------------------------------------------------
def patch_descriptions():
    """s-rack_b200/patches.py as a plain module (pure Python: graph descriptions + per-voice parameter generators),
    WITHOUT importing the package -- importing srack_b200 would dlopen libsrack_b200.so into the reference arm."""
    import importlib.util
    d = os.path.join(ROOT, "s-rack_b200")
    if d not in sys.path:
        sys.path.insert(0, d)  # patches.py falls back to `import constants`
    spec = importlib.util.spec_from_file_location("srk_patch_descriptions", os.path.join(d, "patches.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_run(config, steps, warmup, threads=None, sample_voices=None):
    """The reference's CPU algorithm (oracle port) on `threads` host threads (default: all), bounded sample."""
    from oracle import orc
    patches = patch_descriptions()
    assert "srack_b200" not in sys.modules or threads is not None, "the reference arm must not load the product"
    cores = os.cpu_count() or 1
    threads = threads or cores
    builders = patches.CFG5_GRAPHS if config == "cfg5" else [patches.CONFIGS[config][0]]
    V = sample_voices or max((64 * threads) // len(builders), 8)  # cfg5: the sample is spread over its 8 graphs
    banks = []
    for b in builders:
        p = orc.OraclePatch(48000, 1024, 2)
        b(p, V)
        p.render(V, 0, stems=False, mix=False, n_threads=threads)  # build the voice bank outside the timed region
        banks.append(p)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in banks:
            p.render(V, N_SAMPLES, stems=False, mix=True, n_threads=threads)  # state carries over, like execute()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    V = V * len(builders)
    return dict(value=V * N_SAMPLES / per_step, ms_per_step=per_step * 1e3, cores=threads, host_cores=cores, voices=V,
                sample=f"{V} voices x {N_SAMPLES} samples per step ({V} of the workload's voices; block-based execute(), "
                       f"buffer_size 1024, one patch instance per voice, {threads} thread{'s' if threads > 1 else ''}; mix only)")


def reference_arm(args):
    patches = patch_descriptions()
    desc = patches.CONFIGS[args.config][2] if args.config in patches.CONFIGS else "mixed batch: 8 distinct patch graphs x 32768 voices each"
    r = cpu_reference_run(args.config, args.steps, args.warmup)
    one = cpu_reference_run(args.config, steps=1, warmup=1, threads=1, sample_voices=32)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": f"{args.config}: {desc}", "sample_rate": 48000, "n_samples": N_SAMPLES, "buffer_size": 1024,
                   "note": "CPU oracle = C++ restatement of the Rust reference (rustc/cargo absent), timed on host cores; "
                           "throughput is per voice-sample, the sample is a bounded number of the workload's voices"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "cpu_baseline_1thread": {"value": one["value"], "unit": UNIT, "cores": 1, "kind": "port", "sample": one["sample"],
                                 "note": "the reference's own execution model: one audio thread (main.rs:59-63)"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "product_library_loaded": "srack_b200" in sys.modules or "libsrack_b200" in open("/proc/self/maps").read(),
    }))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide handles of the GPU arm."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        import srack_b200 as srk
        self.np, self.torch, self.dist, self.srk, self.args = np, torch, dist, srk, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl ours needs a CUDA device: srack_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # stdout carries the one JSON line: NCCL's banner / debug log (whatever level the box sets) goes to stderr
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.builders = {n: c[0] for n, c in srk.patches.CONFIGS.items()}
        for g in srk.patches.CFG5_GRAPHS:
            self.builders[g.__name__] = g
        self._stems = None

    def stems(self, V):
        """One stems buffer, grown to the largest request and reused (views): f32 [C][N][V]."""
        need = C * N_SAMPLES * V
        if self._stems is None or self._stems.numel() < need:
            self._stems = None
            self.torch.cuda.empty_cache()
            self._stems = self.torch.empty(need, dtype=self.torch.float32, device=self.dev)
        return self._stems[:need].view(C, N_SAMPLES, V)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def patch(self, name, V_total):
        p = self.srk.Patch(device=self.local_rank)
        self.builders[name](p, V_total)
        p.plan()
        return p


def timed_steps(ctx, fn, steps, warmup):
    """W warm-up steps, then exactly K steps between CUDA events on the launching stream, barrier + synchronize on
    both sides, max over ranks -> ms per step."""
    torch = ctx.torch
    for _ in range(warmup):
        fn()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    ctx.barrier()
    return ctx.max_over_ranks(e0.elapsed_time(e1) / steps)[0]


def kernel_ms_of(ctx, patches, render_one, reps=3):
    """Voice-kernel device time (CUDA events on the render stream, inside the library), averaged live."""
    out = []
    for p in patches:
        ks = []
        for _ in range(reps):
            render_one(p)
            ctx.torch.cuda.synchronize()
            ks.append(p.last_render_ms()[0])
        out.append(sum(ks) / len(ks))
    return out


def launch_shape(info):
    return {"block_threads": info["block_threads"], "step_samples": info["step_samples"], "smem_bytes": info["smem_bytes"],
            "warps_per_voice_group": info["n_warps"], "pipeline_stages": info["n_stages"],
            "voice_groups_per_block": info["groups_per_block"], "fused": bool(info["fused"]),
            "registers": info["fused_regs"] or None}


def check_mix_against_stems(ctx, mix, stems, V):
    """N = 1: |mix - sum_v stems| <= 1e-5 * max(|sum|, sqrt(V)) with the sum in f64."""
    torch = ctx.torch
    ref = torch.zeros((C, N_SAMPLES), dtype=torch.float64, device=ctx.dev)
    for c in range(C):  # channel by channel: the f64 copy of one channel of 65536 voices is 25 GB otherwise
        for n0 in range(0, N_SAMPLES, 6000):
            ref[c, n0:n0 + 6000] = stems[c, n0:n0 + 6000].sum(dim=1, dtype=torch.float64)
    err = (mix.double() - ref).abs()
    bound = 1e-5 * torch.maximum(ref.abs(), torch.tensor(float(V) ** 0.5, dtype=torch.float64, device=ctx.dev))
    ok = bool((err <= bound).all()) and bool(torch.isfinite(mix).all()) and float(mix.abs().max()) > 0.0
    return {"result": "ok" if ok else "FAILED", "what": "mix vs f64 sum of the stems over voices", "max_err": float(err.max()),
            "max_err_over_bound": float((err / bound).max()), "voices": V}


def check_reduced_mix(ctx, reduced, partial, V_total, rerender=None, what=""):
    """N > 1.  (a) the NCCL-reduced mix on rank 0 against the f64 sum of the per-rank mixes gathered to rank 0:
    |reduced - sum| <= 1e-5 * max(|sum|, sqrt(V)); (b) rank 0 renders the work of the LAST rank again on its own GPU
    (`rerender()` -> mix tensor) and compares it with what that rank produced, bit for bit."""
    torch, dist = ctx.torch, ctx.dist
    parts = [torch.empty_like(partial) for _ in range(ctx.world)] if ctx.rank == 0 else None
    dist.gather(partial, parts, dst=0)
    if ctx.rank != 0:
        return None
    ref = torch.stack(parts).double().sum(dim=0)
    err = (reduced.double() - ref).abs()
    bound = 1e-5 * torch.maximum(ref.abs(), torch.tensor(float(V_total) ** 0.5, dtype=torch.float64, device=ctx.dev))
    ok = bool((err <= bound).all()) and bool(torch.isfinite(reduced).all()) and float(reduced.abs().max()) > 0.0
    out = {"what": what or "NCCL-reduced mix vs f64 sum of the gathered per-rank mixes", "max_err": float(err.max()),
           "max_err_over_bound": float((err / bound).max()), "voices": V_total, "ranks": ctx.world}
    if rerender is not None:
        again = rerender()
        ctx.torch.cuda.synchronize()
        same = bool(torch.equal(again.view(torch.int32), parts[-1].view(torch.int32)))
        out["last_rank_rerendered_on_rank0_bit_identical"] = same
        ok = ok and same
    out["result"] = "ok" if ok else "FAILED"
    return out


def voice_shard(ctx, name, V_gpu, steps, warmup, e2e=True, sampler=None):
    """One patch graph, V_gpu voices per rank (contiguous voice ranges), one NCCL sum of the mix per step."""
    torch, srk = ctx.torch, ctx.srk
    world, rank = ctx.world, ctx.rank
    V_total = V_gpu * world
    off, cnt = srk.shard.voice_range(V_total, rank, world)
    patch = ctx.patch(name, V_total)
    stems = ctx.stems(cnt)
    mix = torch.empty((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev)

    def render(p=patch, with_stems=True):
        p.render_into(cnt, N_SAMPLES, off, stems.data_ptr() if with_stems else None, mix.data_ptr(), device_out=True,
                      async_=True, stream=ctx.stream)

    def step_resident():
        render()
        if world > 1:
            srk.shard.reduce_mix(mix)  # one NCCL sum of the [2][48000] mix

    launches0 = patch.launch_count()
    ms_step = timed_steps(ctx, step_resident, steps, warmup)
    launches = (patch.launch_count() - launches0) * steps // (steps + warmup)
    kernel_ms = kernel_ms_of(ctx, [patch], lambda p: render(p))[0]
    kid = patch.kernel_id(cnt)  # (after the first renders: the launch shape in use may be a measured choice)
    rec = {"workload": f"{name} @ {V_gpu} voices per GPU x {N_SAMPLES} samples, stems + mix in HBM", "voices_per_gpu": V_gpu,
           "voices_total": V_total, "value": V_total * N_SAMPLES / (ms_step * 1e-3), "unit": UNIT, "ms_per_step": ms_step,
           "steps": steps, "warmup": warmup, "kernel_ms": kernel_ms, "gpu_launches": int(launches), **launch_shape(patch.program_info(cnt)),
           "l2": f"each step writes {C * N_SAMPLES * cnt * 4 / 1e6:.0f} MB of stems per GPU (> 126 MB L2), no flush needed",
           "schedule": patch.schedule_report()}

    # ---- the result, checked in the same run
    patch.reset()
    render()
    torch.cuda.synchronize()
    if world == 1:
        rec["mix_check"] = check_mix_against_stems(ctx, mix, stems, cnt)
    else:
        partial = mix.clone()
        srk.shard.reduce_mix(mix)

        def rerender():
            o2, c2 = srk.shard.voice_range(V_total, world - 1, world)
            p2 = ctx.patch(name, V_total)
            m2 = torch.empty((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev)
            p2.render_into(c2, N_SAMPLES, o2, ctx.stems(c2).data_ptr(), m2.data_ptr(), device_out=True, stream=ctx.stream)
            return m2
        rec["mix_check"] = check_reduced_mix(ctx, mix, partial, V_total, rerender)

    # ---- e2e: host buffers (per-voice params H2D + mix D2H every step; the NCCL sum inside at N > 1)
    if e2e:
        pv = list(patch.per_voice.items())  # ((module, param id), host array) set by the patch description
        mix_host = torch.empty((C, N_SAMPLES), dtype=torch.float32).pin_memory()
        h2d = patch.program_info(cnt)["param_words"] * cnt * 4
        d2h = C * N_SAMPLES * 4

        def step_e2e():
            for (m, pid), arr in pv:
                m.set_param_per_voice(pid, arr)  # marks the table dirty -> rebuilt in pinned staging + H2D
            if world == 1:
                patch.render_into(cnt, N_SAMPLES, off, None, mix_host.data_ptr(), device_out=False)  # the plain C-ABI call
            else:
                patch.render_into(cnt, N_SAMPLES, off, None, mix.data_ptr(), device_out=True, async_=True, stream=ctx.stream)
                srk.shard.reduce_mix(mix)
                if rank == 0:
                    mix_host.copy_(mix, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            return float(mix_host[0, -1])

        for _ in range(warmup):
            step_e2e()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_s = ctx.max_over_ranks((time.perf_counter() - t0) / steps)[0]
        rec["e2e"] = {"value": V_total * N_SAMPLES / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "ms_per_step": e2e_s * 1e3,
                      "what": "MIX-ONLY: per-voice params from host memory + render" + (" + NCCL sum of the mix" if world > 1 else "") +
                              " + stereo mix to pinned host memory, every step; the per-voice stems stay on the device (see e2e_stems)"}
    clocks = sampler.summary() if sampler else None
    rec["roofline"] = roofline(name, cnt, kernel_ms, kid, (clocks or {}).get("sm_mhz"))
    if name == "cfg2":
        floor_ms = N_SAMPLES * LADDER_CHAIN_CYCLES / (((clocks or {}).get("sm_mhz") or sm_max_mhz()) * 1e3)
        rec["roofline"]["latency_floor"] = {"ms": floor_ms, "frac": floor_ms / kernel_ms,
                                            "what": f"{N_SAMPLES} samples x {LADDER_CHAIN_CYCLES:.0f} cycles (the ladder filter's dependent chain of "
                                                    "one voice) / SM clock: the bound at few voices per SM"}
    return rec, clocks, patch, (off, cnt)


def single_gpu_config(ctx, name, V, steps=4, warmup=2):
    """One BASELINE config on one GPU as a sub-record: device-timed steps, voice-kernel time, roofline, mix check."""
    torch = ctx.torch
    p = ctx.patch(name, V)
    stems = ctx.stems(V)
    mix = torch.empty((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev)

    def step():
        p.render_into(V, N_SAMPLES, 0, stems.data_ptr(), mix.data_ptr(), device_out=True, async_=True, stream=ctx.stream)

    launches0 = p.launch_count()
    ms = timed_steps(ctx, step, steps, warmup)
    launches = (p.launch_count() - launches0) * steps // (steps + warmup)
    k = kernel_ms_of(ctx, [p], lambda q: step())[0]
    p.reset()
    step()
    torch.cuda.synchronize()
    rec = {"workload": f"{name} @ {V} voices x {N_SAMPLES} samples, stems + mix in HBM", "voices": V,
           "value": V * N_SAMPLES / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kernel_ms": k, "steps": steps, "warmup": warmup,
           "roofline": roofline(name, V, k, p.kernel_id(V)), "gpu_launches": int(launches), **launch_shape(p.program_info(V)),
           "mix_check": check_mix_against_stems(ctx, mix, stems, V), "schedule": p.schedule_report(),
           "timing": f"CUDA events on the launching stream around {steps} steps; {C * N_SAMPLES * V * 4 / 1e9:.1f} GB of stems per step (> L2)"}
    return rec


def block_cadence(ctx, patch, off, cnt):
    """The reference's own call pattern: one execute() per buffer_size block -- srk_render per
    1024 samples, mix to the device, synchronous per call.  47 blocks ~ one second."""
    torch = ctx.torch
    mix = torch.empty((C, 1024), dtype=torch.float32, device=ctx.dev)
    patch.reset()
    for _ in range(3):
        patch.render_into(cnt, 1024, off, None, mix.data_ptr(), device_out=True)
    torch.cuda.synchronize()
    n = 47
    t0 = time.perf_counter()
    for _ in range(n):
        patch.render_into(cnt, 1024, off, None, mix.data_ptr(), device_out=True)
    dt = time.perf_counter() - t0
    return {"block_samples": 1024, "blocks": n, "ms_per_block": dt / n * 1e3, "value": cnt * 1024 * n / dt, "unit": UNIT,
            "what": "srk_render(1024 samples) per call, mix to HBM, stream synchronised per call (wall clock)"}


def graphs_on_ranks(ctx, V, steps, warmup):
    """BASELINE configs[4]: 8 distinct patch graphs x V voices; whole graphs are dealt to the ranks (one graph per GPU at
    N = 8), longest-processing-time first by the kernel times measured in this run; a rank renders its graphs one
    after the other and sums their mixes, then the one NCCL sum."""
    torch, srk, dist = ctx.torch, ctx.srk, ctx.dist
    graphs = srk.patches.CFG5_GRAPHS
    names = [g.__name__ for g in graphs]
    world, rank = ctx.world, ctx.rank
    stems = ctx.stems(V)
    scratch = torch.empty((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev)
    # calibration: every rank measures every graph's kernel once (also loads the kernels), max over ranks -> weights
    weights = []
    for n in names:
        p = ctx.patch(n, V)
        for _ in range(2):
            p.render_into(V, N_SAMPLES, 0, stems.data_ptr(), scratch.data_ptr(), device_out=True, stream=ctx.stream)
        weights.append(p.last_render_ms()[0])
        del p
    weights = ctx.max_over_ranks(*weights)
    owner, load = lpt(weights, world)
    mine = [g for g in range(len(graphs)) if owner[g] == rank]
    patches = [(g, ctx.patch(names[g], V)) for g in mine]
    mixes = [torch.empty((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev) for _ in mine]
    total = torch.zeros((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev)

    def local():
        for (g, p), mx in zip(patches, mixes):
            p.render_into(V, N_SAMPLES, 0, stems.data_ptr(), mx.data_ptr(), device_out=True, async_=True, stream=ctx.stream)
        if mixes:
            torch.sum(torch.stack(mixes), dim=0, out=total)
        else:
            total.zero_()

    def step():
        local()
        if world > 1:
            srk.shard.reduce_mix(total)

    launches0 = sum(p.launch_count() for _, p in patches)
    ms = timed_steps(ctx, step, steps, warmup)
    launches = (sum(p.launch_count() for _, p in patches) - launches0) * steps // (steps + warmup)
    per_graph = {}
    for g, p in patches:
        info = p.program_info(V)
        per_graph[names[g]] = {"rank": rank, "kernel_ms": kernel_ms_of(ctx, [p], lambda q: q.render_into(
            V, N_SAMPLES, 0, stems.data_ptr(), scratch.data_ptr(), device_out=True, stream=ctx.stream), reps=2)[0],
            "kernel": p.kernel_id(V), **launch_shape(info)}
    for _, p in patches:
        p.reset()
    local()
    torch.cuda.synchronize()
    partial = total.clone()
    if world > 1:
        srk.shard.reduce_mix(total)
        gathered = [None] * world
        dist.all_gather_object(gathered, per_graph)
        per_graph = {k: v for d in gathered for k, v in d.items()}

        def rerender():  # everything the last rank rendered, again on rank 0
            acc = []
            for g in [g for g in range(len(graphs)) if owner[g] == world - 1]:
                q = ctx.patch(names[g], V)
                m = torch.empty((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev)
                q.render_into(V, N_SAMPLES, 0, stems.data_ptr(), m.data_ptr(), device_out=True, stream=ctx.stream)
                acc.append(m)
            return torch.sum(torch.stack(acc), dim=0) if acc else torch.zeros_like(total)
        check = check_reduced_mix(ctx, total, partial, len(graphs) * V, rerender)
    else:
        check = {"result": "ok" if bool(torch.isfinite(total).all()) and float(total.abs().max()) > 0 else "FAILED",
                 "what": "single GPU: finite, non-silent sum of the 8 graphs' mixes"}
    n_vs = len(graphs) * V * N_SAMPLES
    rec = None
    if rank == 0:
        peak, _ = measured_peak_hbm()
        algo = sum(bytes_per_voice_sample(n) * V * N_SAMPLES for n in names)
        k_sum = sum(v["kernel_ms"] for v in per_graph.values())
        rec = {"workload": f"cfg5: mixed batch, 8 distinct patch graphs x {V} voices each, 48 kHz x 1 s; whole graphs dealt to the ranks",
               "graphs": names, "graphs_per_gpu": len(graphs) / world, "voices_total": len(graphs) * V, "scaling": "strong",
               "value": n_vs / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
               "assignment": {names[g]: owner[g] for g in range(len(graphs))}, "rank_load_ms": load,
               "limiter": max(per_graph, key=lambda k: per_graph[k]["kernel_ms"]),
               "roofline": {"bound": "hbm", "achieved": algo / (k_sum * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": algo / (k_sum * 1e-3) / 1e9 / peak, "traffic": None,
                            "kernel": "sum over the 8 graphs' voice-kernel launches", "kernel_ms": k_sum},
                "per_graph": per_graph, "gpu_launches": int(launches), "mix_check": check}
    return rec, total.clone(), weights


def lpt(times, world):
    """Longest-processing-time-first: -> (owner per item, load per rank)."""
    load, owner = [0.0] * world, [0] * len(times)
    for i in sorted(range(len(times)), key=lambda i: (-times[i], i)):
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += times[i]
    return owner, load


def graphs_balanced(ctx, V, steps, warmup, whole_ms, reference_mix=None):
    """cfg5 beyond one graph per GPU: the work is dealt as (graph, voice range) PIECES.  A graph whose kernel is longer than
    a rank's fair share is cut into two half ranges (voice_offset) when that shortens the longest rank -- with the half's
    kernel time measured in this run, because a half costs more than half (fewer voice groups per SM: less latency hiding).
    Pieces go to ranks longest first; a rank renders its pieces one after the other, sums their mixes, then the one NCCL sum."""
    torch, srk = ctx.torch, ctx.srk
    graphs = srk.patches.CFG5_GRAPHS
    names = [g.__name__ for g in graphs]
    world, rank = ctx.world, ctx.rank
    stems = ctx.stems(V)
    scratch = torch.empty((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev)
    half = V // 2
    fair = sum(whole_ms) / world
    half_ms = {}
    for g in range(len(graphs)):  # candidates: measured on every rank, max over ranks -> the same decision everywhere
        if whole_ms[g] > 0.6 * fair and world > 1:  # (the greedy step below decides; this only bounds what is measured)
            p = ctx.patch(names[g], V)
            for _ in range(2):
                p.render_into(half, N_SAMPLES, 0, stems.data_ptr(), scratch.data_ptr(), device_out=True, stream=ctx.stream)
            half_ms[g] = p.last_render_ms()[0]
            del p
    if half_ms:
        vals = ctx.max_over_ranks(*[half_ms[g] for g in sorted(half_ms)])
        half_ms = dict(zip(sorted(half_ms), vals))
    split = set()
    for g in sorted(half_ms, key=lambda g: -whole_ms[g]):
        def makespan(sp):
            t = [x for i in range(len(graphs)) for x in ([half_ms[i]] * 2 if i in sp else [whole_ms[i]])]
            return max(lpt(t, world)[1])
        if makespan(split | {g}) < makespan(split):
            split.add(g)
    pieces = []  # (graph, voice offset, voices, estimated ms)
    for g in range(len(graphs)):
        if g in split:
            pieces += [(g, 0, half, half_ms[g]), (g, half, V - half, half_ms[g])]
        else:
            pieces.append((g, 0, V, whole_ms[g]))
    owner, load = lpt([p[3] for p in pieces], world)
    mine = [pieces[i] for i in range(len(pieces)) if owner[i] == rank]
    patches = [(ctx.patch(names[g], V), off, cnt) for g, off, cnt, _ in mine]
    mixes = [torch.empty((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev) for _ in mine]
    total = torch.zeros((C, N_SAMPLES), dtype=torch.float32, device=ctx.dev)

    def local():
        for (p, off, cnt), mx in zip(patches, mixes):
            p.render_into(cnt, N_SAMPLES, off, stems.data_ptr(), mx.data_ptr(), device_out=True, async_=True, stream=ctx.stream)
        if mixes:
            torch.sum(torch.stack(mixes), dim=0, out=total)
        else:
            total.zero_()

    def step():
        local()
        if world > 1:
            srk.shard.reduce_mix(total)

    launches0 = sum(p.launch_count() for p, _, _ in patches)
    ms = timed_steps(ctx, step, steps, warmup)
    launches = (sum(p.launch_count() for p, _, _ in patches) - launches0) * steps // (steps + warmup)
    for p, _, _ in patches:
        p.reset()
    local()
    torch.cuda.synchronize()
    partial = total.clone()
    check = None
    if world > 1:
        srk.shard.reduce_mix(total)
        check = check_reduced_mix(ctx, total, partial, len(graphs) * V)
    if rank != 0:
        return None
    if reference_mix is not None:  # the same job summed in another order: equal within the mix tolerance
        err = (total.double() - reference_mix.double()).abs()
        bound = 1e-5 * torch.maximum(reference_mix.double().abs(), torch.tensor(float(len(graphs) * V) ** 0.5, dtype=torch.float64, device=ctx.dev))
        same = bool((err <= bound).all())
        check = dict(check or {"result": "ok"})
        check["equals_whole_graph_assignment_within_tolerance"] = same
        check["max_err_over_bound_vs_whole_graphs"] = float((err / bound).max())
        if not same:
            check["result"] = "FAILED"
    return {"workload": f"cfg5 balanced: 8 graphs x {V} voices dealt to {world} ranks as (graph, voice range) pieces, longest first",
            "voices_total": len(graphs) * V, "scaling": "strong", "value": len(graphs) * V * N_SAMPLES / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms, "steps": steps, "warmup": warmup, "fair_share_ms": fair,
            "split_in_two": {names[g]: {"whole_ms": whole_ms[g], "half_ms": half_ms[g]} for g in sorted(split)},
            "pieces": [{"graph": names[g], "voice_offset": off, "voices": cnt, "rank": owner[i], "estimated_ms": t}
                       for i, (g, off, cnt, t) in enumerate(pieces)],
            "rank_load_ms": load, "gpu_launches": int(launches), "mix_check": check}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(BYTES_PER_VOICE_SAMPLE) + ["cfg5"])
    ap.add_argument("--voices-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        return reference_arm(args)

    ctx = Ctx(args)
    srk, torch, world, rank = ctx.srk, ctx.torch, ctx.world, ctx.rank
    if args.config == "cfg5":
        rec, total, weights = graphs_on_ranks(ctx, args.voices_per_gpu or srk.patches.CFG5_VOICES, args.steps, args.warmup)
        bal = graphs_balanced(ctx, args.voices_per_gpu or srk.patches.CFG5_VOICES, args.steps, args.warmup, weights, total)
        if rank == 0:
            line = {"metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": DTYPE, "data": "synthetic", "config": {"workload": rec["workload"], "parallelism": f"graph-shard x{world}"},
                    "roofline": rec["roofline"], "cpu_baseline": None, "cfg5": rec, "cfg5_balanced": bal,
                    "gpu_launches": rec["gpu_launches"], "mix_check": rec["mix_check"]["result"]}
            print(json.dumps(line))
        if world > 1:
            ctx.dist.barrier()
            ctx.dist.destroy_process_group()
        return 0

    desc = srk.patches.CONFIGS[args.config][2]
    V_gpu = args.voices_per_gpu or (min(srk.patches.CONFIGS[args.config][1], 65536) if world == 1
                                    else {"cfg4": 32768}.get(args.config, min(srk.patches.CONFIGS[args.config][1], 65536)))
    try:
        uuid = str(torch.cuda.get_device_properties(ctx.local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(ctx.local_rank, uuid) if rank == 0 else None
    if sampler:
        sampler.start()
    head, clocks, patch, (off, cnt) = voice_shard(ctx, args.config, V_gpu, args.steps, args.warmup, e2e=True, sampler=sampler)

    extras = {}
    default_run = args.config == "cfg2" and not args.voices_per_gpu and not args.no_extras
    if world == 1:
        # ---- e2e with all stems to pinned host memory as well (PCIe bound), bounded memory
        if cnt * N_SAMPLES * C * 4 <= 4 << 30:
            stems_host = torch.empty((C, N_SAMPLES, cnt), dtype=torch.float32).pin_memory()
            mix_host = torch.empty((C, N_SAMPLES), dtype=torch.float32).pin_memory()
            pv = list(patch.per_voice.items())

            def step_e2e_stems():
                for (m, pid), arr in pv:
                    m.set_param_per_voice(pid, arr)
                patch.render_into(cnt, N_SAMPLES, off, stems_host.data_ptr(), mix_host.data_ptr(), device_out=False)

            step_e2e_stems()
            t0 = time.perf_counter()
            n_rep = max(2, min(args.steps, 5))
            for _ in range(n_rep):
                step_e2e_stems()
            dt = (time.perf_counter() - t0) / n_rep
            extras["e2e_stems"] = {"value": cnt * N_SAMPLES / dt, "unit": UNIT, "h2d_bytes_per_step": head["e2e"]["h2d_bytes_per_step"],
                                   "d2h_bytes_per_step": C * N_SAMPLES * 4 + C * N_SAMPLES * cnt * 4, "ms_per_step": dt * 1e3,
                                   "what": "as e2e, plus every voice's stems to pinned host memory (PCIe-bound)"}
            del stems_host
        if default_run:
            extras["block_cadence"] = block_cadence(ctx, patch, off, cnt)
            extras["configs"] = {
                "cfg2_65536": single_gpu_config(ctx, "cfg2", 65536),    # the full chip on the headline patch
                "cfg3_65536": single_gpu_config(ctx, "cfg3", 65536),    # BASELINE configs[2], feed-forward FM
                "cfg3b_65536": single_gpu_config(ctx, "cfg3b", 65536),  # BASELINE configs[2], in-graph feedback (one cut wire)
                "cfg4_32768": single_gpu_config(ctx, "cfg4", 32768),    # one GPU's share of BASELINE configs[3]
            }
            extras["full_chip"] = extras["configs"]["cfg2_65536"]
            # BASELINE configs[4] on one GPU: the N = 1 point of its strong-scaling curve (N > 1: graphs dealt to ranks)
            g, _, _ = graphs_on_ranks(ctx, srk.patches.CFG5_VOICES, 3, 3)
            extras["cfg5"] = g
    elif default_run:
        cfg4, _, _, _ = voice_shard(ctx, "cfg4", 32768, max(3, min(args.steps, 5)), 3, e2e=False)
        extras["cfg4_shard"] = cfg4
        g, total, weights = graphs_on_ranks(ctx, srk.patches.CFG5_VOICES, max(3, min(args.steps, 5)), 3)
        extras["cfg5"] = g
        extras["cfg5_balanced"] = graphs_balanced(ctx, srk.patches.CFG5_VOICES, max(3, min(args.steps, 5)), 3, weights, total)

    if rank == 0:
        cpu = cpu1 = None
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(args.config, steps=3, warmup=1, threads=os.cpu_count() or 1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
            r1 = cpu_reference_run(args.config, steps=2, warmup=1, threads=1, sample_voices=32)
            cpu1 = {"value": r1["value"], "unit": UNIT, "cores": 1, "kind": "port", "sample": r1["sample"],
                    "note": "the reference's own execution model: one audio thread (main.rs:59-63)"}
        checks = [head["mix_check"]] + [v.get("mix_check") for v in extras.values() if isinstance(v, dict) and "mix_check" in v]
        checks += [v["mix_check"] for v in extras.get("configs", {}).values()]
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": f"{args.config}: {desc}", "voices_per_gpu": V_gpu, "voices_total": V_gpu * world,
                       "n_samples": N_SAMPLES, "sample_rate": 48000, "parallelism": f"voice-shard x{world}",
                       "outputs": "stems [2][48000][V] + mix [2][48000] in HBM", "l2": head["l2"],
                       **{k: head[k] for k in ("block_threads", "step_samples", "smem_bytes", "warps_per_voice_group",
                                               "pipeline_stages", "voice_groups_per_block", "fused", "registers")}},
            "roofline": dict(head["roofline"], note="latency/issue-bound DSP recurrences: at 4096 voices the chip holds 128 voice groups "
                             "on 148 SMs and the HBM fraction is small by construction (SURVEY.md §8d); see latency_floor, "
                             "roofline_issue and configs.cfg2_65536 for the full chip"),
            "cpu_baseline": cpu, "cpu_baseline_1thread": cpu1,
            "e2e": head["e2e"],
            "mix_check": "ok" if all(c and c.get("result") == "ok" for c in checks) else "FAILED",
            "mix_check_detail": head["mix_check"],
            "schedule": head["schedule"],
            "gpu_launches": head["gpu_launches"],
            "clocks": clocks,
            **extras,
        }
        print(json.dumps(line))
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
