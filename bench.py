#!/usr/bin/env python
"""bench.py - motion frames/s of the EMAGE audio->motion hot path (BASELINE.json metric).

One "step" = the reference demo's timed span (test_emage_audio.py:32-47) over one batch of synthetic
16 kHz audio: EmageAudioModel.inference() + the full-length EmageVQModel.decode(get_global_motion=True).
Workload at N=1: BASELINE configs[1], 32 clips x 10 s (300 frames each, 9 600 frames per step); for
N>1 each rank runs its own 32 clips (weak scaling, configs[4]); weights are broadcast from rank 0 once
at load (NCCL) and there is no collective inside the step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = device-timed frames/s with inputs resident in HBM; `e2e` = the
same span through the public API with pinned HOST buffers (H2D audio + D2H results inside the timed
region, wall clock).  `roofline` = the dominant kernel (tap-GEMM) against the measured bf16 peak,
`roofline_vq` = the VQ lookup kernel against the measured HBM bandwidth on 2^21 rows, `extra` = the other
BASELINE configs (bs = 1 latency, CaMN bs 64, DisCo bs 32).

`--impl reference` times the UNMODIFIED reference modules (byte-compiled into oracle/_ref by
oracle/make_ref.py; falls back to the oracle port when that tree is absent) on the host cores, on the same
32-clip step, same warm-up count.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIPS_PER_GPU = 32
N_SAMPLES = 160000                    # 10 s @ 16 kHz -> 300 frames
FRAMES_PER_CLIP = 300
FLOP_PER_FRAME = 364.7e6              # algorithmic work per emitted frame
DTYPES = {"fp32": "f32", "fp16x3": "fp16x3 (two fp16 operand planes, 3 products, f32 accumulate)",
          "bf16x6": "bf16x6 (three bf16 operand planes, 6 products, f32 accumulate)",
          "bf16x3": "bf16x3 (two bf16 operand planes, 3 products, f32 accumulate)", "bf16": "bf16"}
ENGINES = {"fp32": "fp32 SIMT tap-GEMM", "fp16x3": "tcgen05 tap-GEMM, 3 fp16 products per fp32 product",
           "bf16x6": "tcgen05 tap-GEMM, 6 bf16 products per fp32 product",
           "bf16x3": "tcgen05 tap-GEMM, 3 bf16 products per fp32 product", "bf16": "tcgen05 tap-GEMM, plain bf16"}
MMA_PER_PRODUCT = {"fp32": 0, "bf16": 1, "bf16x3": 3, "bf16x6": 6, "fp16x3": 3}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE representative launch from a committed `ncu --set full` capture
# (a static number, not measured in this run): precision -> (bytes, source file under profiles/)
NCU_TRAFFIC = {"bf16x6": (26.4e6, "profiles/ncu_full_r1_final.md"),
               "fp16x3": (8.711e6, "profiles/r2/ncu_full.md (8.711 MB read + 0 B written; algorithmic operand bytes 8.65 MB)")}
METRIC = "motion_frames_per_sec"
UNIT = "frames/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# -----------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref) or, without it, the oracle port - on the host cores
# -----------------------------------------------------------------------------------------------------


class CpuArm:
    """The reference path on the host cores.  kind = "reference": the unmodified reference modules, byte-compiled by
    oracle/make_ref.py; "port": oracle/emage_oracle.py (same ATen CPU ops) when the staged tree is absent."""

    def __init__(self):
        import torch
        from oracle import ref_loader
        self.torch = torch
        self.kind = "reference" if ref_loader.staged_available() else "port"
        if self.kind == "reference":
            self.ref = ref_loader.import_reference()
            self.model, self.vqm = ref_loader.build_emage(self.ref, seed=0)
            self._drive = ref_loader.drive_like_demo
        else:
            from oracle import emage_oracle as O
            from oracle.weights import make_checkpoint
            self.sd, self.cfg, self.vq = make_checkpoint(seed=0)
            self.O = O
        self.threads = None

    def run(self, clips, seed=1234):
        from oracle.weights import synth_audio
        torch = self.torch
        audio = torch.from_numpy(synth_audio(clips, N_SAMPLES, seed))
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.kind == "reference":
                self._drive(self.model, self.vqm, audio)
            else:
                self.O.emage_generate(self.sd, self.cfg, self.vq, audio, torch.zeros(clips, 1, dtype=torch.long))
        return time.perf_counter() - t0

    def pick_threads(self):
        """Best of a quick probe over {min(cores, 64), 32, 16, 8}: on many-core hosts the small per-window ops of this
        model run slower with every core than with a subset, and the baseline should be the CPU at its best."""
        torch = self.torch
        cores = os.cpu_count() or 1
        cands = sorted({c for c in (min(cores, 64), 32, 16, 8) if c <= cores}, reverse=True) or [cores]
        best, best_t = cands[0], None
        torch.set_num_threads(cands[0])
        self.run(1)                                       # warm-up: allocator, oneDNN primitive caches
        for c in cands:
            torch.set_num_threads(c)
            t = min(self.run(2), self.run(2))
            if best_t is None or t < best_t:
                best, best_t = c, t
        torch.set_num_threads(best)
        self.threads = best
        return best

    def time_steps(self, clips, steps, warmup):
        if self.threads is None:
            self.pick_threads()
        times = []
        for i in range(warmup + steps):
            t = self.run(clips)
            if i >= warmup:
                times.append(t)
        return times

    def describe(self, clips, times):
        fps = [clips * FRAMES_PER_CLIP / t for t in times]
        return {"value": clips * FRAMES_PER_CLIP * len(times) / sum(times), "unit": UNIT, "cores": self.threads,
                "host_cores": os.cpu_count(), "kind": self.kind,
                "sample": f"{clips} clips x {FRAMES_PER_CLIP} frames per run x {len(times)} runs, fp32, "
                          + ("unmodified reference modules (oracle/_ref)" if self.kind == "reference" else "oracle port (torch CPU ops)")
                          + "; threads = best of a probe over the host's cores",
                "runs_frames_per_s": [round(v, 1) for v in fps]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm()
    clips = CLIPS_PER_GPU                                 # the full step of the GPU arm, not a sample of it
    times = arm.time_steps(clips, args.steps, args.warmup)
    base = arm.describe(clips, times)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "EMAGE batch32x300f (BASELINE configs[1]) on the host CPU", "clips": clips,
                   "frames_per_clip": FRAMES_PER_CLIP, "audio_samples": N_SAMPLES,
                   "weights": "synthetic seeded checkpoint (oracle/weights.py), reference key layout"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -----------------------------------------------------------------------------------------------------
# GPU arm
# -----------------------------------------------------------------------------------------------------


def instrumented_gemm_pass(run_step, ops):
    """One extra (untimed-for-`value`) eager step on ONE stream in which every tap-GEMM launch is bracketed by CUDA
    events: returns dict(flop, ms, launches, weight_bytes, wall_ms).  Branches that overlap on forked streams in the
    graph-replayed step run back to back here, so sum(ms) is a per-kernel busy time, not a share of `ms_per_step`."""
    import torch
    from pantomatrix_b200.emage_audio import engine
    records = []
    real = ops._call

    def traced(name, *a):
        if name in ("pm_tapgemm_f32", "pm_tapgemm_tc"):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            real(name, *a)
            e.record()
            if name == "pm_tapgemm_f32":      # (A,a_bs,lda,batch,rows_in,cin,W,bias,taps,stride,pad,rows_out,cout,...)
                batch, cin, taps, rows_out, cout = a[3], a[5], a[8], a[11], a[12]
                wbytes = 4 * taps * cout * cin
            else:      # (A,a_ps,a_bs,lda,batch,rows_in,cin,W,w_ps,w_rows,ldw,taps,pad,nsplit,bias,rows_out,cout,..)
                batch, cin, taps, rows_out, cout = a[4], a[6], a[11], a[15], a[16]
                wbytes = 2 * (a[13] & 0xFF) * taps * a[9] * a[10]
                if taps * cin in (18 * 64, 18 * 128):      # k=15 stride-6/3 convs run as 3x(6C) / 5x(3C) taps:
                    taps, cin = 15, cin * taps // 18       # count the algorithmic 15 taps, not the zero padding
            records.append((2.0 * batch * rows_out * cout * cin * taps, s, e, wbytes))
        else:
            real(name, *a)

    ops._call = traced
    engine._STATE["fork"] = False          # one stream, so each event pair brackets exactly one kernel
    try:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_step()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    finally:
        ops._call = real
        engine._STATE["fork"] = True
    return dict(flop=sum(r[0] for r in records), ms=sum(r[1].elapsed_time(r[2]) for r in records), launches=len(records),
                weight_bytes=sum(r[3] for r in records), wall_ms=1e3 * wall)


def _device_time(fn, steps, warmup, flush=None):
    """ms per call of fn(): CUDA events around each call, `flush` (a > L2 buffer) rewritten between calls."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i, (s, e) in enumerate(ev):
        if flush is not None:
            flush.fill_(i & 0xFF)
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in ev) / steps


def extra_configs(dev, peaks, flush, precision, cpu):
    """BASELINE configs[0] (one 10 s clip: latency, weight-bandwidth bound), [2] CaMN bs 64, [3] DisCo bs 32."""
    import torch
    from oracle.weights import synth_audio
    from pantomatrix_b200 import ops
    from pantomatrix_b200.pipeline import CapturedPipeline, generate
    from synthetic_models import build_lstm_product, build_product
    out = {}
    # ---- configs[0]: bs = 1
    model, vqm = build_product(seed=0, device=dev)
    host1 = torch.from_numpy(synth_audio(1, N_SAMPLES, 99)).pin_memory()
    cap = CapturedPipeline(model, vqm, 1, N_SAMPLES)
    cap.audio.copy_(host1)
    ms = _device_time(cap.graph.replay, 10, 3, flush)
    t_e2e = []
    for i in range(5):
        flush.fill_(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pred = cap(host1)[1]
        host_out = {k: pred[k].cpu() for k in ("motion_axis_angle", "expression", "trans")}
        t_e2e.append(time.perf_counter() - t0)
    del host_out
    inst = instrumented_gemm_pass(lambda: generate(model, vqm, cap.audio), ops)
    wbw = inst["weight_bytes"] / (ms * 1e-3) / 1e9
    out["configs[0] EMAGE single 10 s clip"] = {
        "latency_ms": ms, "frames_per_s": FRAMES_PER_CLIP / ms * 1e3, "e2e_latency_ms": 1e3 * sorted(t_e2e)[len(t_e2e) // 2],
        "timer": "CUDA events around one graph replay, 256 MB L2 flush between replays; e2e = wall clock incl. H2D/D2H",
        "weight_bytes_streamed_per_step": inst["weight_bytes"], "weight_stream_GBps": wbw, "hbm_peak_GBps": peaks["hbm"],
        "weight_bw_frac": wbw / peaks["hbm"],
        "note": "bs = 1 is weight-bandwidth / latency bound: every window re-reads the packed operand planes of the trunk "
                "(they do not fit L2 together with the other windows' working set); weight_bw_frac = packed weight bytes "
                "of all GEMM launches of one step / latency / measured HBM copy bandwidth"}
    del cap, model, vqm
    # ---- configs[2], [3]: CaMN bs 64, DisCo bs 32 (emitted 15-fps frames per second)
    for key, kind, bs in (("configs[2] CaMN batch 64", "camn", 64), ("configs[3] DisCo batch 32", "disco", 32)):
        m = build_lstm_product(kind, device=dev)
        audio = torch.from_numpy(synth_audio(bs, N_SAMPLES, 7)).to(dev)
        spk = torch.zeros(bs, 1, dtype=torch.long, device=dev)
        res = {}

        def step():
            res["o"] = m(audio, spk)
        ms = _device_time(step, 5, 3, flush)
        frames = bs * res["o"]["motion"].shape[1]
        entry = {"ms_per_batch": ms, "frames_per_s_15fps": frames / ms * 1e3, "batch": bs,
                 "frames_per_clip": int(res["o"]["motion"].shape[1]), "precision": precision,
                 "timer": "CUDA events, eager launches, L2 flush between batches"}
        if cpu is not None and cpu.kind == "reference":
            from oracle import ref_loader
            rm = ref_loader.build_lstm(cpu.ref, kind, seed=0)
            a8 = torch.from_numpy(synth_audio(8, N_SAMPLES, 7))
            s8 = torch.zeros(8, 1, dtype=torch.long)
            with torch.no_grad():
                rm(a8[:2], s8[:2])
                t0 = time.perf_counter()
                ro = rm(a8, s8)
                cpu_s = time.perf_counter() - t0
            entry["cpu_reference_frames_per_s_15fps"] = 8 * ro["motion"].shape[1] / cpu_s
            entry["cpu_sample"] = f"8 clips, unmodified reference module, {torch.get_num_threads()} threads"
        out[key] = entry
        del m
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from oracle.weights import synth_audio
    from pantomatrix_b200 import ops
    from pantomatrix_b200.emage_audio import engine
    from pantomatrix_b200.pipeline import CapturedPipeline, generate
    from synthetic_models import build_product
    engine.set_precision(args.precision)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weights: rank 0's checkpoint is the job's checkpoint, broadcast once at load
    model, vqm = build_product(seed=0, device=dev)
    if world > 1:
        from pantomatrix_b200.sharding import broadcast_checkpoint
        broadcast_checkpoint(model, vqm, src=0)

    clips = CLIPS_PER_GPU
    host_audio = torch.from_numpy(synth_audio(clips, N_SAMPLES, 1234 + rank)).pin_memory()
    audio = host_audio.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    frames_per_step = clips * FRAMES_PER_CLIP

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    cap = CapturedPipeline(model, vqm, clips, N_SAMPLES, body_priority=bool(args.body_priority)) if args.graph else None
    if cap is not None:
        cap.audio.copy_(audio)

    def step_resident():
        if cap is not None:
            cap.graph.replay()
            ops.launch_count += cap.kernels_per_replay
        else:
            generate(model, vqm, audio)

    def step_e2e():
        if cap is not None:
            return cap(host_audio)[1]
        return generate(model, vqm, host_audio.to(dev, non_blocking=True))[1]

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step_resident()
    sync_all()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- value: device-timed, inputs resident in HBM ----
    launches0 = ops.launch_count
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sync_all()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                   # evict L2 between timed iterations (outside the bracket)
        starts[i].record()
        step_resident()
        ends[i].record()
    sync_all()
    launches = ops.launch_count - launches0
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))

    # ---- e2e: pinned host audio -> H2D -> public API -> D2H of the emitted SMPL-X parameters ----
    out_host = {k: torch.empty(clips, FRAMES_PER_CLIP, d).pin_memory() for k, d in
                (("motion_axis_angle", 165), ("expression", 100), ("trans", 3))}
    h2d = host_audio.numel() * 4
    d2h = sum(v.numel() * 4 for v in out_host.values())
    e2e_times = []
    sync_all()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pred = step_e2e()
        for k, v in out_host.items():
            v.copy_(pred[k], non_blocking=True)
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    sync_all()
    clocks = sampler.stop() if rank == 0 else None

    t_dev = torch.tensor([dev_ms / 1e3, sum(e2e_times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = t_dev.tolist()

    if rank == 0:
        peaks = _peaks()
        value = world * frames_per_step * args.steps / dev_s
        e2e_value = world * frames_per_step * args.steps / e2e_s
        inst = instrumented_gemm_pass(lambda: generate(model, vqm, audio), ops)
        achieved = inst["flop"] / (inst["ms"] * 1e-3) / 1e12 if inst["ms"] > 0 else 0.0
        peak = peaks["bf16_sustained"]
        traffic = NCU_TRAFFIC.get(args.precision)
        # ---- VQ lookup kernel against HBM, >= 10^6 rows
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        try:
            import bench_vq
            roofline_vq = bench_vq.measure(1 << 21, 20)
        except Exception as exc:                                    # keep the headline line even if the extra fails
            roofline_vq = {"error": repr(exc)}
        cpu = cpu_line = None
        if args.cpu_baseline:
            cpu = CpuArm()
            cpu_line = cpu.describe(clips, cpu.time_steps(clips, 3, 1))
        extra = None
        if world == 1 and args.extra:
            del cap
            try:
                extra = extra_configs(dev, peaks, flush, args.precision, cpu)
            except Exception as exc:
                extra = {"error": repr(exc)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPES[args.precision], "data": "synthetic",
            "config": {"workload": "EMAGE batch32x300f per GPU (BASELINE configs[1]; configs[4] at 8 GPUs)",
                       "clips_per_gpu": clips, "frames_per_clip": FRAMES_PER_CLIP, "audio_samples": N_SAMPLES,
                       "weights": "synthetic seeded checkpoint (oracle/weights.py), reference key layout",
                       "engine": ENGINES[args.precision], "cuda_graph": bool(args.graph), "l2": "256 MB flush between timed steps; weights 0.56 GB > L2",
                       "parallelism": f"dp{world} (clip sharding, NCCL weight broadcast at load only)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "timer": "wall clock incl. H2D/D2H, pinned host buffers"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "tapgemm_tc_kernel (conv1d + linear), all launches of one step",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic[0] if traffic else None,
                         "traffic_source": (f"static: one representative launch (M=2048, N=K=768 trunk GEMM) of the committed "
                                            f"ncu --set full capture {traffic[1]}, not measured in this run") if traffic else None,
                         "launches_per_step": inst["launches"], "kernel_ms_sum": inst["ms"],
                         "kernel_ms_regime": "eager single-stream pass, one CUDA-event pair per launch; its own wall time is "
                                             f"{inst['wall_ms']:.1f} ms.  The graph-replayed step overlaps the face / body / part branches "
                                             "on forked streams, so this sum may exceed ms_per_step - it is kernel busy time, not a share of it",
                         "mma_per_fp32_product": MMA_PER_PRODUCT[args.precision],
                         "tensor_pipe_frac": achieved * MMA_PER_PRODUCT[args.precision] / peak,
                         "note": "achieved = algorithmic FLOP (2*rows*cout*cin*taps) / summed launch durations; "
                                 "tensor_pipe_frac counts the 1/3/6 MMAs issued per fp32 product",
                         "peak_source": f"{peaks['source']} bf16 sustained (MEASURED_PEAKS.json)",
                         "step_frac": FLOP_PER_FRAME * value / world / (peak * 1e12)},
            "roofline_vq": roofline_vq,
            "cpu_baseline": cpu_line,
            "extra": extra,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("PM_EMAGE_PRECISION", "fp16x3"), choices=list(DTYPES),
                    help="fp16x3 (default) and bf16x6 meet the fp32 parity gates; see DESIGN.md section 4")
    ap.add_argument("--cpu-baseline", type=int, default=1, help="0 skips the CPU timing (exploratory runs)")
    ap.add_argument("--extra", type=int, default=1, help="0 skips the other BASELINE configs (bs 1, CaMN, DisCo)")
    ap.add_argument("--body-priority", type=int, default=1, help="capture the critical (body) chain on a high-priority stream")
    ap.add_argument("--graph", type=int, default=1, help="replay the step as one CUDA graph (1) or launch eagerly (0)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
