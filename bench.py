#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the mel + encoder hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--wtype f16|q8_0|q4_0] [--windows B | --total-windows W] [--impl reference]

A step = one pass of the hot path (PCM -> log-mel -> conv stem -> 32 encoder blocks -> pool -> LN) over one batch of
B synthetic 30 s windows per GPU (default B = 64, F16: BASELINE.json configs[1]).  One process per GPU; for N > 1 launch
under torchrun (RANK / LOCAL_RANK / WORLD_SIZE); windows are sharded, weights replicated, no collective on the path.

  value        whole-job audio-s/s with the PCM already resident in HBM; CUDA events on the library's stream, max over ranks
  e2e          the same metric through the reference-facing call (whisper_encode_batch) with pinned HOST buffers:
               H2D of the PCM and D2H of the embeddings inside the timed region
  roofline     the dominant kernel (tcgen05 weight GEMM): algorithmic FLOPs / CUDA-event time, measured live in the timed steps
  parity       what was timed is checked: window 0 of every batch is the golden clip on the golden model; after the B = 1 loop, after the
               last device-resident step and on the host buffer of the last e2e step its rows are compared with tests/golden/full_<wtype>.npz
               (the reference's own output); a mismatch exits non-zero
  configs      the other weight types / BASELINE configurations as short ride-along passes on the same GPUs: Q8_0 and Q4_0 weak passes
               (64 windows per GPU), Q8_0 / Q4_0 256 windows and F16 120 windows (1 h of audio) in total, sharded over the ranks
  cpu_baseline the UNMODIFIED reference (oracle/_ref, ggml CPU backend) on this box's host cores, one window (rank 0, N = 1)
  --impl reference   times only that reference arm and prints the same line with "impl": "reference"
"""
from __future__ import annotations

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

import numpy as np  # noqa: E402

METRIC = "audio_seconds_per_second_mel_plus_encoder"
UNIT = "audio-s/s"
WINDOW_S = 30.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), burst=d["bf16_tflops"], src="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, burst=1590.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        busy = [s for s, p in zip(sm, power) if p > 300] or sm
        return dict(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm), power_w_max=max(power))


def build_model_bytes(wtype_name: str, hp=None) -> bytes:
    from qwen2_audio_whisper_ggml_b200 import ggml_quant as gq, modelfile as mfm, synth
    wt = {"f16": gq.GGML_TYPE_F16, "q8_0": gq.GGML_TYPE_Q8_0, "q4_0": gq.GGML_TYPE_Q4_0, "f32": gq.GGML_TYPE_F32}[wtype_name]
    # (synth caches the seed-1234 F32 weights: they are drawn once per process and converted to each weight type asked for)
    return mfm.to_bytes(synth.synth_model(hp or synth.FULL_HPARAMS, wt, seed=1234))


def synth_windows(B: int, rank: int) -> np.ndarray:
    from qwen2_audio_whisper_ggml_b200 import synth
    base = [synth.synth_pcm(480000, seed=1000 * rank + i, kind="chirp" if i % 4 else "noise") for i in range(min(B, 8))]
    out = np.empty((B, 480000), dtype=np.float32)
    for b in range(B):
        out[b] = np.roll(base[b % len(base)], 997 * (b // len(base)))
    out[0] = synth.synth_pcm(480000, seed=0)      # the golden clip: bench checks what it timed (golden_check)
    return out


GOLDEN_TOL = {"f16": (2e-3, 1e-2), "q8_0": (3e-2, 0.15), "q4_0": (3e-2, 0.15)}   # (rel-L2, max-abs): tests/util.py TOL, SURVEY Appendix F


def golden_check(emb0: np.ndarray, wtype: str) -> dict:
    """window 0 of every benchmarked batch is the golden clip (chirp, seed 0) on the golden model (seed 1234): compare the rows the
    committed fixture holds (tests/golden/full_<wtype>.npz, produced by the unmodified reference's ggml CPU backend)"""
    path = os.path.join(ROOT, "tests", "golden", f"full_{wtype}.npz")
    if not os.path.exists(path):
        return {"checked": False, "why": f"{os.path.relpath(path, ROOT)} missing"}
    g = np.load(path)
    rows = emb0[g["emb_row_idx"]].astype(np.float64)
    want = g["emb_rows"].astype(np.float64)
    rel = float(np.linalg.norm(rows - want) / np.linalg.norm(want))
    mx = float(np.abs(rows - want).max())
    col = float(np.abs(emb0.mean(axis=0) - g["emb_col_mean"]).max())
    tol_rel, tol_abs = GOLDEN_TOL[wtype]
    return {"checked": True, "rel_l2": rel, "max_abs": mx, "col_mean_max_abs": col, "tol_rel_l2": tol_rel, "tol_max_abs": tol_abs,
            "ok": bool(rel < tol_rel and mx < tol_abs and col < tol_abs and np.isfinite(emb0).all()),
            "against": f"tests/golden/full_{wtype}.npz ({len(g['emb_row_idx'])} rows + column means of window 0; reference = ggml CPU backend)"}


def gemm_traffic(B: int):
    """DRAM bytes per GEMM launch from the committed ncu --set full captures of the four per-layer shapes at M = 96000 (B = 64),
    next to the algorithmic bytes of the same launches; None when the summary is missing or the batch differs"""
    path = os.path.join(ROOT, "profiles", "r02_gemm_ncu_summary.json")
    if B != 64 or not os.path.exists(path):
        return None, None, None
    d = json.load(open(path))
    meas, alg = [], []
    M = 96000
    shapes = {"qkv": (3840, 1280, 2, 1), "outproj": (1280, 1280, 4, 2), "fc1": (5120, 1280, 2, 1), "fc2": (1280, 5120, 4, 2)}   # N, K, out bytes, out passes (RMW)
    for nm, (N, K, ob, passes) in shapes.items():
        e = d.get("shapes", {}).get(nm)
        if not e:
            return None, None, None
        meas.append(e["dram_bytes"])
        alg.append(2.0 * M * K + 2.0 * N * K + passes * ob * M * N)
    return sum(meas) / len(meas), sum(alg) / len(alg), os.path.relpath(path, ROOT)


def run_reference(model_bytes: bytes, steps: int, warmup: int, threads: int):
    """the unmodified reference (ggml CPU backend) on host cores: one 30 s window per step"""
    from oracle import refbind
    from qwen2_audio_whisper_ggml_b200 import synth
    ctx = refbind.RefContext(model_bytes)
    pcm = synth.synth_pcm(480000, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        rc = ctx.full(pcm, n_threads=threads)
        dt = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError(f"reference whisper_full -> {rc}")
        if i >= warmup:
            times.append(dt)
    ctx.free()
    return times


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--wtype", default="f16", choices=["f16", "q8_0", "q4_0"])
    ap.add_argument("--windows", type=int, default=64, help="30 s windows per GPU per step")
    ap.add_argument("--total-windows", type=int, default=0, help="strong scaling: this many windows in total, sharded over the ranks "
                    "(BASELINE configs[2]/[3]: 256; configs[4], 1 h of audio: 120)")
    ap.add_argument("--max-batch", type=int, default=0, help="windows per micro-batch (default: = --windows)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the B = 1 latency leg (ncu launch lists of the throughput step)")
    ap.add_argument("--no-second-wtype", action="store_true", help="skip the short pass with the other weight type (Q8_0 next to F16)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = a.cpu_threads or (os.cpu_count() or 1)
    scaling = "weak"
    if a.total_windows:
        from qwen2_audio_whisper_ggml_b200.parallel import shard_bounds
        lo, hi = shard_bounds(a.total_windows, rank, world)
        a.windows = hi - lo
        scaling = "strong"
    if a.total_windows:
        workload = (f"BASELINE configs[2]-[4] shape: Qwen2-Audio encoder {a.wtype.upper()} (32L, d=1280, 20 heads, 128 mel), {a.total_windows} x 30 s windows "
                    f"({a.total_windows * 30 / 3600:.2f} h of audio) sharded over {world} GPU(s), mel + encoder")
    else:
        workload = (f"BASELINE configs[1]: Qwen2-Audio encoder {a.wtype.upper()} (32L, d=1280, 20 heads, 128 mel), "
                    f"{a.windows} x 30 s windows per GPU, mel + encoder")
    config = {"workload": workload, "windows_per_gpu": a.windows, "weights": a.wtype, "sharding": f"dp{world} (independent windows, replicated weights, no collective)",
              "total_windows": a.total_windows or a.windows * world,
              "l2": "no explicit flush: each step streams ~2.7 GB of activations + 1.26 GB of weights, >> 126 MB L2"}

    # ------------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        try:
            mb = build_model_bytes(a.wtype)
            times = run_reference(mb, a.steps, a.warmup, threads)
        except Exception as ex:   # oracle/_ref not shipped / wrong ISA: say so in one line, exit 0 (driver contract)
            print(json.dumps({"impl": "reference", "unavailable": f"{type(ex).__name__}: {ex}"[:300]}))
            return
        sec = statistics.median(times)
        val = WINDOW_S / sec
        sample = f"1 x 30 s window per step ({a.steps} steps, p50), {a.wtype} weights, n_threads={threads}"
        print(json.dumps({"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f16" if a.wtype == "f16" else f"{a.wtype} weights -> f16 / q8_0 activations (ggml vec_dot)",
                          "data": "synthetic", "impl": "reference", "config": config,
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from qwen2_audio_whisper_ggml_b200 import api
    from qwen2_audio_whisper_ggml_b200 import lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product arm; use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to stdout (NCCL_DEBUG is set on the GPU boxes), so
        # fd 1 points at stderr while the communicator is created
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    L.load_library()
    api.log_set(lambda lvl, txt: None)

    env = dict(rank=rank, local_rank=local_rank, world=world, threads=threads)
    # BASELINE configs[2]-[4] ride along as short strong-scaling passes (fixed total number of windows, sharded over the ranks), so
    # that one default run -- and the driver's 1/2/4/8-GPU runs of it -- records every configuration the metric names:
    #   F16 context:  + 120 windows (1 h of audio, configs[4]);   Q8_0 context: weak pass + 256 windows (configs[2]);   Q4_0: 256 (configs[3])
    full = not a.no_second_wtype and not a.total_windows
    main_run = measure(a.wtype, a.windows, a.steps, a.warmup, a.max_batch, env, latency=not a.no_latency, total_windows=a.total_windows,
                       extra_totals=(120,) if full and a.wtype == "f16" else ())
    others = []
    if full:
        for other, totals in (("q8_0", (256,)), ("q4_0", (256,))) if a.wtype == "f16" else (("f16", ()),):
            others.append(measure(other, a.windows, max(2, min(a.steps, 3)), 2, a.max_batch, env, latency=True, total_windows=0, extra_totals=totals))

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            times = run_reference(build_model_bytes(a.wtype), 1, 1, threads)
            cpu_baseline = {"value": WINDOW_S / times[0], "unit": UNIT, "cores": threads, "kind": "reference",
                            "sample": f"1 x 30 s window (1/{a.windows} of a step) after 1 warm-up window, {a.wtype} weights, unmodified reference ggml CPU backend (oracle/_ref), n_threads={threads}",
                            "seconds_per_window": times[0]}
        except Exception as ex:  # the checker is optional for the product arm
            cpu_baseline = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"unavailable: {ex}"}

    ok = True
    if rank == 0:
        line = assemble(main_run, a, world, scaling, config)
        line["cpu_baseline"] = cpu_baseline
        cfgs = {}
        for r_ in [main_run] + others:
            if r_ is not main_run:
                s2 = assemble(r_, a, world, scaling, dict(config, weights=r_["wtype"]))
                cfgs[f"{r_['wtype']}_weak_{a.windows}_per_gpu"] = {k: s2[k] for k in ("value", "unit", "ms_per_step", "dtype", "e2e", "roofline", "kernels", "parity",
                                                                                     "p50_ms_per_window_b1", "steps", "warmup", "gpu_launches")}
            for T, x in r_["extras"].items():
                cfgs[f"{r_['wtype']}_total_{T}_windows"] = {
                    "workload": f"{T} x 30 s windows ({T * 30 / 3600:.2f} h of audio) in total, {r_['wtype'].upper()} weights, sharded over {world} GPU(s) (strong scaling)",
                    "value": x["value"], "unit": UNIT, "ms_per_pass": x["ms_per_pass"], "steps": x["steps"], "scaling": "strong",
                    "e2e": {"value": x["e2e_value"], "unit": UNIT, "ms_per_pass": x["e2e_ms_per_pass"], "h2d_bytes_per_step": x["h2d"], "d2h_bytes_per_step": x["d2h"]},
                    "windows_rank0": x["windows_this_rank"], "parity": x["parity"]}
        if cfgs:
            line["configs"] = cfgs
        ok = all(p_.get("ok", True) for r_ in [main_run] + others for p_ in r_["parity"].values() if p_.get("checked"))
        line["parity_ok"] = ok
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: the benchmarked outputs do NOT match the reference golden (see \"parity\" in the line above)")


def measure(wtype: str, B: int, steps: int, warmup: int, max_batch: int, env: dict, latency: bool, total_windows: int, extra_totals=()) -> dict:
    """one weight type through every leg: B = 1 latency, device-resident throughput with the live per-kernel profile, end to end
    with host buffers -- and a golden check of window 0 after each of them"""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from qwen2_audio_whisper_ggml_b200 import Context, api
    from qwen2_audio_whisper_ggml_b200 import lib as L
    rank, local_rank, world = env["rank"], env["local_rank"], env["world"]
    lib = L.load_library()
    t_setup = time.time()
    mb = build_model_bytes(wtype)
    cp = api.default_context_params()
    cp.gpu_device = local_rank
    ctx = Context.init_from_buffer(mb, cp)
    del mb
    assert ctx.set_max_batch(max_batch or B) == 0
    st = ctx.q2w_state()
    stream = torch.cuda.ExternalStream(lib.q2w_state_stream(st), device=torch.device("cuda", local_rank))
    host = torch.from_numpy(synth_windows(B, rank)).pin_memory()
    dev = host.cuda()
    outs = [torch.empty((B, 750, 1280), dtype=torch.float32).pin_memory() for _ in range(2)]
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup
    parity = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        rc = ctx.encode_batch_device(dev.data_ptr(), 480000, B)
        if rc != 0:
            raise RuntimeError("whisper_encode_batch_device failed")

    def window0_device():
        e = np.empty((750, 1280), dtype=np.float32)
        L.check(lib.q2w_get_embeddings(st, e.ctypes.data, 0, e.size))
        return e

    def run_host_steps(n):
        """n end-to-end steps through whisper_encode_batch_async / _wait: every step copies its PCM from pinned host memory and
        lands its embeddings in pinned host memory; step i + 1 is queued before step i is waited for (two in flight), which is
        how a serving loop uses the call. Q2W_BENCH_E2E_SYNC=1 falls back to one blocking whisper_encode_batch per step."""
        if os.environ.get("Q2W_BENCH_E2E_SYNC") == "1":
            for i in range(n):
                ctx.encode_batch(host.numpy(), out=outs[i & 1].numpy())
            return
        prev = None
        for i in range(n):
            t = ctx.encode_batch_async(host.numpy(), outs[i & 1].numpy())
            if prev is not None:
                ctx.wait(prev)
            prev = t
        ctx.wait(prev)

    # ---- single-window latency (p50 ms per 30 s window, B = 1), device-resident PCM. Measured FIRST: it is clock-bound, and right
    #      after the power-capped throughput phase the SM clock is still held down
    lat = []
    if latency:
        for i in range(23):
            barrier() if i == 0 else None
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            ctx.encode_batch_device(dev.data_ptr(), 480000, 1)
            a1.record(stream)
            torch.cuda.synchronize()
            if i >= 3:
                lat.append(a0.elapsed_time(a1))
        parity["b1_graph_replay"] = golden_check(window0_device(), wtype)

    # ---- device-resident throughput ("value") with the live per-kernel roofline
    for _ in range(warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    L.check(lib.q2w_profile_enable(st, 1))
    launches0 = lib.q2w_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step_device()
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = lib.q2w_kernel_launches() - launches0
    clocks = sampler.stop()
    parity["device_batch"] = golden_check(window0_device(), wtype)          # window 0 of the LAST timed step

    prof = {}
    names = ["gemm", "attention", "layernorm", "mel", "im2col", "dequant"]
    for ci, nm in enumerate(names):
        ms, cnt, fl, by = C.c_double(), C.c_long(), C.c_double(), C.c_double()
        L.check(lib.q2w_profile_read(st, ci, C.byref(ms), C.byref(cnt), C.byref(fl), C.byref(by)))
        prof[nm] = dict(ms=ms.value, count=cnt.value, flops=fl.value, bytes=by.value)
    L.check(lib.q2w_profile_enable(st, 0))

    # ---- end-to-end through the reference-facing API with host buffers
    for o in outs:
        o.fill_(float("nan"))
    run_host_steps(max(1, min(warmup, 2)))
    barrier()
    t0 = time.perf_counter()
    run_host_steps(steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    parity["e2e_host_batch"] = golden_check(outs[(steps - 1) & 1][0].numpy(), wtype)   # window 0 as it landed in host memory, last timed step
    finite = bool(torch.isfinite(outs[(steps - 1) & 1]).all())
    parity["e2e_host_batch"]["all_windows_finite"] = finite
    parity["e2e_host_batch"]["ok"] = parity["e2e_host_batch"].get("ok", True) and finite

    # ---- optional: gather every rank's embeddings on rank 0 over NCCL (the only collective this path ever issues; NOT part of
    #      the timed region -- SURVEY 8(e): "NCCL only to gather embeddings when a caller asks for them on one device")
    gather_ms = None
    if world > 1:
        emb = outs[0].cuda()                        # this rank's embeddings of an end-to-end step
        parts = [torch.empty_like(emb) for _ in range(world)] if rank == 0 else None
        dist.gather(emb, parts, dst=0)              # first use builds NCCL's p2p channels; time the second
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.gather(emb, parts, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        if rank == 0:
            assert all(torch.isfinite(p_).all() for p_ in parts) and torch.equal(parts[0], emb)

    t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = float(t[0]), float(t[1])
    h2d, d2h = int(host.numel() * 4), int(outs[0].numel() * 4)

    # ---- the other BASELINE configurations on the same context: a fixed TOTAL number of windows sharded over the ranks (strong scaling;
    #      configs[2]/[3]: 256 windows, configs[4]: 1 h of audio = 120 windows), larger than max_batch -> chunked into micro-batches
    extras = {}
    for T in extra_totals:
        from qwen2_audio_whisper_ggml_b200.parallel import shard_bounds
        lo, hi = shard_bounds(T, rank, world)
        n = hi - lo
        reps = -(-n // B)
        xh = host.repeat(reps, 1)[:n].contiguous().pin_memory() if n > 0 else None
        xd = xh.cuda() if n > 0 else None
        xo = [torch.empty((max(n, 1), 750, 1280), dtype=torch.float32).pin_memory() for _ in range(2)]
        def xdev():
            if n > 0 and ctx.encode_batch_device(xd.data_ptr(), 480000, n) != 0:
                raise RuntimeError("whisper_encode_batch_device failed")
        def xhost(k):
            prev = None
            for i in range(k):
                if n == 0:
                    continue
                tk = ctx.encode_batch_async(xh.numpy(), xo[i & 1][:n].numpy())
                if prev is not None:
                    ctx.wait(prev)
                prev = tk
            if prev is not None:
                ctx.wait(prev)
        xsteps = 3
        for _ in range(2):
            xdev()
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(stream)
        for _ in range(xsteps):
            xdev()
        x1.record(stream)
        barrier()
        x_ms = x0.elapsed_time(x1)
        par = golden_check(window0_device(), wtype) if n > 0 else {"checked": False, "why": "empty shard"}
        xhost(1)
        barrier()
        tx = time.perf_counter()
        xhost(xsteps)
        barrier()
        x_e2e = time.perf_counter() - tx
        tt = torch.tensor([x_ms, x_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        extras[T] = dict(total_windows=T, windows_this_rank=n, steps=xsteps, ms_per_pass=float(tt[0]) / xsteps,
                         value=WINDOW_S * T * xsteps / (float(tt[0]) / 1e3), e2e_value=WINDOW_S * T * xsteps / float(tt[1]),
                         e2e_ms_per_pass=1e3 * float(tt[1]) / xsteps, parity=par, h2d=int(n) * 480000 * 4, d2h=int(n) * 750 * 1280 * 4)
        parity[f"total{T}_device_batch"] = par
        del xh, xd, xo
    ctx.free()
    del dev, host, outs
    torch.cuda.empty_cache()
    return dict(wtype=wtype, B=B, steps=steps, warmup=warmup, dev_ms=dev_ms, e2e_s=e2e_s, launches=int(launches), clocks=clocks, prof=prof, lat=lat,
                parity=parity, gather_ms=gather_ms, setup_s=setup_s, h2d=h2d, d2h=d2h, total_windows=total_windows, extras=extras)


def assemble(r: dict, a, world: int, scaling: str, config: dict) -> dict:
    B, steps, dev_ms, e2e_s, prof, clocks = r["B"], r["steps"], r["dev_ms"], r["e2e_s"], r["prof"], r["clocks"]
    total_audio = WINDOW_S * (r["total_windows"] or B * world) * steps
    value = total_audio / (dev_ms / 1e3)
    e2e_value = total_audio / e2e_s
    pk = peaks()
    g = prof["gemm"]
    gemm_tflops = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    traffic, traffic_alg, traffic_src = gemm_traffic(B)
    roofline = {"bound": "tensor", "kernel": "gemm_kernel<EPI> (tcgen05.mma kind::f16, TMA, TMEM)", "achieved": gemm_tflops, "peak": pk["tflops"],
                "unit": "TFLOP/s", "frac": gemm_tflops / pk["tflops"], "peak_source": f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
                # dram__bytes_read + dram__bytes_write per launch, mean over the four per-layer GEMM shapes, read from the committed ncu
                # --set full captures (null until they exist for this batch size); traffic_algorithmic = operands + output of the same launches
                "traffic": traffic, "traffic_algorithmic": traffic_alg, "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
                "launches": g["count"], "avg_launch_ms": g["ms"] / max(1, g["count"]),
                "flops_per_launch": g["flops"] / max(1, g["count"]), "share_of_step": g["ms"] / dev_ms}
    kernels = {}
    for nm, p in prof.items():
        if p["count"] == 0:
            continue
        kernels[nm] = {"ms_per_step": p["ms"] / steps, "launches_per_step": p["count"] / steps, "share": p["ms"] / dev_ms}
        if p["flops"] > 0:
            kernels[nm]["tflops"] = p["flops"] / (p["ms"] * 1e-3) / 1e12
        if p["bytes"] > 0 and p["flops"] == 0:
            kernels[nm]["gbs"] = p["bytes"] / (p["ms"] * 1e-3) / 1e9
            kernels[nm]["hbm_frac"] = kernels[nm]["gbs"] / pk["hbm_gbs"]
    if "attention" in kernels and clocks.get("sm_mhz"):
        # at head dim 64 one exponential carries 4 * 64 tensor FLOPs and the SM retires 16 ex2 per clock (tools/ubench/xu_pipe.cu), so the
        # MUFU, not the tensor pipe, bounds this kernel: ceiling = 148 SMs * 16 * clock * 256 FLOP
        ceil = 148 * 16 * clocks["sm_mhz"] * 1e6 * 256 / 1e12
        kernels["attention"]["mufu_ceiling_tflops_at_sampled_clock"] = ceil
        kernels["attention"]["frac_of_mufu_ceiling"] = kernels["attention"]["tflops"] / ceil
    lat = r["lat"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": r["warmup"],
            "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            # the arithmetic the path computes in: F16 tensor-core operands with F32 accumulation for every weight type; Q8_0 / Q4_0
            # weights stay quantised in HBM and are decoded to F16 on the way into the GEMM
            "dtype": "f16" if r["wtype"] == "f16" else f"{r['wtype']} weights -> f16",
            "data": "synthetic", "config": dict(config, weights=r["wtype"]), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "ms_per_step": 1e3 * e2e_s / steps,
                    "api": "whisper_encode_batch (blocking)" if os.environ.get("Q2W_BENCH_E2E_SYNC") == "1" else "whisper_encode_batch_async + whisper_encode_batch_wait, two batches in flight"},
            "gpu_launches": r["launches"], "roofline": roofline, "kernels": kernels, "parity": r["parity"],
            "p50_ms_per_window_b1": statistics.median(lat) if lat else None,
            "p10_p90_ms_per_window_b1": [sorted(lat)[len(lat) // 10], sorted(lat)[(9 * len(lat)) // 10]] if lat else None,
            "ms_per_window": dev_ms / steps / B, "nccl_gather_ms": r["gather_ms"], "setup_s": r["setup_s"]}
    line["tflops_whole_step"] = 2.2738e12 * B * steps / (dev_ms / 1e3) / 1e12
    return line


if __name__ == "__main__":
    main()
