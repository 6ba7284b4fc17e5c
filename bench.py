#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 zstd decode path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F]

A "step" is one pass of the hot path (scan -> parse -> link -> huffman -> fse -> offsets -> execute ->
finish) over one batch of synthetic input already resident in HBM.  Workload at every N: BASELINE.json
configs[1], 1 GiB of synthetic log-like text per GPU cut into 16384 independent 64 KiB frames, level 3
(frames compressed by the real libzstd 1.5.7).  Weak scaling: each rank decodes its own 1 GiB shard
(frames [rank*F, (rank+1)*F) of the corpus), no collective on the data path.

`--impl reference` times the reference's own CPU implementation of the path: libzstd (the library
behind Kompressor's ZSTD_decompressStream call) on the host cores through oracle/_ref, with all host
threads, on a bounded sample of the same workload.
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

METRIC = "zstd decode GB/s (output) at 1/2/4/8 B200, 64 KiB level-3 frames; % HBM peak"
FRAME = 65536
WORKLOAD = "configs[1]: 1 GiB synthetic log-like text per GPU, 16384 independent 64 KiB frames, zstd level 3"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                smax.append(float(f[2]))
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active") and t0 - 0.05 <= ts <= t1 + 0.15:
                    reasons.add(name)
        if not sm:
            sm = [float(r[1].split(",")[1]) for r in self.rows[-3:] if len(r[1].split(",")) > 2] or [0.0]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_corpus(rank: int, frames: int):
    import corpus
    t = time.time()
    c = corpus.build("logs", frames, FRAME, 3, False, first_frame=rank * frames, keep_raw=True)
    return c, time.time() - t


def cpu_reference_run(c, sample_frames: int, threads: int, passes: int):
    """libzstd on the host cores over the first `sample_frames` frames (oracle/_ref, kind 'reference')."""
    import numpy as np
    import corpus as corpus_mod
    import oracle
    oracle.build()
    n = min(sample_frames, c.n_frames)
    end = int(c.offsets[n - 1] + c.sizes[n - 1])
    sub = corpus_mod.Corpus(c.kind, c.frame_size, n, c.comp[:end], c.offsets[:n], c.sizes[:n], None, c.level, c.checksum, c.streamed)
    a = oracle.cpu_baseline(sub, threads=threads, passes=passes, mode=0, prefer="fast")
    b = oracle.cpu_baseline(sub, threads=1, passes=2, mode=1, prefer="fast")
    return a, b, n


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    frames = args.frames
    sample = min(frames, 4096)                      # 256 MiB of output per step: bounded sample
    c, gen_s = build_corpus(0, sample)
    threads = os.cpu_count() or 1
    a, b, n = cpu_reference_run(c, sample, threads, args.warmup + args.steps)
    timed = a["seconds"][args.warmup:] or a["seconds"]
    t_step = sum(timed) / len(timed)
    value = n * FRAME / t_step / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"first {n} frames ({n * FRAME >> 20} MiB out) per step, host memory"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": "reference",
                         "sample": f"libzstd {a['version']} ({a['path']}), parallel one-shot ZSTD_decompressDCtx per frame, "
                                   f"{threads} threads, {n} frames of 64 KiB; reference call pattern (ZSTD_decompressStream, "
                                   f"slices of max(8192,n/10)) on 1 thread: {b['gbps']:.3f} GB/s"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=16384, help="frames per GPU (16384 = 1 GiB of output)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the kzg path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from kompressor_b200 import build as kb
    from kompressor_b200.batch import BatchDecoder
    kb.build()
    c, gen_s = build_corpus(rank, args.frames)
    n = c.n_frames
    dec = BatchDecoder(local)
    dec.set_profiling(True)
    caps = np.full(n, FRAME, np.uint64)
    batch = dec.upload(c.comp, c.offsets, c.sizes, caps, device=dev)
    launches_per_step = dec.run(batch)                     # first pass sizes the workspace
    # correctness gate outside the timed region: byte for byte against the generator's data
    ok = bool((batch.status == 0).all()) and bool((batch.out_size == FRAME).all())
    got = batch.dst[:n * FRAME].cpu().numpy()
    ok = ok and bool(np.array_equal(got, c.raw))
    del got

    stream = torch.cuda.current_stream(dev)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        dec.enqueue(batch)
        dec.finish()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    stage_acc: dict = {}
    barrier()
    t0 = time.time()
    ev[0].record(stream)
    for k in range(args.steps):
        dec.enqueue(batch)
        ev[k + 1].record(stream)
        rc = dec.finish()          # host sync per step: the per-frame status is part of the step's result
        if rc != 0:
            ok = False
        for name, ms in dec.stage_ms().items():
            stage_acc[name] = stage_acc.get(name, 0.0) + ms
    barrier()
    t1 = time.time()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    clocks = sampler.stop(t0, t1)
    ok = ok and bool((batch.status == 0).all())

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    out_bytes_all = n * FRAME * world
    value = out_bytes_all / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the host-pointer C ABI (pinned host buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        src_pin = torch.from_numpy(c.comp).pin_memory()
        dst_pin = torch.empty(n * FRAME, dtype=torch.uint8).pin_memory()
        doff = np.arange(n, dtype=np.uint64) * FRAME
        offs = np.ascontiguousarray(c.offsets, dtype=np.uint64)
        sizes = np.ascontiguousarray(c.sizes, dtype=np.uint64)
        srcv, dstv = src_pin.numpy(), dst_pin.numpy()
        for _ in range(2):
            dec.decompress_host(srcv, offs, sizes, caps, out=dstv, out_offsets=doff)
        barrier()
        e2e_steps = max(2, min(args.steps, 5))
        te = time.perf_counter()
        for _ in range(e2e_steps):
            _, _, st_h, sz_h = dec.decompress_host(srcv, offs, sizes, caps, out=dstv, out_offsets=doff)
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - te) / e2e_steps
        ok = ok and bool((st_h == 0).all()) and bool(np.array_equal(dstv, c.raw))
        te_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
        e2e = {"value": out_bytes_all / float(te_t.item()) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(c.comp_bytes + n * 32), "d2h_bytes_per_step": int(n * FRAME + n * 12),
               "api": "kzg_decompress_batch_host (C ABI), pinned host buffers; H2D, decode and D2H pipelined over 128 MiB sub-batches"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (live CUDA-event durations recorded by the library on its stream)
    peak, peak_src = measured_peak_gbs()
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
    dom = max(stage_ms, key=stage_ms.get) if stage_ms else None
    alg_bytes = n * FRAME + c.comp_bytes                    # per launch: every compressed byte read once, every output byte written once
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if dom and os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom)
        except Exception:
            traffic = None
    roofline = None
    if dom:
        ach = alg_bytes / (stage_ms[dom] * 1e-3) / 1e9
        kname = {"execute": "kzg_exec_kernel", "huffman": "kzg_huf_kernel"}.get(dom, f"kzg_{dom}_kernel")
        roofline = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": stage_ms[dom],
                    "stage_ms": stage_ms,
                    "pipeline_frac_of_peak": (alg_bytes / (ms_per_step * 1e-3) / 1e9) / peak}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        a, b, ns = cpu_reference_run(c, 2048, threads, 4)
        cpu = {"value": a["gbps"], "unit": "GB/s", "cores": threads, "kind": "reference",
               "sample": f"libzstd {a['version']} via oracle/_ref, parallel one-shot per frame on {threads} threads, first {ns} "
                         f"frames (128 MiB out), best of 3; reference call pattern (ZSTD_decompressStream with "
                         f"max(8192,n/10) slices) on 1 thread: {b['gbps']:.3f} GB/s"}

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": n, "frame_bytes": FRAME, "ratio": round(c.ratio, 3),
                   "compressed_bytes_per_gpu": c.comp_bytes, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                   "l2": "inputs larger than L2 (0.22 GiB in + 1 GiB out per step vs 126 MB L2), no flush needed",
                   "parity": "byte-for-byte against the generator's data outside the timed region: " + ("ok" if ok else "FAILED"),
                   "corpus_build_s": round(gen_s, 1)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 2


if __name__ == "__main__":
    sys.exit(main())
