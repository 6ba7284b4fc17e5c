#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 zstd decode path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F]

A "step" is one pass of the hot path (scan -> parse -> link -> fsetab -> fse -> huffman -> offsets -> bulk ->
execute -> finish) over one batch of synthetic input already resident in HBM.  Headline workload at every N:
BASELINE.json configs[1], 1 GiB of synthetic log-like text per GPU cut into 16384 independent 64 KiB frames, level 3
(frames compressed by the real libzstd 1.5.7).  Weak scaling: each rank decodes its own 1 GiB shard (frames
[rank*F, (rank+1)*F) of the corpus), no collective on the data path.

Beside the headline the same line carries, under `config`:
  other_configs  (N = 1)  BASELINE.json configs[2], [3], [4] on one GPU: GB/s, ms, parity, fraction of the HBM
                          roofline, and libzstd on the host cores on the same frames;
  sharded_json   (N > 1)  BASELINE.json configs[2] as written: the 8 GiB JSON corpus with XXH64 checksums,
                          131072 frames, STRONG-sharded over the N ranks with kompressor_b200.shard.partition
                          (balanced by compressed + decompressed bytes), per-frame {status, size} gathered with
                          shard.gather_frame_results;
and under `e2e` the ceiling of the node's host path (`ceiling_gbs`: plain pinned cudaMemcpyAsync of the same
bytes in both directions on all ranks at once).

`--impl reference` times the reference's own CPU implementation of the path: libzstd (the library behind Kompressor's
ZSTD_decompressStream call, kompressor-zstd--nativelib/src/jvmCommonMain/jni/Wrapper.cpp:178) on the host cores through
oracle/_ref, with all host threads, on the same frames.  Both arms use ONE routine (cpu_reference_run) for the CPU
figure: all frames of the workload per step, mean over the timed steps.
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
JSON_TOTAL_FRAMES = 131072          # configs[2]: 8 GiB of JSON lines in 64 KiB frames


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


def cpu_reference_run(c, warmup: int, steps: int, threads: int | None = None, frames: int | None = None):
    """THE CPU routine of both arms: libzstd on the host cores (oracle/_ref, kind 'reference'), parallel one-shot
    ZSTD_decompressDCtx per frame, static contiguous partition over `threads` threads, over the first `frames` frames of
    corpus `c` (default: all of them).  One step = one pass over those frames.
    -> dict(gbps = output bytes / mean time of the timed steps, ms_per_step, frames, threads, version, path, pattern_gbps =
       the reference's own call pattern, ZSTD_decompressStream with max(8192, n/10) output slices, on one thread)."""
    import corpus as corpus_mod
    import oracle
    oracle.build()
    threads = threads or (os.cpu_count() or 1)
    n = min(frames or c.n_frames, c.n_frames)
    end = int(c.offsets[n - 1] + c.sizes[n - 1])
    sub = corpus_mod.Corpus(c.kind, c.frame_size, n, c.comp[:end], c.offsets[:n], c.sizes[:n], None, c.level, c.checksum, c.streamed)
    warmup = max(warmup, 1)
    a = oracle.cpu_baseline(sub, threads=threads, passes=warmup + steps, mode=0, prefer="fast")
    timed = a["seconds"][warmup:]
    t_step = sum(timed) / len(timed)
    # the reference's call pattern on a bounded part (one thread): a reported side figure
    m = min(n, max(1, (64 << 20) // c.frame_size))
    endm = int(c.offsets[m - 1] + c.sizes[m - 1])
    subm = corpus_mod.Corpus(c.kind, c.frame_size, m, c.comp[:endm], c.offsets[:m], c.sizes[:m], None, c.level, c.checksum, c.streamed)
    b = oracle.cpu_baseline(subm, threads=1, passes=2, mode=1, prefer="fast")
    return {"gbps": n * c.frame_size / t_step / 1e9, "ms_per_step": t_step * 1e3, "frames": n, "threads": threads,
            "version": a["version"], "path": a["path"], "pattern_gbps": b["gbps"]}


def cpu_sample_text(r, what):
    return (f"libzstd {r['version']} via oracle/_ref, parallel one-shot ZSTD_decompressDCtx per frame on {r['threads']} threads, "
            f"{what}, mean of the timed steps; reference call pattern (ZSTD_decompressStream with max(8192,n/10) slices) "
            f"on 1 thread: {r['pattern_gbps']:.3f} GB/s")


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    c, gen_s = build_corpus(0, args.frames)
    r = cpu_reference_run(c, args.warmup, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["gbps"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": c.n_frames, "frame_bytes": FRAME, "ratio": round(c.ratio, 3),
                   "compressed_bytes_per_gpu": c.comp_bytes,
                   "sample": f"all {r['frames']} frames ({r['frames'] * FRAME >> 20} MiB out) per step, host memory"},
        "cpu_baseline": {"value": r["gbps"], "unit": "GB/s", "cores": r["threads"], "kind": "reference",
                         "sample": cpu_sample_text(r, f"all {r['frames']} frames of 64 KiB")},
        "e2e": {"value": r["gbps"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_numa_node(local: int):
    """Pins this process to the CPUs next to its GPU (so that pinned arenas are first-touched there). Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        if ids and node >= 0:
            os.sched_setaffinity(0, ids)
            return {"gpu": bdf, "numa_node": node, "cpus": cpus, "bound": True}
        return {"gpu": bdf, "numa_node": node, "cpus": cpus, "bound": False}
    except Exception as e:          # virtualised boxes often report no topology
        return {"bound": False, "why": str(e)[:80]}


def time_decode(dec, batch, dev, steps, warm):
    """Device-timed steps of one resident batch -> (ms_per_step, rc_ok)."""
    import torch
    stream = torch.cuda.current_stream(dev)
    for _ in range(warm):
        dec.enqueue(batch)
        dec.finish()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ok = True
    torch.cuda.synchronize(dev)
    e0.record(stream)
    for _ in range(steps):          # back to back, one host synchronisation at the end (as in the headline loop)
        dec.enqueue(batch)
    e1.record(stream)
    if dec.finish() != 0:
        ok = False
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps, ok


def other_config(dec, dev, name, kind, frames, frame_size, level, checksum, streamed, peak, cpu_frames):
    """One of BASELINE.json's other configs on this GPU: parity, device-timed GB/s, roofline fraction, CPU beside it."""
    import numpy as np
    import corpus
    t = time.time()
    c = corpus.build(kind, frames, frame_size, level, checksum, streamed=streamed, keep_raw=True)
    gen_s = time.time() - t
    n = c.n_frames
    caps = np.full(n, frame_size, np.uint64)
    batch = dec.upload(c.comp, c.offsets, c.sizes, caps, device=dev)
    dec.run(batch)
    ok = bool((batch.status == 0).all()) and bool((batch.out_size == frame_size).all())
    ok = ok and bool(np.array_equal(batch.dst[:n * frame_size].cpu().numpy(), c.raw))
    ms, ok2 = time_decode(dec, batch, dev, steps=3, warm=2)
    stage = {k: round(v, 4) for k, v in dec.stage_ms().items()}
    del batch
    r = cpu_reference_run(c, 1, 3, frames=cpu_frames)
    alg = n * frame_size + c.comp_bytes
    return {"config": name, "frames": n, "frame_bytes": frame_size, "level": level, "checksum": checksum, "streamed": streamed,
            "ratio": round(c.ratio, 3), "GBps": n * frame_size / (ms * 1e-3) / 1e9, "ms": ms, "parity": "ok" if ok and ok2 else "FAILED",
            "pipeline_frac_of_peak": alg / (ms * 1e-3) / 1e9 / peak, "stage_ms_last_step": stage,
            "cpu_gbs": r["gbps"], "cpu_cores": r["threads"], "cpu_frames": r["frames"], "corpus_build_s": round(gen_s, 1)}


def sharded_json(dec, dev, rank, world, peak):
    """BASELINE.json configs[2]: the 8 GiB JSON corpus, XXH64 checksums on, strong-sharded over the ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import corpus
    from kompressor_b200 import shard
    total = JSON_TOTAL_FRAMES
    margin = 128
    per = total // world
    g_lo, g_hi = max(0, rank * per - margin), min(total, (rank + 1) * per + margin if rank + 1 < world else total)
    t = time.time()
    c = corpus.build("json", g_hi - g_lo, FRAME, 3, True, first_frame=g_lo, keep_raw=True)
    gen_s = time.time() - t
    # the whole frame table's compressed sizes: every rank contributes its own equal-count slice
    core_lo, core_hi = rank * per, ((rank + 1) * per if rank + 1 < world else total)
    mine = torch.zeros(total - (world - 1) * per if world > 1 else total, dtype=torch.int64, device=dev)
    mine[:core_hi - core_lo] = torch.from_numpy(c.sizes[core_lo - g_lo:core_hi - g_lo].astype(np.int64)).to(dev)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    sizes_all = np.concatenate([parts[r][:(((r + 1) * per if r + 1 < world else total) - r * per)].cpu().numpy() for r in range(world)]).astype(np.uint64)
    ranges = shard.partition(sizes_all, np.full(total, FRAME, np.uint64), world)
    lo, hi = ranges[rank]
    balanced = lo >= g_lo and hi <= g_hi
    if not balanced:                # (never with this corpus: the byte-balanced cut is a few frames from the equal-count one)
        ranges = [(r * per, (r + 1) * per if r + 1 < world else total) for r in range(world)]
        lo, hi = ranges[rank]
    n = hi - lo
    k0 = lo - g_lo
    off = c.offsets[k0:k0 + n] - c.offsets[k0]
    comp = c.comp[int(c.offsets[k0]):int(c.offsets[k0 + n - 1] + c.sizes[k0 + n - 1])]
    caps = np.full(n, FRAME, np.uint64)
    batch = dec.upload(comp, off, c.sizes[k0:k0 + n], caps, device=dev)
    dec.run(batch)
    st, sz = shard.gather_frame_results(batch.status.cpu().numpy(), batch.out_size.cpu().numpy(), ranges, rank, world, device=dev)
    ok = bool((st == 0).all()) and bool((sz == FRAME).all())          # status 0 includes the XXH64 check of every frame on the device
    probe = np.linspace(0, n - 1, 64).astype(np.int64)                 # plus a byte-for-byte probe of 64 frames per rank
    got = batch.dst[:n * FRAME]
    for i in probe:
        ok = ok and bool(np.array_equal(got[int(i) * FRAME:(int(i) + 1) * FRAME].cpu().numpy(), c.raw[(k0 + int(i)) * FRAME:(k0 + int(i) + 1) * FRAME]))
    dist.barrier()
    ms, ok2 = time_decode(dec, batch, dev, steps=3, warm=2)
    tt = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    okt = torch.tensor([1 if ok and ok2 else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    ms = float(tt.item())
    comp_total = int(sizes_all.sum())
    del batch
    return {"config": "configs[2]: 8 GiB synthetic JSON lines, 131072 frames of 64 KiB, level 3, XXH64 checksums, strong-sharded",
            "scaling": "strong", "frames_total": total, "ranges": [list(r) for r in ranges],
            "partition": "shard.partition (compressed + decompressed bytes)" if balanced else "equal frame counts (fallback)",
            "GBps": total * FRAME / (ms * 1e-3) / 1e9, "ms": ms, "parity": "ok" if int(okt.item()) else "FAILED",
            "pipeline_frac_of_peak_per_gpu": (total * FRAME + comp_total) / world / (ms * 1e-3) / 1e9 / peak,
            "gather": "shard.gather_frame_results over NCCL: {status, size} of all 131072 frames on every rank",
            "corpus_build_s": round(gen_s, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=16384, help="frames per GPU (16384 = 1 GiB of output)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs[2..4] (other_configs / sharded_json)")
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
    numa = bind_to_gpu_numa_node(local) if world > 1 else {"bound": False, "why": "single rank"}
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
    # The steps are enqueued back to back on the stream, as a pipeline would, with one host synchronisation at the end
    # (a host round trip per step left the GPU idle for ~0.07 ms between steps: Python reading twelve events and
    # enqueueing again).  Every step still writes the per-frame status and sizes; they are checked after the last one.
    for k in range(args.steps):
        dec.enqueue(batch)
        ev[k + 1].record(stream)
    if dec.finish() != 0:          # workspace overflow: cannot happen after the sizing pass above
        ok = False
    barrier()
    t1 = time.time()
    total_ms = ev[0].elapsed_time(ev[args.steps])
    # per-stage device times of the timed steps: the library keeps the CUDA events of the last 32 profiled batches
    staged_steps = 0
    for age in range(min(args.steps, 32)):
        st = dec.stage_ms(age)
        if not st:
            break
        staged_steps += 1
        for name, ms in st.items():
            stage_acc[name] = stage_acc.get(name, 0.0) + ms
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
        # the ceiling of the host path: the same bytes as plain pinned copies, both directions at once, all ranks at once
        # (one cudaMemcpyAsync per direction; no decode)
        d_in = torch.empty(c.comp_bytes, dtype=torch.uint8, device=dev)
        d_out = batch.dst[:n * FRAME]
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ceil_s = []
        for rep in range(4):
            barrier()
            tc = time.perf_counter()
            with torch.cuda.stream(s_in):
                d_in.copy_(src_pin, non_blocking=True)
            with torch.cuda.stream(s_out):
                dst_pin.copy_(d_out, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
            if rep:
                ceil_s.append(time.perf_counter() - tc)
        tc_t = torch.tensor([min(ceil_s)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc_t, op=dist.ReduceOp.MAX)
        del d_in
        e2e = {"value": out_bytes_all / float(te_t.item()) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(c.comp_bytes + n * 32), "d2h_bytes_per_step": int(n * FRAME + n * 12),
               "ceiling_gbs": out_bytes_all / float(tc_t.item()) / 1e9,
               "ceiling": "same bytes as plain pinned cudaMemcpyAsync (H2D of the compressed bytes and D2H of the output concurrently, "
                          "no decode) on all ranks at once, best of 3, slowest rank",
               "numa": numa,
               "api": "kzg_decompress_batch_host (C ABI), pinned host buffers; H2D, decode and D2H pipelined over 128 MiB sub-batches"}

    # ---- BASELINE.json's other configs (outside the timed region of the headline)
    peak, peak_src = measured_peak_gbs()
    other, sharded = None, None
    if not args.no_other_configs:
        del batch
        torch.cuda.empty_cache()
        if world == 1:
            other = [
                other_config(dec, dev, "configs[2] on one GPU: 1 GiB of the JSON corpus, 64 KiB frames, level 3, XXH64 checksums", "json", 16384, FRAME, 3, True, False, peak, 4096),
                other_config(dec, dev, "configs[3]: literal-heavy mix, 64 KiB frames, level 1, 1 GiB", "literal", 16384, FRAME, 1, False, False, peak, 4096),
                other_config(dec, dev, "configs[4] on one GPU: 16 frames of 4 MiB (2 MiB logs + 2 MiB JSON), level 19, streamed, 8 MiB window", "mixed", 16, 4 << 20, 19, False, True, peak, 16),
                # the same frames, four times as many in the batch: the CTA-per-frame execute kernel is bound by the serial chains
                # inside ONE frame, so its time does not grow until every SM has a frame ("64 per GPU")
                other_config(dec, dev, "configs[4] on one GPU, 64 frames in the batch: 4 MiB each (2 MiB logs + 2 MiB JSON), level 19, streamed, 8 MiB window", "mixed", 64, 4 << 20, 19, False, True, peak, 64),
            ]
        else:
            sharded = sharded_json(dec, dev, rank, world, peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (live CUDA-event durations recorded by the library on its stream)
    stage_ms = {k: v / max(staged_steps, 1) for k, v in stage_acc.items()}
    dom = max(stage_ms, key=stage_ms.get) if stage_ms else None
    alg_bytes = n * FRAME + c.comp_bytes                    # per launch: every compressed byte read once, every output byte written once
    roofline = None
    if dom:
        ach = alg_bytes / (stage_ms[dom] * 1e-3) / 1e9
        kname = {"execute": "kzg_exec_kernel<false>", "huffman": "kzg_huf_kernel"}.get(dom, f"kzg_{dom}_kernel")
        # DRAM traffic of that kernel: from the committed ncu capture of this very command, if it is of this kernel and
        # of a build whose kernel time agrees within 15 %; otherwise null (a stale constant is worse than none)
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic_r02.json")
        try:
            tj = json.load(open(tp))
            e = tj.get(dom)
            if e and abs(e["kernel_ms_at_capture"] - stage_ms[dom]) <= 0.15 * stage_ms[dom]:
                traffic, traffic_src = e["dram_bytes_per_launch"], tj.get("_note")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": stage_ms[dom],
                    "stage_ms": stage_ms,
                    "pipeline_frac_of_peak": (alg_bytes / (ms_per_step * 1e-3) / 1e9) / peak,
                    "note": "frac divides the whole batch's algorithmic bytes by the dominant kernel's time alone; "
                            "pipeline_frac_of_peak (all kernels of a step) is the number that maps to the 50 % target"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(c, 1, 5)
        cpu = {"value": r["gbps"], "unit": "GB/s", "cores": r["threads"], "kind": "reference",
               "sample": cpu_sample_text(r, f"all {r['frames']} frames of 64 KiB (the same routine and sample as --impl reference)")}

    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": n, "frame_bytes": FRAME, "ratio": round(c.ratio, 3),
                   "compressed_bytes_per_gpu": c.comp_bytes, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                   "l2": "inputs larger than L2 (0.22 GiB in + 1 GiB out per step vs 126 MB L2), no flush needed",
                   "parity": "byte-for-byte against the generator's data outside the timed region: " + ("ok" if ok else "FAILED"),
                   "stepping": "steps enqueued back to back on one stream, one host synchronisation after the last; per-stage times from the library's CUDA events of each step",
                   "corpus_build_s": round(gen_s, 1), "other_configs": other, "sharded_json": sharded},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 2


if __name__ == "__main__":
    sys.exit(main())
