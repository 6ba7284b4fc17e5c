"""Development aid (GPU): per-stage times of the batch decoder on one corpus for several kernel variants.
   python scripts/dev_decode_bench.py [--kind logs] [--frames 16384] [--level 3] [--steps 5] [--variants "KZG_EXEC=1;KZG_EXEC=2,KZG_EXEC_TILE=8192"]
Every variant is a ';'-separated list of comma-separated NAME=VALUE environment settings (read by the library at
every launch).  Parity is checked byte for byte against the generator's data for every variant."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import corpus
from kompressor_b200.batch import BatchDecoder

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="logs")
ap.add_argument("--frames", type=int, default=16384)
ap.add_argument("--frame-size", type=int, default=65536)
ap.add_argument("--level", type=int, default=3)
ap.add_argument("--checksum", action="store_true")
ap.add_argument("--streamed", action="store_true")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--variants", default="KZG_EXEC=1;KZG_EXEC=2,KZG_EXEC_TILE=8192")
args = ap.parse_args()
dev = torch.device("cuda", 0)
t = time.time()
c = corpus.build(args.kind, args.frames, args.frame_size, args.level, args.checksum, streamed=args.streamed, keep_raw=True)
print(f"corpus {args.kind} x{args.frames} of {args.frame_size} l{args.level}: ratio {c.ratio:.2f}, built in {time.time()-t:.1f}s", flush=True)
n = c.n_frames
dec = BatchDecoder(0)
dec.set_profiling(True)
caps = np.full(n, args.frame_size, np.uint64)
batch = dec.upload(c.comp, c.offsets, c.sizes, caps, device=dev)
for var in args.variants.split(";"):
    for kv in var.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
    batch.dst.zero_()
    dec.run(batch)
    ok = bool((batch.status == 0).all()) and bool((batch.out_size == args.frame_size).all())
    got = batch.dst[:n * args.frame_size].cpu().numpy()
    ok = ok and bool(np.array_equal(got, c.raw))
    if not ok:
        bad = [i for i in range(n) if int(batch.status[i]) != 0 or not np.array_equal(got[i*args.frame_size:(i+1)*args.frame_size], c.raw[i*args.frame_size:(i+1)*args.frame_size])]
        print("  PARITY FAILED frames", bad[:10], len(bad), "status", [int(batch.status[i]) for i in bad[:10]], flush=True)
    for _ in range(2):
        dec.enqueue(batch); dec.finish()
    acc = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(args.steps):
        ev0.record(torch.cuda.current_stream(dev)); dec.enqueue(batch); ev1.record(torch.cuda.current_stream(dev)); dec.finish()
        tot += ev0.elapsed_time(ev1)
        for name, ms in dec.stage_ms().items(): acc[name] = acc.get(name, 0.0) + ms
    ms = tot / args.steps
    st = {k: round(v / args.steps, 4) for k, v in acc.items()}
    print(json.dumps({"variant": var, "parity": ok, "ms_per_step": round(ms, 4), "GBps_out": round(n * args.frame_size / ms / 1e6, 1), "stage_ms": st}), flush=True)
    for kv in var.split(","):
        if "=" in kv: os.environ.pop(kv.split("=")[0], None)
