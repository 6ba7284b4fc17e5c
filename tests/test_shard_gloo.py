"""N > 1 path on CPU: two gloo ranks shard a frame table, 'decode' their shard (here: the oracle, as
the checker of the host logic only) and gather the per-frame status / size words."""
import os
import subprocess
import sys

import numpy as np

from kompressor_b200.shard import partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    comp = rng.integers(1000, 20000, 1000)
    out = np.full(1000, 65536)
    for w in (1, 2, 3, 4, 8):
        r = partition(comp, out, w)
        assert r[0][0] == 0 and r[-1][1] == 1000 and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        loads = [float((comp[lo:hi] + out[lo:hi]).sum()) for lo, hi in r]
        assert max(loads) - min(loads) <= 2 * (65536 + 20000)
    assert partition([], [], 4) == [(0, 0)] * 4
    assert partition([5], [5], 4)[-1] == (1, 1) or sum(hi - lo for lo, hi in partition([5], [5], 4)) == 1


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["KZG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["KZG_ROOT"], "tests"))
import numpy as np, torch.distributed as dist
import corpus, oracle
from kompressor_b200.shard import partition, gather_frame_results
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
c = corpus.build("json", 12, 65536, 3, True, threads=1)
ranges = partition(c.sizes, np.full(c.n_frames, 65536), world)
lo, hi = ranges[rank]
st, sz = [], []
for i in range(lo, hi):
    out, err = oracle.decompress(bytes(c.comp[int(c.offsets[i]):int(c.offsets[i] + c.sizes[i])]), 65536)
    st.append(err); sz.append(0 if err else len(out))
status, sizes = gather_frame_results(st, sz, ranges, rank, world)
assert (status == 0).all() and (sizes == 65536).all() and len(status) == 12, (status, sizes)
if rank == 0: print("GATHER_OK", ranges)
dist.destroy_process_group()
'''


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, KZG_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK" in r.stdout
