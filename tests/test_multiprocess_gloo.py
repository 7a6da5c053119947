"""The N > 1 host logic on CPU: 2 processes over gloo, each runs its point-range shard of an MSM and its share of
an NTT batch through the kernel-emulation build, partials are all-gathered and folded; result == oracle."""
import os
import socket
import subprocess
import sys

import helpers as H

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["BBG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["BBG_ROOT"], "tests"))
import torch.distributed as dist
import barretenberg_b200 as bb
from barretenberg_b200 import parallel
import helpers as H
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = bb.Library(os.path.join(os.environ["BBG_ROOT"], "tests", "emul", "libbbgpu_emul.so"))
n = 600
table, a0, d = H.generator_multiples_table(5, n)
sc = H.random_scalars_mont(6, n)
got = parallel.sharded_msm(lib, sc, table, rank, world)
assert (got == H.oracle_msm(sc, table)).all(), "sharded MSM mismatch on rank %d" % rank
# NTT batch sharded by polynomial, no collective
batch, m = 5, 1 << 12
od = H.OracleDomain(m)
mine = parallel.shard_batch(batch, rank, world)
assert sorted(sum((parallel.shard_batch(batch, r, world) for r in range(world)), [])) == list(range(batch))
for i in mine:
    x = H.random_scalars_mont(100 + i, m)
    assert (lib.ntt("coset_fft", x.copy()) == od.ntt(H.NTT_OPS["coset_fft"], x)).all()
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
'''


def test_two_rank_sharded_msm_and_ntt(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, BBG_ROOT=H.ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2", OMP_NUM_THREADS="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-2000:])
        assert "rank %d ok" % r in o
