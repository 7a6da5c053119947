"""The N > 1 path on real GPUs: one process per GPU over NCCL (what bench.py --gpus N runs under torchrun), each rank runs
its point-range shard of an MSM and its share of an NTT batch on ITS device through the C ABI, the 128-byte partials are
all-gathered over NCCL and folded on the host; every rank must hold the oracle's result.  Also the registered-SRS form
(fixed-base windows per shard) and the device-resident partial path the bench step uses.
Needs >= 2 GPUs (skipped on a one-GPU box: the same logic runs there over gloo on the CPU, tests/test_multiprocess_gloo.py)."""
import os
import socket
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["BBG_ROOT"]); sys.path.insert(0, os.path.join(os.environ["BBG_ROOT"], "tests"))
import torch
import torch.distributed as dist
import barretenberg_b200 as bb
from barretenberg_b200 import parallel
from barretenberg_b200 import synthetic as S
import helpers as H
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = bb.Library(device=local)  # raises without libbbgpu.so or a GPU: no fallback
# 1. host-buffer MSM, sharded by point range, vs the plain-C oracle
n = 3000
table, a0, d = H.generator_multiples_table(5, n)
sc = H.random_scalars_mont(6, n)
got = parallel.sharded_msm(lib, sc, table, rank, world, device="cuda")
assert (got == H.oracle_msm(sc, table)).all(), "sharded MSM mismatch on rank %d" % rank
# 2. the bench step's form at 2^16: device-resident shard over a registered table (fixed-base windows), closed form
n = 1 << 16
lo, hi = parallel.shard_range(n, rank, world)
table, a0, d = H.generator_multiples_table(7, n)
sc = H.random_scalars_mont(8, n)
lib.set_srs_precompute(True)
keep = lib.srs_register(np.ascontiguousarray(table[2 * lo:2 * hi]))
d_tab, c, w = lib.srs_device_table(keep)
assert w >= 6
d_sc = lib.dev_alloc((hi - lo) * 32)
lib.h2d(d_sc, np.ascontiguousarray(sc[lo:hi]))
part = lib.msm_partial_dev(d_sc, d_tab, hi - lo)
got = lib.fold_partials(parallel.gather_partials(part, world, device="cuda"))
assert (got == H.closed_form_msm(sc, a0, d)).all(), "device-resident sharded MSM mismatch on rank %d" % rank
lib.srs_unregister(keep)
# 3. NTT batch sharded by polynomial, no collective
batch, m = 5, 1 << 13
od = H.OracleDomain(m)
mine = parallel.shard_batch(batch, rank, world)
assert sorted(sum((parallel.shard_batch(batch, r, world) for r in range(world)), [])) == list(range(batch))
for i in mine:
    x = H.random_scalars_mont(100 + i, m)
    assert (lib.ntt("coset_fft", x.copy()) == od.ntt(H.NTT_OPS["coset_fft"], x)).all()
assert lib.launch_count() > 0
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
'''


def test_two_rank_nccl_sharded_msm_and_ntt(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one-GPU box: the gloo test covers the host logic)")
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
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o[-2000:])
        assert "rank %d ok" % r in o
