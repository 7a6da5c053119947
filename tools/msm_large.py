"""BASELINE configs[2]/[3]: one MSM of 2^log_n synthetic points (random multiples of the generator, generated on the
GPU) sharded by point range over the ranks of a torchrun launch.  Prints one JSON line on rank 0.

    python tools/msm_large.py --log-n 26
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/msm_large.py --log-n 26

Correctness gate (no oracle): sum_i k_i (a0 + i d) G must equal the 1-point MSM [(sum_i k_i (a0 + i d)) mod r] G."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=26)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--profile", action="store_true", help="add the per-kernel CUDA-event breakdown of one MSM")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import barretenberg_b200 as bb
    from barretenberg_b200 import parallel
    from barretenberg_b200 import synthetic as S

    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = bb.Library(device=local_rank)
    n = 1 << args.log_n
    lo, hi = parallel.shard_range(n, rank, world)
    n_loc = hi - lo
    a0, d = 0x1234567, 0x89ABC
    t0 = time.perf_counter()
    d_points = lib.dev_alloc(n_loc * 64)
    d_table = lib.dev_alloc(n_loc * 128)
    lib.generate_multiples_dev(S.to_limbs(S.mont(a0 + lo * d)), S.to_limbs(S.mont(d)), d_points, n_loc)
    lib.generate_pippenger_point_table_dev(d_points, d_table, n_loc)
    lib.sync()
    lib.dev_free(d_points)
    gen_s = time.perf_counter() - t0
    # scalars: 1 Mi-element seeded blocks so every rank can rebuild any range
    BLK = 1 << 20

    def scalars_range(a, b):
        out = np.empty((b - a, 4), dtype=np.uint64)
        pos = a
        while pos < b:
            blk = pos // BLK
            chunk = S.random_field(5000 + blk, min(BLK, n))
            s, e = pos - blk * BLK, min(b - blk * BLK, chunk.shape[0])
            out[pos - a:pos - a + (e - s)] = chunk[s:e]
            pos += e - s
        return out

    h_scalars = scalars_range(lo, hi)
    d_scalars = lib.dev_alloc(n_loc * 32)
    lib.h2d(d_scalars, h_scalars)

    def run():
        part = lib.msm_partial_dev(d_scalars, d_table, n_loc)
        return lib.fold_partials(parallel.gather_partials(part, world, device="cuda"))

    for _ in range(args.warmup):
        res = run()
    times = []
    for _ in range(args.steps):
        lib.sync()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = run()
        dt = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        times.append(dt)
    prof = None
    if args.profile:
        lib.profile_enable(True)
        run()
        prof = {k: round(v[0] / v[1], 3) for k, v in lib.profile_read().items()}
        lib.profile_enable(False)
    ok = None
    if rank == 0 and not args.no_check:
        t0 = time.perf_counter()
        s = 0
        pos = 0
        while pos < n:  # blockwise so the 2^26 case never holds more than one block of scalars
            e = min(pos + BLK, n)
            blk = scalars_range(pos, e)
            s = (s + S.dot_mod_r(blk, a0 + pos * d, d)) % S.FR_MODULUS
            pos = e
        d_g, d_gt, d_s1 = lib.dev_alloc(64), lib.dev_alloc(128), lib.dev_alloc(32)
        lib.generate_multiples_dev(S.to_limbs(S.mont(1)), S.to_limbs(0), d_g, 1)
        lib.generate_pippenger_point_table_dev(d_g, d_gt, 1)
        lib.h2d(d_s1, S.to_limbs(S.mont(s)).reshape(1, 4))
        expect = lib.msm_dev(d_s1, d_gt, 1)
        ok = bool((res == expect).all())
        check_s = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"msm_log_n": args.log_n, "n_gpus": world, "ms": min(times), "ms_all": times, "points_per_rank": n_loc,
                          "closed_form_ok": ok, "kernels_ms": prof, "point_generation_s": gen_s, "check_s": None if ok is None else check_s}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
