#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs, one JSON line on stdout.

Workload ("step"), synthetic and seeded:
    configs[0]  one 2^20-point BN254 G1 MSM (points = (a0 + i d) G generated on the GPU, uniform scalars), plus
    configs[1]  the NTT batch of 8 prover polynomials: fft and ifft at n = 2^20 and coset_fft at 4n = 2^22
                (low n coefficients non-zero, the prover's pattern — prover.cpp:418-425).
`value` = milliseconds per step with every input resident in HBM (time-like: lower is better);
`e2e`   = the same step through the calls the reference's signatures really make (polynomial_arithmetic.hpp:28-39,
          scalar_multiplication.hpp:60-61): ONE blocking bbg_msm_g1 and 24 single-polynomial bbg_ntt_fr calls on plain
          aligned_alloc (pageable) host memory, H2D/D2H inside the timed region, SRS points registered once like
          ReferenceString does, long-lived buffers page-locked in place by the library on their second sighting
          (bbg_set_host_register_cache, what the shims switch on);
`e2e_pinned` = the batched / launched form on caller-pinned buffers (bbg_msm_g1_launch + 3 x bbg_ntt_fr_batched), an API
          extension the reference's signatures do not reach — reported next to it, not as the headline.
`msm_2p26` = BASELINE configs[3]: one MSM of 2^26 synthetic points sharded by point range over the N ranks.
With --gpus N (torchrun) the MSM is sharded by point range and the NTT batch by polynomial (configs[2]): total work
fixed => "scaling": "strong"; per-rank device time, max over ranks.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified sources compiled
by oracle/Makefile; multithreaded x86-asm path) on the same workload — all 8 polynomials of every transform — on the
box's host cores (every core the process may run on, whatever OMP_NUM_THREADS a launcher exported).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 20
BATCH = 8
MAC_PER_FIELD_MUL = 136  # 64 + 64 + 8 32x32->64 multiply-adds per 8-limb Montgomery product


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# --------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        self.thread = None
        self.device_index = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------
# workload description shared by both arms
# --------------------------------------------------------------------------------------------------------
def workload_config(args, world):
    return {
        "workload": "configs[0]+configs[1]: 1x MSM 2^%d (BN254 G1, points (a0+i*d)G, uniform scalars) + NTT batch of %d polys: "
                    "fft 2^%d, ifft 2^%d, coset_fft 2^%d" % (args.log_n, BATCH, args.log_n, args.log_n, args.log_n + 2),
        "msm_points": 1 << args.log_n,
        "msm_table": "plain Pippenger windows" if getattr(args, "no_srs_precompute", False) else
                     "the SRS is registered once with the library (ReferenceString::monomials): fixed-base windows built at registration "
                     "(GPU generate_pippenger_precompute_table / pippenger_precomputed), not per step; components_ms.msm_*_plain_windows "
                     "is the same MSM over an unregistered table",
        "ntt_batch": BATCH,
        "ntt_sizes": [1 << args.log_n, 1 << args.log_n, 1 << (args.log_n + 2)],
        "streams": "device-resident step: the MSM runs on a second stream beside the NTT batch; e2e: sequential blocking calls, as the reference's signatures are",
        "sharding": "msm by point range, ntt batch by polynomial (no data-path collective; %d-rank gather of 128-byte partials)" % world,
        "untimed_steps": "warm-up W + 10 more while the clock sampler spins up",
        "l2": "inputs larger than L2 (polynomial batch %d MiB, point table %d MiB), no flush" % (
            BATCH * 32 * (1 << args.log_n) * 5 // (1 << 20) // 1, 128 * (1 << args.log_n) // (1 << 20)),
    }


def algorithmic_macs(log_n):
    """Reference-algorithm 32x32->64 multiply-add counts."""
    n = 1 << log_n
    # get_optimal_bucket_width (scalar_multiplication.cpp:21-81), the thresholds that matter at these sizes
    c = 21 if n >= 14617149 else 18 if n >= 2139094 else 15 if n >= 100000 else 12
    R = (127 + c) // (c + 1)
    msm_muls = 11 * 2 * n * R + 16 * 2 * (1 << c) * R + 7 * (R - 1) * (c + 1)
    fft = (log_n - 1) * n // 2
    ifft = fft + n
    cfft = (log_n + 1) * (4 * n) // 2 + 4 * n
    return {"msm": msm_muls * MAC_PER_FIELD_MUL, "fft": fft * MAC_PER_FIELD_MUL, "ifft": ifft * MAC_PER_FIELD_MUL,
            "coset_fft_4n": cfft * MAC_PER_FIELD_MUL, "msm_mixed_adds": 11 * 2 * n * R * MAC_PER_FIELD_MUL}


# --------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference compiled into oracle/_ref (multithreaded CPU)
# --------------------------------------------------------------------------------------------------------
class ReferenceCpu:
    """Only place (with tests/ and smoke()) allowed to execute oracle/: as the measured CPU baseline, never as
    part of the product path."""

    def __init__(self, log_n):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers as H

        self.H = H
        self.r = H.ref()
        self.kind = "reference"
        if self.r is None:
            raise RuntimeError("oracle/_ref/libbb_ref.so is not built (run __graft_entry__.build() where /root/reference exists)")
        # every core this process may run on — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its ranks,
        # which left the round-1 reference arm on one thread at N = 2, 4, 8
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        t = 1
        while t * 2 <= cores:
            t *= 2
        # evaluation_domain silently needs a power-of-two thread count (SURVEY.md §5 hazard)
        self.r.ref_set_omp_threads(t)
        self.threads = t
        self.log_n = log_n
        self.n = 1 << log_n
        self._setup()

    def _alloc(self, nbytes):
        import ctypes as C

        p = self.r.ref_aligned_alloc(nbytes)
        return p, np.ctypeslib.as_array((C.c_uint64 * (nbytes // 8)).from_address(p))

    def _setup(self):
        H, r, n = self.H, self.r, self.n
        from barretenberg_b200 import synthetic as S

        self.s_ptr, s = self._alloc(32 * n)
        s.reshape(n, 4)[:] = S.random_field(1001, n)
        self.t_ptr, t = self._alloc(128 * n)
        pts = np.zeros((n, 8), dtype=np.uint64)
        # same point family as the GPU arm; built with the reference's own group code
        r.ref_g1_arith_progression(H.ptr(S.to_limbs(S.mont(0x1234567))), H.ptr(S.to_limbs(S.mont(0x89ABC))), H.ptr(pts), n)
        r.ref_generate_pippenger_point_table(H.ptr(pts), H.ptr(t.reshape(2 * n, 8)), n)
        self.p_ptrs, self.q_ptrs = [], []
        for j in range(BATCH):  # the same 8 + 8 polynomials as the GPU arm
            ptr_, p = self._alloc(32 * n)
            p.reshape(n, 4)[:] = S.random_field(2001 + j, n)
            self.p_ptrs.append(ptr_)
            ptr_, q = self._alloc(32 * 4 * n)
            q[:] = 0
            q.reshape(4 * n, 4)[:n] = S.random_field(3001 + j, n)
            self.q_ptrs.append(ptr_)
        self.dom_n = r.ref_domain_new(n)
        self.dom_4n = r.ref_domain_new(4 * n)

    def step_sample(self):
        """The whole step, nothing extrapolated: the MSM and all 8 polynomials of every transform.
        Returns (ms per step, parts)."""
        import ctypes as C

        r = self.r
        out = np.zeros(12, dtype=np.uint64)
        ptrs = (C.c_void_p * 1)(self.s_ptr)
        t0 = time.perf_counter()
        r.ref_batched_scalar_multiplications(ptrs, self.t_ptr, self.n, 1, self.H.ptr(out))
        t1 = time.perf_counter()
        for p in self.p_ptrs:
            r.ref_ntt(self.dom_n, 0, p, None)
        t2 = time.perf_counter()
        for p in self.p_ptrs:
            r.ref_ntt(self.dom_n, 1, p, None)
        t3 = time.perf_counter()
        for q in self.q_ptrs:
            r.ref_ntt(self.dom_4n, 2, q, None)
        t4 = time.perf_counter()
        parts = {"msm_ms": (t1 - t0) * 1e3, "fft_x%d_ms" % BATCH: (t2 - t1) * 1e3, "ifft_x%d_ms" % BATCH: (t3 - t2) * 1e3,
                 "coset_fft_4n_x%d_ms" % BATCH: (t4 - t3) * 1e3}
        return (t4 - t0) * 1e3, parts

    def describe(self):
        return ("the whole step, measured: batched_scalar_multiplications(1 x 2^%d) + %d polynomials x (fft, ifft 2^%d; coset_fft 2^%d); "
                "OMP threads = %d" % (self.log_n, BATCH, self.log_n, self.log_n + 2, self.threads))


def prove_leg(log_gates, with_cpu, num_gpus=1):
    """BASELINE configs[4] next to the step metric: the reference's waffle StandardComposer prover, prebuilt by
    tests/cpp/Makefile (build/ travels to the GPU box), once with Prover::construct_proof on the GPU (HBM-resident rounds)
    and once all-CPU; proofs compared field for field.  Returns None when the binaries are not there."""
    import subprocess

    b = os.path.join(ROOT, "build")
    need = [os.path.join(b, f) for f in ("make_srs", "prover_gpu") + (("prover_cpu",) if with_cpu else ())]
    if not all(os.path.exists(f) for f in need):
        return None
    srs = os.path.join(b, "srs", "transcript.dat")
    os.makedirs(os.path.dirname(srs), exist_ok=True)
    n = 1 << log_gates
    if not os.path.exists(srs) or os.path.getsize(srs) < 28 + 64 * (n - 1) + 256 + 64:
        subprocess.run([need[0], str(n), srs], cwd=ROOT, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    env = dict(os.environ)
    threads = 1
    while threads * 2 <= (os.cpu_count() or 1):
        threads *= 2
    env["OMP_NUM_THREADS"] = str(threads)

    def run(binary, repeat, extra_env=None):
        e = dict(env)
        e.update(extra_env or {})
        out = subprocess.run([os.path.join(b, binary), str(log_gates), str(repeat)], cwd=ROOT, capture_output=True, text=True, timeout=900, env=e)
        if out.returncode != 0:
            raise RuntimeError("%s failed: %s" % (binary, (out.stderr or out.stdout)[-300:]))
        return json.loads(out.stdout.strip().splitlines()[-1])

    multi = {"BBG_NUM_GPUS": str(num_gpus)} if num_gpus > 1 else {}
    gpu = run("prover_gpu", 5, multi)
    cold = run("prover_gpu", 3, dict(multi, BBG_PLONK_KEY_CACHE="0"))
    res = {"workload": "configs[4]: waffle StandardComposer prove, n = 2^%d (bench_plonk.cpp:25-37 circuit, seeded witnesses, synthetic SRS)" % log_gates,
           "gpu_prove_ms": cold["prove_ms_best"], "gpu_prove_ms_warm_key": gpu["prove_ms_best"], "gpu_prove_ms_first": gpu["prove_ms_first"],
           "gpus": num_gpus, "gpu_verified": gpu["verified"] and cold["verified"],
           "note": "gpu_prove_ms (the headline): circuit constants (permutation, selectors) uploaded and transformed every proof, the work the "
                   "reference does per proof; warm_key: constants found unchanged on the device by their fingerprint, only the witness "
                   "uploaded; first: first proof of the process (tables, workspaces, SRS upload); gpus > 1: commitments cut into point "
                   "ranges over the devices inside libbbgpu.so (BBG_NUM_GPUS -> bbg_init_multi)",
           "path": "Prover::construct_proof -> shim/prover_gpu.cpp -> bbg_plonk_* (witness, mappings and selectors uploaded from pageable host memory every proof)"}
    if with_cpu:
        cpu = run("prover_cpu", 1)
        res.update({"cpu_reference_prove_ms": cpu["prove_ms_best"], "cpu_threads": threads, "cpu_verified": cpu["verified"],
                    "proofs_identical": all(cpu["proof"][k] == gpu["proof"][k] and cpu["proof"][k] == cold["proof"][k] for k in cpu["proof"])})
    else:
        res["proofs_identical_warm_vs_cold"] = all(gpu["proof"][k] == cold["proof"][k] for k in gpu["proof"])
    return res


def msm_large_leg(lib, parallel, S, rank, world, log_big, barrier, max_over_ranks, fold, steps=3, warmup=1):
    """BASELINE configs[3]: one MSM of 2^log_big synthetic points ((a0 + i d) G, generated on the device), each rank a
    contiguous point range, partial sums gathered (128 bytes per rank) and folded on the host.  Scalars: independent seeded
    blocks of 2^20.  Gate: the result must equal [(sum_i k_i (a0 + i d)) mod r] G computed as a 1-point MSM; every rank adds
    up its own range's share of that dot product."""
    n = 1 << log_big
    lo, hi = parallel.shard_range(n, rank, world)
    n_loc = hi - lo
    a0, d = 0x7654321, 0xABCDE
    BLK = 1 << 20
    t_setup = time.perf_counter()
    d_points = lib.dev_alloc(max(n_loc, 1) * 64)
    d_table = lib.dev_alloc(max(n_loc, 1) * 128)
    lib.generate_multiples_dev(S.to_limbs(S.mont(a0 + lo * d)), S.to_limbs(S.mont(d)), d_points, n_loc)
    lib.generate_pippenger_point_table_dev(d_points, d_table, n_loc)
    lib.sync()
    lib.dev_free(d_points)
    d_scalars = lib.dev_alloc(max(n_loc, 1) * 32)
    share = 0
    pos = lo
    while pos < hi:  # block by block: never more than 32 MiB of scalars on the host
        blk = pos // BLK
        chunk = S.random_field(7000 + blk, min(BLK, n))
        s_, e_ = pos - blk * BLK, min(hi - blk * BLK, chunk.shape[0])
        piece = np.ascontiguousarray(chunk[s_:e_])
        lib.h2d(d_scalars + (pos - lo) * 32, piece)
        share = (share + S.dot_mod_r(piece, a0 + pos * d, d)) % S.FR_MODULUS
        pos += e_ - s_
    setup_s = time.perf_counter() - t_setup

    def run():
        return fold(lib.msm_partial_dev(d_scalars, d_table, n_loc))

    res = None
    for _ in range(warmup):
        res = run()
    times = []
    for _ in range(steps):
        barrier()
        t0 = time.perf_counter()
        res = run()
        times.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
    # closed form: gather every rank's share of the dot product (as an Fr element through the same partial gather)
    share_limbs = np.zeros(16, dtype=np.uint64)
    share_limbs[:4] = S.to_limbs(share)
    shares = parallel.gather_partials(share_limbs, world, device="cuda" if world > 1 else None)
    ok = None
    if rank == 0:
        total = sum(S.from_limbs(shares[r][:4]) for r in range(world)) % S.FR_MODULUS
        d_g, d_gt, d_s1 = lib.dev_alloc(64), lib.dev_alloc(128), lib.dev_alloc(32)
        lib.generate_multiples_dev(S.to_limbs(S.mont(1)), S.to_limbs(0), d_g, 1)
        lib.generate_pippenger_point_table_dev(d_g, d_gt, 1)
        lib.h2d(d_s1, S.to_limbs(S.mont(total)).reshape(1, 4))
        expect = lib.msm_dev(d_s1, d_gt, 1)
        ok = bool((res == expect).all())
        for p_ in (d_g, d_gt, d_s1):
            lib.dev_free(p_)
    lib.dev_free(d_scalars)
    lib.dev_free(d_table)
    return {"workload": "configs[3]: MSM 2^%d synthetic points (a0 + i d) G, uniform scalars, point ranges over %d rank(s)" % (log_big, world),
            "ms": min(times), "ms_all": times, "n_gpus": world, "points_per_rank": n_loc, "closed_form_ok": ok, "setup_s": setup_s,
            "timing": "wall clock around launch + host fold + gather, barrier before, max over ranks, best of %d after %d warm-up" % (steps, warmup)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    try:
        ref = ReferenceCpu(args.log_n)
    except Exception as e:  # oracle/_ref missing
        print(json.dumps({"impl": "reference", "unavailable": str(e).splitlines()[0]}))
        return
    for _ in range(args.warmup):
        ref.step_sample()
    vals, parts = [], None
    for _ in range(args.steps):
        v, parts = ref.step_sample()
        vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "ms_per_step_msm2p%d_plus_ntt_batch" % args.log_n, "value": value, "unit": "ms", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64x4 Montgomery (x86-64 MULX/ADX asm)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "ms", "cores": ref.threads, "kind": ref.kind, "sample": ref.describe(), "parts_ms": parts},
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import barretenberg_b200 as bb
    from barretenberg_b200 import parallel
    from barretenberg_b200 import synthetic as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = bb.Library(device=local_rank)
    store = dist.distributed_c10d._get_default_store() if world > 1 else None

    def barrier():
        lib.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    log_n, n = args.log_n, 1 << args.log_n
    # ---- shards --------------------------------------------------------------------------------------
    lo, hi = parallel.shard_range(n, rank, world)
    n_loc = hi - lo
    polys = parallel.shard_batch(BATCH, rank, world)  # batch sharded by polynomial; empty for ranks beyond the batch
    P = len(polys)

    # ---- MSM inputs: points generated on the GPU, scalars seeded on the host -------------------------------
    a0, d = 0x1234567, 0x89ABC
    d_points = lib.dev_alloc(max(n_loc, 1) * 64)
    d_table = lib.dev_alloc(max(n_loc, 1) * 128)
    lib.generate_multiples_dev(S.to_limbs(S.mont(a0 + lo * d)), S.to_limbs(S.mont(d)), d_points, n_loc)
    lib.generate_pippenger_point_table_dev(d_points, d_table, n_loc)
    lib.dev_free(d_points)
    scalars_all = S.random_field(1001, n)
    pin = lambda shape: torch.empty(shape, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)  # noqa: E731
    h_scalars = pin((max(n_loc, 1), 4))
    h_scalars[:n_loc] = scalars_all[lo:hi]
    d_scalars = lib.dev_alloc(max(n_loc, 1) * 32)
    lib.h2d(d_scalars, h_scalars[:n_loc])
    # host copy of the table for the e2e arm, registered once (what ReferenceString + the shim do)
    h_table = pin((2 * max(n_loc, 1), 8))
    lib.d2h(h_table, d_table)
    # A ReferenceString is a fixed base: with bbg_set_srs_precompute on (the shims' default), registering it also builds its
    # pre-doubled windows on the device, once — the GPU form of generate_pippenger_precompute_table / pippenger_precomputed
    # (scalar_multiplication.cpp:90-129, :478-573) — and every MSM that names the table adds all windows into one bucket set.
    lib.set_srs_precompute(not args.no_srs_precompute)
    t_reg = time.perf_counter()
    lib.srs_register(h_table)
    srs_register_ms = (time.perf_counter() - t_reg) * 1e3
    d_plain_table = d_table  # an unregistered copy of the same table: plain Pippenger windows
    d_table, fb_c, fb_windows = lib.srs_device_table(h_table) if n_loc else (d_table, 0, 0)

    # ---- NTT inputs --------------------------------------------------------------------------------------
    h_poly_n = pin((max(P, 1), n, 4))
    h_poly_4n = pin((max(P, 1), 4 * n, 4))
    h_poly_4n[:] = 0
    for j, pi in enumerate(polys):
        h_poly_n[j] = S.random_field(2001 + pi, n)
        h_poly_4n[j, :n] = S.random_field(3001 + pi, n)
    d_poly_n = lib.dev_alloc(max(P, 1) * n * 32)
    d_poly_4n = lib.dev_alloc(max(P, 1) * 4 * n * 32)
    lib.h2d(d_poly_n, h_poly_n)
    lib.h2d(d_poly_4n, h_poly_4n)

    # ---- the same inputs in plain aligned_alloc memory: what the reference's callers hand over ---------------------------
    import ctypes as C

    libc = C.CDLL(None)
    libc.aligned_alloc.restype = C.c_void_p
    libc.aligned_alloc.argtypes = [C.c_size_t, C.c_size_t]

    def pageable(shape):
        nbytes = int(np.prod(shape)) * 8
        ptr_ = libc.aligned_alloc(64, (nbytes + 63) // 64 * 64)
        if not ptr_:
            raise MemoryError("aligned_alloc(%d)" % nbytes)
        return np.ctypeslib.as_array((C.c_uint64 * (nbytes // 8)).from_address(ptr_)).reshape(shape)

    pg_scalars = pageable((max(n_loc, 1), 4))
    pg_scalars[:n_loc] = scalars_all[lo:hi]
    pg_table = pageable((2 * max(n_loc, 1), 8))
    pg_table[:] = h_table
    lib.srs_register(pg_table)  # ReferenceString::monomials: uploaded once, like the shim does
    pg_poly_n = [pageable((n, 4)) for _ in range(P)]
    pg_poly_4n = [pageable((4 * n, 4)) for _ in range(P)]
    for j in range(P):
        pg_poly_n[j][:] = h_poly_n[j]
        pg_poly_4n[j][:] = h_poly_4n[j]

    def fold(partial16):
        """Tiny NCCL all-gather of the 128-byte partials over NVLink, then the host-side fold."""
        return lib.fold_partials(parallel.gather_partials(partial16, world, device="cuda"))

    def step_device(overlap=True):
        """One pass of the hot path with everything resident in HBM.  The MSM and the NTT batch are independent, as a
        wire commitment and the other wires' transforms are inside a prover, so the MSM is queued on the library's second
        stream (bbg_msm_g1_partial_dev_launch) and finished after the NTTs have been queued: its latency-bound sort /
        reduction kernels run beside the transforms.  overlap=False is the strictly sequential form, used for the
        per-kernel attribution pass."""
        if not overlap:
            part = lib.msm_partial_dev(d_scalars, d_table, n_loc)
        else:
            ticket = lib.msm_partial_dev_launch(d_scalars, d_table, n_loc)
        if P:
            lib.ntt_dev("fft", d_poly_n, log_n, batch=P)
            lib.ntt_dev("ifft", d_poly_n, log_n, batch=P)
            lib.ntt_dev("coset_fft", d_poly_4n, log_n + 2, batch=P)
        if overlap:
            part = lib.msm_partial_finish(ticket)
        return fold(part)

    def step_host():
        """The calls the reference's signatures make (pippenger: scalar_multiplication.hpp:60-61; fft / ifft / coset_fft:
        polynomial_arithmetic.hpp:28-39): blocking, one polynomial per call, pageable buffers in and out."""
        if n_loc:
            part = parallel.normalized_to_partial(lib.msm(pg_scalars[:n_loc], pg_table, n_loc))  # bbg_msm_g1
        else:
            part = np.zeros(16, dtype=np.uint64)
        for j in range(P):
            lib.ntt("fft", pg_poly_n[j])  # bbg_ntt_fr
        for j in range(P):
            lib.ntt("ifft", pg_poly_n[j])
        for j in range(P):
            lib.ntt("coset_fft", pg_poly_4n[j])
        return fold(part)

    def step_host_pinned():
        """API extension beyond the reference's signatures: caller-pinned buffers, batched transforms, launched MSM."""
        ticket = lib.msm_launch(h_scalars[:n_loc], h_table, n_loc) if n_loc else None  # bbg_msm_g1_launch: beside the NTT copies
        if P:
            lib.ntt("fft", h_poly_n[:P])
            lib.ntt("ifft", h_poly_n[:P])
            lib.ntt("coset_fft", h_poly_4n[:P])
        if n_loc:
            part = parallel.normalized_to_partial(lib.msm_finish(ticket))
        else:
            part = np.zeros(16, dtype=np.uint64)
        return fold(part)

    # ---- parity gate before any timing counts (no oracle here: self-consistency through different code paths) --
    res = step_device()
    if rank == 0:
        s = S.dot_mod_r(scalars_all, a0, d)
        d_g = lib.dev_alloc(64)
        d_gt = lib.dev_alloc(128)
        lib.generate_multiples_dev(S.to_limbs(S.mont(1)), S.to_limbs(0), d_g, 1)
        lib.generate_pippenger_point_table_dev(d_g, d_gt, 1)
        d_s1 = lib.dev_alloc(32)
        lib.h2d(d_s1, S.to_limbs(S.mont(s)).reshape(1, 4))
        expect = lib.msm_dev(d_s1, d_gt, 1)
        for p_ in (d_g, d_gt, d_s1):
            lib.dev_free(p_)
        if not (res == expect).all():
            raise SystemExit("bench.py: MSM parity gate failed (sum k_i P_i != (sum k_i a_i) G)")
    if P:
        chk = np.zeros((n, 4), dtype=np.uint64)
        lib.h2d(d_poly_n, h_poly_n)
        lib.ntt_dev("fft", d_poly_n, log_n, batch=1)
        lib.ntt_dev("ifft", d_poly_n, log_n, batch=1)
        lib.d2h(chk, d_poly_n)
        if not (chk == h_poly_n[0]).all():
            raise SystemExit("bench.py: NTT parity gate failed (ifft(fft(x)) != x)")
        lib.h2d(d_poly_n, h_poly_n)
        lib.h2d(d_poly_4n, h_poly_4n)

    # ---- device-resident timing ----------------------------------------------------------------------------
    # the clock sampler runs from the warm-up through the timed region (a 5-step region lasts only ~75 ms, fewer than
    # two nvidia-smi periods): every sample is taken under this workload's load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3) + 10):
        step_device()
    barrier()
    launches0 = lib.launch_count()
    lib.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    dev_ms = lib.timer_stop()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    dev_ms = max_over_ranks(max(dev_ms, wall_ms))  # the MSM's host-side fold is part of the step: take the longer clock
    value = dev_ms / args.steps

    if args.steps_only:
        # launch lists under ncu: nothing but identical steps after the set-up, so kernel shares compare with `kernels`
        if rank == 0:
            print(json.dumps({"metric": "ms_per_step_msm2p%d_plus_ntt_batch" % log_n, "value": value, "unit": "ms", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "gpu_launches": int(launches), "steps_only": True}))
        barrier()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-op device timings (events on the library stream) ------------------------------------------------
    def time_op(fn, reps=3, host_finish=False):
        fn()
        best = 1e30
        for _ in range(reps):
            lib.sync()
            t = time.perf_counter()
            lib.timer_start()
            fn()
            ev = lib.timer_stop()
            wall = (time.perf_counter() - t) * 1e3
            best = min(best, wall if host_finish else ev)  # the MSM ends with its host-side fold: wall clock covers it
        return best

    def msm_only():
        lib.msm_partial_dev(d_scalars, d_table, n_loc)

    ops_ms = {"msm_2p%d" % log_n: time_op(msm_only, host_finish=True),
              "msm_2p%d_plain_windows" % log_n: time_op(lambda: lib.msm_partial_dev(d_scalars, d_plain_table, n_loc), host_finish=True)}
    if n_loc >= 1024:
        # degenerate digit distribution (constant polynomials do occur in a prover): every scalar equal, so each window's
        # entries fall into one bucket — the block-reduction fix-up path
        d_const = lib.dev_alloc(n_loc * 32)
        lib.h2d(d_const, np.ascontiguousarray(np.broadcast_to(scalars_all[lo], (n_loc, 4))))
        ops_ms["msm_2p%d_constant_scalars" % log_n] = time_op(lambda: lib.msm_partial_dev(d_const, d_table, n_loc), host_finish=True)
        lib.dev_free(d_const)
    if P:
        ops_ms["fft_2p%d_per_poly" % log_n] = time_op(lambda: lib.ntt_dev("fft", d_poly_n, log_n, batch=P)) / P
        ops_ms["ifft_2p%d_per_poly" % log_n] = time_op(lambda: lib.ntt_dev("ifft", d_poly_n, log_n, batch=P)) / P
        ops_ms["coset_fft_2p%d_per_poly" % (log_n + 2)] = time_op(lambda: lib.ntt_dev("coset_fft", d_poly_4n, log_n + 2, batch=P)) / P

    if args.device_only:
        if rank == 0:
            print(json.dumps({"metric": "ms_per_step_msm2p%d_plus_ntt_batch" % log_n, "value": value, "unit": "ms", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "components_ms": ops_ms, "gpu_launches": int(launches), "device_only": True}))
        barrier()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e through the host-buffer ABI ----------------------------------------------------------------------
    def time_host(step, warm):
        for _ in range(warm):
            step()
        barrier()
        lib.timer_start()
        t0 = time.perf_counter()
        steps = max(1, min(args.steps, 5))
        for _ in range(steps):
            step()
        ev_ms = lib.timer_stop()
        return max_over_ranks(max(ev_ms, (time.perf_counter() - t0) * 1e3)) / steps

    # the reference-signature calls: first on cold buffers (every call staged through the pinned ring: what a caller whose
    # buffers never repeat would see), then with the registration cache on (a buffer is page-locked in place at its 6th
    # copy: seven warm-up steps bring every buffer of the step there — the polynomials in steps 2-3, the scalars, which are
    # copied once per step, in step 6 — and the timed steps are the steady state of a caller whose polynomials live as long
    # as a prover's do)
    e2e_staged_ms = time_host(step_host, 1)
    lib.set_host_register_cache(True)
    e2e_ms = time_host(step_host, 7)  # (the scalars are copied once per step: their 6th copy, which page-locks them, is in warm-up step 6)
    reg_stats = lib.host_register_stats()
    for a in [pg_scalars] + pg_poly_n + pg_poly_4n:
        lib.host_buffer_forget(a)
    lib.set_host_register_cache(False)
    e2e_pinned_ms = time_host(step_host_pinned, min(args.warmup, 2))
    h2d_bytes = n_loc * 32 + P * (2 * n * 32 + 4 * n * 32)
    d2h_bytes = 96 + P * (2 * n * 32 + 4 * n * 32)
    # the pageable polynomials went through fft then ifft every step: still the inputs
    if P and not (pg_poly_n[0] == S.random_field(2001 + polys[0], n)).all():
        raise SystemExit("bench.py: e2e parity gate failed (ifft(fft(x)) != x through the pageable host path)")

    # ---- BASELINE configs[3]: MSM 2^26 synthetic points, sharded by point range over the ranks ---------------------------
    msm_big = None
    if not args.no_msm26:
        msm_big = msm_large_leg(lib, parallel, S, rank, world, args.log_big, barrier, max_over_ranks, fold)

    # ---- kernel attribution (separate untimed pass) + measured integer-pipe peak ------------------------------
    lib.profile_enable(True)
    prof_steps = 2
    for _ in range(prof_steps):
        step_device(overlap=False)
    prof = lib.profile_read()
    lib.profile_enable(False)
    peaks = {}
    for mode, name in ((0, "imad_per_s"), (1, "imad_wide_mac_per_s"), (2, "imad_wide_carry_mac_per_s"), (3, "fq_mul_per_s")):
        peaks[name] = lib.microbench(mode, 4096 if mode < 3 else 512)[0]

    if rank == 0:
        macs = algorithmic_macs(log_n)
        mac_peak = max(peaks["imad_wide_mac_per_s"], peaks["imad_wide_carry_mac_per_s"])
        kern = {k: {"ms_per_launch": v[0] / v[1], "launches_per_step": v[1] / prof_steps, "ms_per_step": v[0] / prof_steps} for k, v in prof.items()}
        step_kernel_ms = sum(k["ms_per_step"] for k in kern.values()) or 1.0
        for k in kern.values():
            k["share"] = k["ms_per_step"] / step_kernel_ms
        top = max((k for k in kern if k != "msm_host_finish"), key=lambda k: kern[k]["ms_per_step"])
        measured_peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                measured_peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = measured_peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in measured_peaks else "fallback"
        # algorithmic work of the dominant kernel, per launch
        n4 = 4 * n
        if top in ("ntt_pass_a", "ntt_pass_b"):
            # per step the pass kernel runs on: P polys x (fft, ifft at n) and P polys x coset_fft at 4n
            # a transform's reference-algorithm products are attributed to the two pass kernels by the butterfly stages each
            # one executes (2^20 = 2^9 x 2^11: 9 of 20 stages in pass A; 2^22 = 2^11 x 2^11: half each); ifft's n extra
            # products (x 1/n) and coset_fft's 4n (x g^i) ride on pass A's matrix
            a_share_n = (9.0 / 20.0) if log_n == 20 else ((log_n + 1) // 2) / float(log_n)
            a_share_4n = ((log_n + 3) // 2) / float(log_n + 2)
            mpm = MAC_PER_FIELD_MUL
            pass_a_macs = (macs["fft"] * a_share_n + (macs["ifft"] - n * mpm) * a_share_n + n * mpm
                           + (macs["coset_fft_4n"] - n4 * mpm) * a_share_4n + n4 * mpm)
            all_macs = macs["fft"] + macs["ifft"] + macs["coset_fft_4n"]
            total_macs = P * (pass_a_macs if top == "ntt_pass_a" else all_macs - pass_a_macs)
            total_bytes = P * (2 * n + n4) * 64.0  # read + write 32 B per element per pass
            per_launch_macs = total_macs / kern[top]["launches_per_step"]
            per_launch_bytes = total_bytes / kern[top]["launches_per_step"]
        else:
            per_launch_macs = macs["msm_mixed_adds"] * (n_loc / n) if top == "msm_accumulate" else 0.0
            per_launch_bytes = n_loc * 96.0
        dur_s = kern[top]["ms_per_launch"] * 1e-3
        # ncu evidence for the same kernel (tools/ncu_metrics_json.py on the committed --set full capture): DRAM traffic per
        # launch and the multiply pipe's measured activity — what the pipe really did, next to the algorithmic fraction
        traffic, ncu_rows = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_kernel_metrics.json")) as f:
                ncu_all = json.load(f)
            ncu_rows = {k: v for k, v in ncu_all.items() if k.startswith(top) and isinstance(v, dict)}
            # traffic is quoted per average launch like `achieved`: weight the captured launches by the step's mix
            if top in ("ntt_pass_a", "ntt_pass_b"):
                l1_n, l2_n = (9, 11) if log_n == 20 else ((log_n + 1) // 2, log_n - (log_n + 1) // 2)
                l1_4n = (log_n + 3) // 2
                ka = "%s_L%d" % (top, l1_n if top == "ntt_pass_a" else l2_n)
                kb = "%s_L%d" % (top, l1_4n if top == "ntt_pass_a" else log_n + 2 - l1_4n)
                if ka in ncu_rows and kb in ncu_rows and P == BATCH:
                    traffic = (2 * ncu_rows[ka]["traffic_bytes"] + ncu_rows[kb]["traffic_bytes"]) / 3.0
            elif top in ncu_rows:
                traffic = ncu_rows[top]["traffic_bytes"]
        except (OSError, ValueError, KeyError):
            pass
        roofline = {
            "kernel": top, "bound": "imad", "achieved": per_launch_macs / dur_s / 1e9, "peak": mac_peak / 1e9, "unit": "GMAC/s",
            "frac": per_launch_macs / dur_s / mac_peak, "traffic": traffic,
            "algorithmic_bytes": per_launch_bytes,
            "ncu": {"source": "profiles/ncu_kernel_metrics.json (ncu --set full of tools/ncu_target.py, same kernels, 1 x B200)",
                    "launches": ncu_rows,
                    "pipe_active_pct": (None if not ncu_rows else
                                        {k: v["fmaheavy_pct"] for k, v in ncu_rows.items()})},
            "note": "INT32 multiply pipe, not HBM or tensor: achieved = reference-algorithm 32x32->64 multiply-adds (SURVEY.md §8d; a "
                    "transform's products attributed to the pass kernels by the butterfly stages each executes) / CUDA-event kernel time; "
                    "peak = dependency-free mad.wide.u32 rate measured in this run (narrow IMAD issue rate %.3g/s).  frac is ALGORITHMIC: "
                    "the kernels issue fewer multiply slots than the reference's 136-MAC products imply (precomputed-quotient twiddle "
                    "products 214 instead of 264 slots, trivial twiddles skipped), so what the pipe really did is ncu.pipe_active_pct "
                    "(sm__pipe_fmaheavy_cycles_active of the committed capture), not frac" % peaks["imad_per_s"],
            "hbm": {"achieved": per_launch_bytes / dur_s / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": per_launch_bytes / dur_s / 1e9 / hbm_peak,
                    "peak_source": hbm_src},
        }
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                ref = ReferenceCpu(log_n)
                v, parts = ref.step_sample()
                cpu_baseline = {"value": v, "unit": "ms", "cores": ref.threads, "kind": ref.kind, "sample": ref.describe(), "parts_ms": parts}
            except Exception as e:  # noqa: BLE001
                cpu_baseline = {"value": None, "unit": "ms", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % str(e).splitlines()[0]}
        line = {
            "metric": "ms_per_step_msm2p%d_plus_ntt_batch" % log_n, "value": value, "unit": "ms", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32x8 Montgomery (bn254 Fq/Fr)", "data": "synthetic", "config": workload_config(args, world),
            "components_ms": ops_ms, "clocks": clocks,
            "msm_fixed_base": {"window_bits": fb_c, "windows": fb_windows, "srs_register_ms": srs_register_ms,
                               "table_mib": (fb_windows * n_loc * 128) >> 20},
            "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "note": "the reference-signature calls: 1 x bbg_msm_g1 (blocking, registered SRS) + %d x single-polynomial bbg_ntt_fr on "
                            "aligned_alloc (pageable) buffers; long-lived buffers page-locked in place by the library on their second "
                            "sighting (bbg_set_host_register_cache, as the shims do)" % (3 * P),
                    "staged_first_sighting_ms": e2e_staged_ms, "page_locking": reg_stats},
            "e2e_pinned": {"value": e2e_pinned_ms, "unit": "ms", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                           "note": "API extension, not the reference's call: caller-pinned buffers, bbg_msm_g1_launch / _finish around 3 x bbg_ntt_fr_batched"},
            "gpu_launches": int(launches), "roofline": roofline, "kernels": kern, "imad_peaks": peaks,
            "algorithmic_gmac": {k: v / 1e9 for k, v in macs.items()},
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if msm_big is not None:
            msm_big["frac_of_imad_peak"] = (algorithmic_macs(args.log_big)["msm"] / (msm_big["ms"] * 1e-3)) / (mac_peak * world) if msm_big.get("ms") else None
            line["msm_2p%d" % args.log_big] = msm_big
        if not args.no_prove:
            # configs[4].  N = 1: one GPU behind the shim (+ the all-CPU reference prover beside it).  N > 1 (torchrun): rank 0
            # runs the same binary with BBG_NUM_GPUS = N — one process driving N devices through bbg_init_multi — while the
            # other ranks idle at the barrier below.
            try:
                prove = prove_leg(log_n, with_cpu=(world == 1 and not args.no_cpu_baseline), num_gpus=world)
            except Exception as e:  # noqa: BLE001
                prove = {"unavailable": str(e).splitlines()[0][:200]}
            if prove is not None:
                line["prove"] = prove
        print(json.dumps(line))
        if store is not None:
            store.set("bbg_bench_rank0_done", "1")
    elif store is not None:
        # wait on the rendezvous store, not in an NCCL barrier: a collective would keep a spinning kernel on this GPU while
        # rank 0's prover leg drives it through bbg_init_multi
        import datetime

        store.wait(["bbg_bench_rank0_done"], datetime.timedelta(minutes=30))
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=LOG_N, help="log2 of the MSM / NTT size (BASELINE: 20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prove", action="store_true", help="skip the full-prover leg (BASELINE configs[4])")
    ap.add_argument("--no-msm26", action="store_true", help="skip the 2^26-point MSM leg (BASELINE configs[3])")
    ap.add_argument("--log-big", type=int, default=26, help="log2 of the large synthetic MSM (BASELINE configs[3]: 26)")
    ap.add_argument("--no-srs-precompute", action="store_true", help="plain Pippenger windows for the registered SRS too (no fixed-base tables)")
    ap.add_argument("--steps-only", action="store_true", help="set-up, warm-up and timed steps only (launch lists under ncu)")
    ap.add_argument("--device-only", action="store_true", help="skip the e2e, microbench and CPU-baseline legs (short runs under ncu)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    # The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on
    # stdout): point fd 1 at stderr while working and restore it only for the final line.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import contextlib
    import io

    captured = io.StringIO()
    with contextlib.redirect_stdout(captured):
        if args.impl == "reference":
            run_reference(args, rank, world)
        else:
            run_b200(args, rank, local_rank, world)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    out = captured.getvalue().strip()
    if out:
        os.write(1, (out.splitlines()[-1] + "\n").encode())


if __name__ == "__main__":
    main()
