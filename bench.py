#!/usr/bin/env python3
"""bench.py - BN254 pairings/sec (and BLS verifies/sec) on N B200s, beside the host-core CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 20]

Workload (BASELINE.json configs[1]): a batch of 2^20 independent optimal-ate pairings on random
G1 x G2 points per GPU (weak scaling: every rank runs its own contiguous 2^20 slice, no data-path
collective).  One "step" = one pass of pairing_batch (fused Miller loop kernel + final-exponentiation
kernel) over the whole batch.  `value` is timed with CUDA events with inputs resident in HBM;
`e2e` goes through the host-pointer C-ABI call with pinned host buffers (H2D + D2H inside the timed
region).  The inputs (192 MiB) and outputs (384 MiB) are larger than the 126 MB L2.
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

LIMB_PRODUCTS_PER_FP_MUL = 136        # 8-limb CIOS: 64 (a*b) + 64 (m*p) + 8 (m)        SURVEY.md 8(d)
FP_MUL_MILLER_FUSED = 8444            # fused Miller loop, per pair                     SURVEY.md 8(d)
FP_MUL_FINAL_EXP = 8822 + 380         # final exponentiation + one Fermat inversion     SURVEY.md 8(d)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_miller launch at 2^20 pairs, from the committed
# `ncu --set full` capture profiles/r01l_k_miller_2pow20.json (local-memory frame spill traffic; the
# algorithmic bytes are 576 B per pairing: this kernel is bound by the integer-multiply pipe, not by HBM - the
# 89 GB are write-backs of the 2.8 KB/thread frame from L2, 506 GB/s or 8 % of the measured HBM bandwidth).
NCU_DRAM_BYTES_K_MILLER_2POW20 = 89.1e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=20, help="log2 of pairings per GPU per step")
    ap.add_argument("--verify-log2n", type=int, default=20, help="log2 of signatures for the verify_batch leg (0 = skip)")
    ap.add_argument("--extras", type=int, default=1, help="also time the Groth16-shaped check and scalar-mul configs")
    ap.add_argument("--groth-log2n", type=int, default=18)
    ap.add_argument("--mul-log2n", type=int, default=22)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle timed on the host cores ("port": sylow itself cannot be built here)
# ------------------------------------------------------------------------------------------------
def cpu_pairings_per_s(budget_s: float):
    """Times the CPU restatement of sylow's pairing on a bounded sample of the same workload, with
    every host core.  Returns (pairings/s, cores, kind, sample description)."""
    try:
        from oracle import c_oracle

        if c_oracle.available():
            return c_oracle.time_pairings(budget_s)
    except Exception as e:  # the C oracle is optional test infrastructure; fall back to the Python one
        print("bench: C oracle unavailable (%s); timing the Python oracle" % e, file=sys.stderr)
    import random

    from oracle import bn254_py as o
    from tests import wire as w

    rng = random.Random(1)
    pts = [(w.rand_g1(rng), w.rand_g2(rng)) for _ in range(4)]
    t0 = time.perf_counter()
    done = 0
    while time.perf_counter() - t0 < budget_s:
        p, q = pts[done % 4]
        o.pairing_affine(p, q)
        done += 1
    dt = time.perf_counter() - t0
    return done / dt, 1, "port", "%d pairings, Python big-int oracle, 1 thread" % done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals = []
    steps = max(1, args.steps)
    per_step = max(1.0, min(20.0, 150.0 / (steps + args.warmup)))
    for i in range(args.warmup + steps):
        v, cores, kind, sample = cpu_pairings_per_s(per_step)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "pairings_per_s", "value": value, "unit": "pairings/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "2^%d independent optimal-ate pairings (Miller loop + final exponentiation) on "
                               "random G1xG2 points per GPU, BASELINE configs[1]" % args.log2n,
                   "pairings_per_gpu_per_step": 1 << args.log2n,
                   "sample": "each step times a bounded sample of that workload on all host cores"},
        "reference_published": {"pairing_ms": 8.183, "pairings_per_s_per_core": 1e3 / 8.183,
                                "source": "sylow_devguide.pdf (hardware unstated)"},
        "cpu_baseline": {"value": value, "unit": "pairings/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "pairings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import sylow_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = sylow_b200.Engine(local)
    n = 1 << args.log2n
    K, W = args.steps, max(3, args.warmup)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- synthetic inputs, generated on the device: P_i = a_i * G1, Q_i = b_i * G2 (seed 1 + rank)
    rs = np.random.RandomState(1 + rank)
    R_ORDER = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001

    def rand_scalars(m):
        raw = rs.randint(0, 256, size=(m, 32), dtype=np.uint8)
        raw[:, 31] &= 0x1F  # < 2^253 < r: uniform enough for synthetic inputs, never 0 in practice
        return raw

    g1gen = np.zeros((1, 64), np.uint8)
    g1gen[0, 0], g1gen[0, 32] = 1, 2
    G2_GEN = (10857046999023057135944570762232829481370756359578518086990519993285655852781,
              11559732032986387107991004021392285783925812861821192530917403151452391805634,
              8495653923123431417604973247489272438418190587263600148770280649306958101930,
              4082367875863433681332203403145435568316851327593401208105741076214120093531)  # g2.rs:47-77
    g2gen = np.frombuffer(b"".join(c.to_bytes(32, "little") for c in G2_GEN), dtype=np.uint8).reshape(1, 128)
    t_gen = time.perf_counter()
    d_g1gen = torch.from_numpy(np.repeat(g1gen, n, axis=0)).to(dev)
    d_g2gen = torch.from_numpy(np.repeat(g2gen, n, axis=0)).to(dev)
    d_a = torch.from_numpy(rand_scalars(n)).to(dev)
    d_b = torch.from_numpy(rand_scalars(n)).to(dev)
    d_g1 = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_g2 = torch.empty((n, 128), dtype=torch.uint8, device=dev)
    eng.g1_mul_batch_dev(d_g1gen, d_a, d_g1)
    eng.g2_mul_batch_dev(d_g2gen, d_b, d_g2)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    del d_g1gen, d_g2gen
    d_out = torch.empty((n, 384), dtype=torch.uint8, device=dev)

    # ---- IMAD roofline denominator, measured live (IMAD.WIDE.U32.X carry chains, all SMs)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    best = 0.0
    for _ in range(5):
        ms, ops = eng.imad_probe(12, sms * 8, 256, 2000)
        best = max(best, ops / (ms * 1e-3))
    peak_limb_products = best  # one IMAD.WIDE.U32 = one 32x32->64 limb product

    # ---- warm-up, then the timed region (CUDA events on torch's current stream = the launch stream)
    for _ in range(W):
        eng.pairing_batch_dev(d_g1, d_g2, d_out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        eng.pairing_batch_dev(d_g1, d_g2, d_out)
    e1.record()
    torch.cuda.synchronize()
    launches = eng.launch_count - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    barrier()
    clocks = sampler.stop()
    value = world * n * K / (ms_total * 1e-3)

    # ---- per-kernel durations for the roofline (same buffers, same stream, CUDA events)
    def time_ms(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    d_f = torch.empty((n, 384), dtype=torch.uint8, device=dev)
    ms_miller = time_ms(lambda: eng.miller_loop_batch_dev(d_g1, d_g2, d_f))
    ms_fexp = time_ms(lambda: eng.final_exp_batch_dev(d_f, d_out))
    lp_miller = n * FP_MUL_MILLER_FUSED * LIMB_PRODUCTS_PER_FP_MUL / (ms_miller * 1e-3)
    lp_fexp = n * FP_MUL_FINAL_EXP * LIMB_PRODUCTS_PER_FP_MUL / (ms_fexp * 1e-3)
    del d_f

    # ---- e2e: host-pointer C-ABI call, pinned host buffers, H2D + D2H inside the timed region
    h_g1 = torch.empty((n, 64), dtype=torch.uint8).pin_memory()
    h_g2 = torch.empty((n, 128), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((n, 384), dtype=torch.uint8).pin_memory()
    h_g1.copy_(d_g1)
    h_g2.copy_(d_g2)
    import ctypes

    def e2e_step():
        st = eng._lib.sylow_b200_pairing_batch(eng._h, ctypes.c_void_p(h_g1.data_ptr()), None,
                                               ctypes.c_void_p(h_g2.data_ptr()), None, n,
                                               ctypes.c_void_p(h_out.data_ptr()))
        if st != 0:
            raise RuntimeError("pairing_batch failed: %d" % st)

    e2e_step()
    Ke = max(1, min(K, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    dt = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * Ke / dt
    same = bool((h_out[:4096].to(dev) == d_out[:4096]).all().item())

    # ---- secondary metric: BLS verify_batch (hash-to-curve + 2 Miller loops per signature + one final exp)
    verify = None
    if args.verify_log2n > 0:
        nv = 1 << args.verify_log2n
        msgs = np.zeros((nv, 32), np.uint8)
        msgs[:, :8] = np.arange(nv, dtype=np.uint64).view(np.uint8).reshape(nv, 8)
        msgs[:, 8:] = rs.randint(0, 256, size=(nv, 24), dtype=np.uint8)
        offs = (np.arange(nv + 1, dtype=np.uint64) * 32)
        sks = rand_scalars(nv)
        sigs = eng.sign_batch(sks, (msgs.reshape(-1), offs))
        d_sk = torch.from_numpy(sks).to(dev)
        d_pk = torch.empty((nv, 128), dtype=torch.uint8, device=dev)
        eng.g2_mul_batch_dev(torch.from_numpy(np.repeat(g2gen, nv, axis=0)).to(dev), d_sk, d_pk)
        d_msgs = torch.from_numpy(msgs.reshape(-1)).to(dev)
        d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
        d_sigs = torch.from_numpy(sigs).to(dev)
        d_part = torch.empty(384, dtype=torch.uint8, device=dev)
        from sylow_b200 import sharding

        def verify_step():
            # per-GPU partial (hash + 2 Miller loops per signature + tree product), then the ONE exchange
            # of the path: all-gather of the 384-byte partials over NCCL, product + one final exponentiation
            eng.verify_batch_partial_dev(d_pk, d_msgs, d_offs, d_sigs, d_part)
            parts = sharding.all_gather_partials(d_part.cpu().numpy(), device=dev)
            return eng.verify_batch_finish(parts)

        ok = verify_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            ok = verify_step() and ok
        dt_v = max_over_ranks(time.perf_counter() - t0) / 2
        ms_v_kernels = time_ms(lambda: eng.verify_batch_partial_dev(d_pk, d_msgs, d_offs, d_sigs, d_part), reps=2)
        d_hm = torch.empty((nv, 64), dtype=torch.uint8, device=dev)
        ms_hash = time_ms(lambda: eng.hash_to_g1_batch_dev(d_msgs, d_offs, d_hm), reps=2)
        # S = 1 variant (the reference example's own setting: one signer, many messages), host-pointer call
        same_signer = None
        if world == 1:
            sk1 = np.repeat(sks[:1], nv, axis=0)
            sigs1 = eng.sign_batch(sk1, (msgs.reshape(-1), offs))
            pk1, _ = eng.g2_mul_batch(g2gen, sks[:1])
            eng.verify_batch_same_signer(pk1[0], (msgs.reshape(-1), offs), sigs1)
            t1 = time.perf_counter()
            ok1 = eng.verify_batch_same_signer(pk1[0], (msgs.reshape(-1), offs), sigs1)
            same_signer = {"verifies_per_s": nv / (time.perf_counter() - t1), "batch_ok": bool(ok1),
                    "note": "one signer: n hashes + 2n point additions + two Miller loops per batch, host buffers (e2e)"}
        verify = {"verifies_per_s": world * nv / dt_v, "signatures_per_gpu": nv, "distinct_signers": nv,
                  "same_signer": same_signer,
                  "batch_ok": bool(ok), "ms_per_batch": dt_v * 1e3, "ms_partial_kernels": ms_v_kernels,
                  "ms_hash_to_curve": ms_hash,
                  "verifies_per_s_excl_hashing": world * nv / (max_over_ranks(dt_v * 1e3 - ms_hash) * 1e-3),
                  "note": "per GPU: hash-to-curve + one fused Miller loop per signature, signatures summed into one "
                          "Miller loop against the generator, tree product; all-gather of the 384-byte partials; one "
                          "final exponentiation per batch (device-resident inputs)"}

    # ---- further BASELINE configs, as extra keys (device-resident inputs, CUDA events)
    extras = {}
    if args.extras:
        # config #4: Groth16-shaped 4-pair product check, three G2 points fixed across all checks
        nc = 1 << args.groth_log2n
        co = eng.g2_precompute(d_g2[:3].cpu().numpy())
        d_tab = torch.empty(3 * 87 * 192, dtype=torch.uint8, device=dev)
        eng.tables_to_device(co, d_tab)
        torch.cuda.synchronize()
        d_g1c = d_g1[: 4 * nc].contiguous() if 4 * nc <= n else d_g1.repeat((4 * nc + n - 1) // n, 1)[: 4 * nc].contiguous()
        d_g2c = d_g2[:nc].contiguous()
        d_ok = torch.empty(nc, dtype=torch.uint8, device=dev)
        ms_g = time_ms(lambda: eng.pairing_check_fixed_batch_dev(d_g1c, d_g2c, d_tab, 1, 3, d_ok), reps=2)
        extras["groth16_4pair_checks_per_s"] = world * nc / (max_over_ranks(ms_g) * 1e-3)
        extras["groth16_checks_per_gpu"] = nc
        # config #5: variable-base scalar multiplication, 254-bit scalars
        nm = 1 << args.mul_log2n
        d_k = torch.from_numpy(rand_scalars(nm)).to(dev)
        d_p1 = d_g1[:nm].contiguous() if nm <= n else d_g1.repeat((nm + n - 1) // n, 1)[:nm].contiguous()
        d_p2 = d_g2[:nm].contiguous() if nm <= n else d_g2.repeat((nm + n - 1) // n, 1)[:nm].contiguous()
        d_o1 = torch.empty((nm, 64), dtype=torch.uint8, device=dev)
        d_o2 = torch.empty((nm, 128), dtype=torch.uint8, device=dev)
        ms_1 = time_ms(lambda: eng.g1_mul_batch_dev(d_p1, d_k, d_o1), reps=2)
        ms_2 = time_ms(lambda: eng.g2_mul_batch_dev(d_p2, d_k, d_o2), reps=2)
        extras["g1_scalar_muls_per_s"] = world * nm / (max_over_ranks(ms_1) * 1e-3)
        extras["g2_scalar_muls_per_s"] = world * nm / (max_over_ranks(ms_2) * 1e-3)
        extras["scalar_muls_per_gpu"] = nm
        del d_k, d_p1, d_p2, d_o1, d_o2
        # threshold aggregation, 2^15 sets of 8 partial signatures (host buffers, wall clock)
        tt = 8
        ns = max(1, min(1 << 15, n // tt))
        ids = np.tile(np.arange(1, tt + 1, dtype=np.uint64), (ns, 1)) + (np.arange(ns, dtype=np.uint64) % 5)[:, None]
        h_sig = d_g1[: ns * tt].cpu().numpy().reshape(ns, tt, 64)
        eng.threshold_aggregate_batch(ids, h_sig)
        t0 = time.perf_counter()
        eng.threshold_aggregate_batch(ids, h_sig)
        extras["threshold_aggregations_per_s"] = world * ns / max_over_ranks(time.perf_counter() - t0)
        extras["threshold_shares_per_set"] = tt
        # SURVEY 8(f) ranks 1-3 through the host-pointer calls (wall clock, copies included; 2^17 items, 2^15 for Gt)
        def wall(fn):
            fn()
            t0 = time.perf_counter()
            fn()
            return max_over_ranks(time.perf_counter() - t0)

        nx = min(1 << 17, n)
        h1 = d_g1[:nx].cpu().numpy()
        h2 = d_g2[:nx].cpu().numpy()
        hk = rand_scalars(nx)
        hm = (np.arange(nx, dtype=np.uint64).view(np.uint8).reshape(nx, 8).copy().reshape(-1),
              np.arange(nx + 1, dtype=np.uint64) * 8)
        extras["next_rows"] = {
            "items": nx,
            "g1_validate_per_s": world * nx / wall(lambda: eng.g1_validate_batch(h1)),
            "g2_validate_per_s": world * nx / wall(lambda: eng.g2_validate_batch(h2)),
            "g2_be_roundtrip_per_s": world * nx / wall(
                lambda: eng.g2_from_be_bytes_batch(eng.g2_to_be_bytes_batch(h2), eip_mode=False)),
            "sign_per_s": world * nx / wall(lambda: eng.sign_batch(hk, hm)),
            "note": "host buffers, wall clock: G1Affine::new / G2Projective::new checks, big-endian codec round trip, sign",
        }
        # one large multi-scalar multiplication (bucket method) next to the same sum by ladders + tree sum
        nm2 = min(1 << 20, n)
        hp, hs = d_g1[:nm2].cpu().numpy(), rand_scalars(nm2)
        t_b = wall(lambda: eng.g1_msm_bucket(hp, hs))
        extras["next_rows"]["g1_msm_2pow%d_ms" % int(np.log2(nm2))] = {"bucket": t_b * 1e3}
        if nm2 <= 1 << 20:
            t_l = wall(lambda: eng.g1_sum(*eng.g1_mul_batch(hp, hs)))
            extras["next_rows"]["g1_msm_2pow%d_ms" % int(np.log2(nm2))]["ladders"] = t_l * 1e3
        ng = min(1 << 15, n)
        hgt = d_out[:ng].cpu().numpy()
        extras["next_rows"]["gt_mul_per_s"] = world * ng / wall(lambda: eng.gt_mul_batch(hgt, hk[:ng]))

    cpu = None
    if rank == 0 and world == 1:
        v, cores, kind, sample = cpu_pairings_per_s(args.cpu_seconds)
        cpu = {"value": v, "unit": "pairings/s", "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": "pairings_per_s", "value": value, "unit": "pairings/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32x8 Montgomery (integer)", "data": "synthetic",
            "config": {"workload": "2^%d independent optimal-ate pairings (Miller loop + final exponentiation) on "
                                   "random G1xG2 points per GPU, BASELINE configs[1]" % args.log2n,
                       "pairings_per_gpu_per_step": n, "l2_policy": "inputs+outputs (576 MiB) larger than L2",
                       "parallelism": "contiguous slice per GPU, no data-path collective", "input_gen_s": t_gen},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairings/s", "h2d_bytes_per_step": n * 192,
                    "d2h_bytes_per_step": n * 384, "steps": Ke, "matches_device_path": same},
            "gpu_launches": launches,
            "roofline": {"bound": "imad", "kernel": "k_miller (fused Miller loop)",
                         "achieved": lp_miller / 1e12, "peak": peak_limb_products / 1e12,
                         "unit": "T limb-products/s (32x32->64 multiply-adds)",
                         "frac": lp_miller / peak_limb_products,
                         "traffic": NCU_DRAM_BYTES_K_MILLER_2POW20 if args.log2n == 20 else None,
                         "traffic_note": "bytes per k_miller launch from profiles/r01l (ncu --set full); algorithmic "
                                         "bytes per launch = n * 576",
                         "peak_source": "measured live: IMAD.WIDE.U32.X carry-chain probe on all SMs",
                         "ms_per_launch": ms_miller,
                         "ncu_fmaheavy_pipe_pct": 69.2 if args.log2n == 20 else None,
                         "ncu_note": "sm__pipe_fmaheavy_cycles_active of one k_miller launch at 2^20 pairs, "
                                     "profiles/r01l_k_miller_2pow20.json",
                         "hbm_gbs_load_store": n * (192 + 384) / (ms_miller * 1e-3) / 1e9,
                         "final_exp": {"ms_per_launch": ms_fexp, "achieved": lp_fexp / 1e12,
                                       "frac": lp_fexp / peak_limb_products}},
            "reference_published": {"pairing_ms": 8.183, "pairings_per_s_per_core": 1e3 / 8.183,
                                    "source": "sylow_devguide.pdf (hardware unstated)"},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if verify:
            line["verify_batch"] = verify
        if extras:
            line["other_configs"] = extras
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
