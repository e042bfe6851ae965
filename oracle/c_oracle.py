"""ctypes wrapper of oracle/sylow_oracle.c (TEST INFRASTRUCTURE: the multi-threaded CPU restatement that
follows sylow's formulas literally; the timed CPU arm of bench.py and the fast oracle of the large
parity tests).  Build with `make -C oracle`."""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libsylow_oracle.so")
_lib = None
DST = b"WARLOCK-CHAOS-V01-CS01-SHA-256"


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def available() -> bool:
    if not os.path.exists(SO):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_build/libsylow_oracle.so missing; run `make -C oracle`")
        try:
            _lib = ctypes.CDLL(SO)
        except OSError:  # built on another machine / stale: rebuild from source here
            subprocess.check_call(["make", "-s", "-B", "-C", HERE])
            _lib = ctypes.CDLL(SO)
        _lib.so_verify_batch.restype = ctypes.c_int
        _lib.so_verify_batch_glued.restype = ctypes.c_int
    return _lib


def cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype=np.uint8):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def pairing_batch(g1, g2, g1_inf=None, g2_inf=None, threads=None):
    g1, g2, g1_inf, g2_inf = _c(g1), _c(g2), _c(g1_inf), _c(g2_inf)
    n = g1.shape[0]
    out = np.empty((n, 384), np.uint8)
    lib().so_pairing_batch(_p(g1), _p(g1_inf), _p(g2), _p(g2_inf), ctypes.c_size_t(n), _p(out), threads or cores())
    return out


def miller_loop_batch(g1, g2, threads=None):
    g1, g2 = _c(g1), _c(g2)
    n = g1.shape[0]
    out = np.empty((n, 384), np.uint8)
    lib().so_miller_loop_batch(_p(g1), _p(g2), ctypes.c_size_t(n), _p(out), threads or cores())
    return out


def final_exp_batch(f, threads=None):
    f = _c(f)
    out = np.empty_like(f)
    lib().so_final_exp_batch(_p(f), ctypes.c_size_t(f.shape[0]), _p(out), threads or cores())
    return out


def miller_product(g1, g2, threads=None):
    g1, g2 = _c(g1), _c(g2)
    out = np.empty(384, np.uint8)
    lib().so_miller_product(_p(g1), _p(g2), ctypes.c_size_t(g1.shape[0]), _p(out), threads or cores())
    return out


def _mul(fn, width, pts, scalars, pts_inf, threads):
    pts, scalars, pts_inf = _c(pts), _c(scalars), _c(pts_inf)
    n = pts.shape[0]
    out = np.empty((n, width), np.uint8)
    inf = np.empty(n, np.uint8)
    fn(_p(pts), _p(pts_inf), _p(scalars), ctypes.c_size_t(n), _p(out), _p(inf), threads or cores())
    return out, inf


def g1_mul_batch(pts, scalars, pts_inf=None, threads=None):
    return _mul(lib().so_g1_mul_batch, 64, pts, scalars, pts_inf, threads)


def g2_mul_batch(pts, scalars, pts_inf=None, threads=None):
    return _mul(lib().so_g2_mul_batch, 128, pts, scalars, pts_inf, threads)


def _msgs(msgs):
    if isinstance(msgs, tuple):
        return _c(msgs[0]), _c(msgs[1], np.uint64)
    offs = np.zeros(len(msgs) + 1, np.uint64)
    if len(msgs):
        offs[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
    return buf, offs


def hash_to_g1_batch(msgs, dst=DST, threads=None):
    buf, offs = _msgs(msgs)
    n = offs.size - 1
    out = np.empty((n, 64), np.uint8)
    inf = np.empty(n, np.uint8)
    lib().so_hash_to_g1_batch(_p(buf), _p(offs), ctypes.c_size_t(n), dst, ctypes.c_size_t(len(dst)), _p(out), _p(inf),
                              threads or cores())
    return out, inf


def sign_batch(sks, msgs, dst=DST, threads=None):
    buf, offs = _msgs(msgs)
    sks = _c(sks)
    n = offs.size - 1
    out = np.empty((n, 64), np.uint8)
    lib().so_sign_batch(_p(sks), _p(buf), _p(offs), ctypes.c_size_t(n), dst, ctypes.c_size_t(len(dst)), _p(out),
                        threads or cores())
    return out


def verify_each(pks, msgs, sigs, dst=DST, threads=None):
    buf, offs = _msgs(msgs)
    pks, sigs = _c(pks), _c(sigs)
    n = offs.size - 1
    ok = np.empty(n, np.uint8)
    lib().so_verify_each(_p(pks), _p(buf), _p(offs), _p(sigs), ctypes.c_size_t(n), dst, ctypes.c_size_t(len(dst)),
                         _p(ok), threads or cores())
    return ok.astype(bool)


def verify_batch(pks, msgs, sigs, dst=DST, threads=None) -> bool:
    buf, offs = _msgs(msgs)
    pks, sigs = _c(pks), _c(sigs)
    n = offs.size - 1
    return bool(lib().so_verify_batch(_p(pks), _p(buf), _p(offs), _p(sigs), ctypes.c_size_t(n), dst,
                                      ctypes.c_size_t(len(dst)), threads or cores()))


def verify_batch_glued(pks, msgs, sigs, dst=DST, threads=None) -> bool:
    """The batch form through glued_miller_loop (pairing.rs:970-1022), as the reference's example runs it."""
    buf, offs = _msgs(msgs)
    pks, sigs = _c(pks), _c(sigs)
    n = offs.size - 1
    return bool(lib().so_verify_batch_glued(_p(pks), _p(buf), _p(offs), _p(sigs), ctypes.c_size_t(n), dst,
                                            ctypes.c_size_t(len(dst)), threads or cores()))


def constants():
    buf = np.zeros(160 + 64 * 24 + 64 * 3, np.uint8)
    lib().so_constants(_p(buf))
    return buf


def _sample_points(m: int, seed: int = 1):
    """m random (G1, G2) pairs: a_i * G1gen, b_i * G2gen, computed by this oracle."""
    rs = np.random.RandomState(seed)
    k = rs.randint(0, 256, size=(2 * m, 32), dtype=np.uint8)
    k[:, 31] &= 0x1F
    g1gen = np.zeros((m, 64), np.uint8)
    g1gen[:, 0], g1gen[:, 32] = 1, 2
    G2 = (10857046999023057135944570762232829481370756359578518086990519993285655852781,
          11559732032986387107991004021392285783925812861821192530917403151452391805634,
          8495653923123431417604973247489272438418190587263600148770280649306958101930,
          4082367875863433681332203403145435568316851327593401208105741076214120093531)
    g2gen = np.tile(np.frombuffer(b"".join(c.to_bytes(32, "little") for c in G2), dtype=np.uint8), (m, 1))
    p, _ = g1_mul_batch(g1gen, k[:m])
    q, _ = g2_mul_batch(g2gen, k[m:])
    return p, q


def time_pairings(budget_s: float):
    """pairings/s of sylow-restated `pairing()` over a bounded batch with one worker per host core.
    Returns (value, cores, kind, sample)."""
    c = cores()
    m = 8 * c
    p, q = _sample_points(m)
    t0 = time.perf_counter()
    pairing_batch(p, q, threads=c)
    dt = time.perf_counter() - t0
    reps = max(1, int(budget_s / max(dt, 1e-3)) - 1)
    reps = min(reps, 512)
    P, Q = np.tile(p, (reps, 1)), np.tile(q, (reps, 1))
    t0 = time.perf_counter()
    pairing_batch(P, Q, threads=c)
    dt = time.perf_counter() - t0
    n = m * reps
    return n / dt, c, "port", ("%d pairings (precompute + Miller loop + final exp, sylow's formulas restated in C, "
                               "4x64 Montgomery), %d worker threads, %.1f s" % (n, c, dt))


def _sample_signatures(m: int, seed: int = 2):
    """m (public key, 32-byte message, signature) triples with distinct signers, made by this oracle."""
    rs = np.random.RandomState(seed)
    sks = rs.randint(0, 256, size=(m, 32), dtype=np.uint8)
    sks[:, 31] &= 0x1F
    msgs = np.zeros((m, 32), np.uint8)
    msgs[:, :8] = np.arange(m, dtype=np.uint64).view(np.uint8).reshape(m, 8)
    msgs[:, 8:] = rs.randint(0, 256, size=(m, 24), dtype=np.uint8)
    packed = (msgs.reshape(-1).copy(), np.arange(m + 1, dtype=np.uint64) * 32)
    G2 = (10857046999023057135944570762232829481370756359578518086990519993285655852781,
          11559732032986387107991004021392285783925812861821192530917403151452391805634,
          8495653923123431417604973247489272438418190587263600148770280649306958101930,
          4082367875863433681332203403145435568316851327593401208105741076214120093531)
    g2gen = np.tile(np.frombuffer(b"".join(c.to_bytes(32, "little") for c in G2), dtype=np.uint8), (m, 1))
    pks, _ = g2_mul_batch(g2gen, sks)
    sigs = sign_batch(sks, packed)
    return pks, packed, sigs


def _tile_msgs(packed, reps):
    buf, offs = packed
    m = offs.size - 1
    return np.tile(buf, reps), np.arange(m * reps + 1, dtype=np.uint64) * 32


def time_verifies(budget_s: float, batch_form: bool):
    """verifies/s of the CPU restatement with one worker per host core on a bounded sample.
    batch_form=False: `verify` per signature (hash-to-curve + two full pairings, src/lib.rs:223-236).
    batch_form=True: the example's batch form (hash + glued Miller loop + ONE final exponentiation per call).
    Returns (value, cores, kind, sample)."""
    c = cores()
    m = 4 * c
    pks, packed, sigs = _sample_signatures(m)
    fn = verify_batch_glued if batch_form else verify_each
    t0 = time.perf_counter()
    r = fn(pks, packed, sigs, threads=c)
    dt = time.perf_counter() - t0
    if not np.all(r):
        raise RuntimeError("CPU oracle rejected its own signatures")
    reps = min(512, max(1, int(budget_s / max(dt, 1e-3)) - 1))
    PK, MS, SG = np.tile(pks, (reps, 1)), _tile_msgs(packed, reps), np.tile(sigs, (reps, 1))
    t0 = time.perf_counter()
    r = fn(PK, MS, SG, threads=c)
    dt = time.perf_counter() - t0
    n = m * reps
    what = ("batch form: hash-to-curve + glued Miller loop over (sig_i, G2gen), (-H(m_i), pk_i) with every G2 point "
            "precomputed + one final exponentiation per call" if batch_form else
            "verify per signature: hash-to-curve + two full pairings")
    return n / dt, c, "port", "%d signatures (%s; sylow's formulas restated in C), %d worker threads, %.1f s" % (n, what, c, dt)
