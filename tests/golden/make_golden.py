#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into tests/golden/reference_kats.json.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It copies *numbers* (test vectors and constants), never code.  Each entry records the
file:line range it came from.  The committed JSON is what the tests read.
"""
from __future__ import annotations

import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def lines(path, lo, hi):
    with open(os.path.join(REF, path)) as f:
        return "".join(f.readlines()[lo - 1:hi])


def _int(tok):
    tok = tok.strip().replace("_", "")
    return int(tok, 16) if tok.lower().startswith("0x") else int(tok)


def word_arrays(path, lo, hi):
    """All `[w0, w1, w2, w3]` little-endian u64 word arrays in a line range, as integers."""
    txt = lines(path, lo, hi)
    out = []
    for m in re.finditer(r"\[\s*((?:0x[0-9a-fA-F_]+|\d[\d_]*)\s*,\s*){3}(?:0x[0-9a-fA-F_]+|\d[\d_]*)\s*,?\s*\]", txt):
        toks = [t for t in re.split(r"[\[\],\s]+", m.group(0)) if t]
        w = [_int(t) for t in toks]
        out.append(w[0] | (w[1] << 64) | (w[2] << 128) | (w[3] << 192))
    return out


def decimals(path, lo, hi):
    return [int(m) for m in re.findall(r'"(\d{20,})"', lines(path, lo, hi))]


def hexblob(path, lo, hi):
    """Concatenate the `//! <hex>\\` continuation lines of one hex::decode literal."""
    return "".join(re.findall(r"//!\s+([0-9a-fA-F]{64})\\?", lines(path, lo, hi)))


def h(v):
    return hex(v)


def main():
    k = {}

    # --- Fp multiplication (src/fields/fp.rs:1035-1099)
    w = word_arrays("src/fields/fp.rs", 1035, 1099)
    # arrays in order: a,b,6 | c,d,res | e,f,res | a,zero,one | large,res
    k["fp_mul"] = {
        "source": "src/fields/fp.rs:1035-1099",
        "cases": [[h(w[0]), h(w[1]), h(w[2])], [h(w[3]), h(w[4]), h(w[5])], [h(w[6]), h(w[7]), h(w[8])],
                  [h(w[12]), h(w[12]), h(w[13])]],
    }

    # --- Fp2 multiplication / division (src/fields/fp2.rs:574-609, 712-737)
    w = word_arrays("src/fields/fp2.rs", 574, 609)
    k["fp2_mul"] = {
        "source": "src/fields/fp2.rs:574-609",
        "cases": [[[h(x) for x in w[0:2]], [h(x) for x in w[2:4]], [h(x) for x in w[4:6]]],
                  [[h(x) for x in w[6:8]], [h(x) for x in w[8:10]], [h(x) for x in w[10:12]]]],
    }
    txt_lo = 700
    w = word_arrays("src/fields/fp2.rs", txt_lo, 737)
    # a, b then c (a/b)
    k["fp2_div"] = {
        "source": "src/fields/fp2.rs:%d-737" % txt_lo,
        "cases": [[[h(x) for x in w[-6:-4]], [h(x) for x in w[-4:-2]], [h(x) for x in w[-2:]]]],
    }

    # --- Fp6 multiplication (src/fields/fp6.rs:668-762)
    w = word_arrays("src/fields/fp6.rs", 668, 762)
    k["fp6_mul"] = {
        "source": "src/fields/fp6.rs:668-762",
        "cases": [[[h(x) for x in w[0:6]], [h(x) for x in w[6:12]], [h(x) for x in w[12:18]]],
                  [[h(x) for x in w[18:24]], [h(x) for x in w[18:24]], [h(x) for x in w[24:30]]]],
    }

    # --- Frobenius tables (src/fields/fp6.rs:40-179, src/fields/fp12.rs:29-172)
    def fp2_table(path, lo, hi):
        txt = lines(path, lo, hi)
        # entries are Fp2::new(&[ A, B ]) where A/B are Fp::ONE / Fp::ZERO / Fp::new(U256::from_words([..]))
        entries = []
        for m in re.finditer(r"Fp2::new\(&\[(.*?)\]\),\n(?=\s*(?://|Fp2::new|\];))", txt, re.S):
            body = m.group(1)
            vals = []
            for t in re.finditer(r"Fp::ONE|Fp::ZERO|from_words\(\[(.*?)\]\)", body, re.S):
                if t.group(0) == "Fp::ONE":
                    vals.append(1)
                elif t.group(0) == "Fp::ZERO":
                    vals.append(0)
                else:
                    ws = [_int(x) for x in t.group(1).split(",") if x.strip()]
                    vals.append(ws[0] | (ws[1] << 64) | (ws[2] << 128) | (ws[3] << 192))
            assert len(vals) == 2, body
            entries.append([h(vals[0]), h(vals[1])])
        return entries

    c1 = fp2_table("src/fields/fp6.rs", 40, 108)
    c2 = fp2_table("src/fields/fp6.rs", 109, 179)
    c12 = fp2_table("src/fields/fp12.rs", 29, 172)
    assert len(c1) == 6 and len(c2) == 6 and len(c12) == 12, (len(c1), len(c2), len(c12))
    k["frobenius_coeff_fp6_c1"] = {"source": "src/fields/fp6.rs:40-108", "table": c1}
    k["frobenius_coeff_fp6_c2"] = {"source": "src/fields/fp6.rs:109-179", "table": c2}
    k["frobenius_coeff_fp12_c1"] = {"source": "src/fields/fp12.rs:29-172", "table": c12}

    # --- Fp2 constants
    w = word_arrays("src/fields/fp2.rs", 18, 55)
    k["two_inv"] = {"source": "src/fields/fp2.rs:18-23", "value": h(w[0])}
    k["fp2_twist_curve_constant"] = {"source": "src/fields/fp2.rs:42-55", "value": [h(w[3]), h(w[4])]}

    # --- G2 generator, psi constants (src/groups/g2.rs:47-112)
    w = word_arrays("src/groups/g2.rs", 47, 112)
    k["g2_generator"] = {"source": "src/groups/g2.rs:47-77", "x": [h(w[0]), h(w[1])], "y": [h(w[2]), h(w[3])]}
    k["eps_exp0"] = {"source": "src/groups/g2.rs:80-94", "value": [h(w[4]), h(w[5])]}
    k["eps_exp1"] = {"source": "src/groups/g2.rs:95-109", "value": [h(w[6]), h(w[7])]}
    k["bls_x"] = {"source": "src/groups/g2.rs:112", "value": h(w[8] & ((1 << 64) - 1))}

    # --- Gt generator e(G1,G2) (src/groups/gt.rs:20-109)
    w = word_arrays("src/groups/gt.rs", 20, 109)
    assert len(w) == 12
    k["gt_generator"] = {"source": "src/groups/gt.rs:20-109 (asserted at src/pairing.rs:1052-1057)",
                         "fp12": [h(x) for x in w]}

    # --- pairing test_cases (src/pairing.rs:1122-1189)
    d = decimals("src/pairing.rs", 1122, 1189)
    assert len(d) == 14
    k["pairing_test_cases"] = {"source": "src/pairing.rs:1122-1189", "g1_scalar": h(d[0]), "g2_scalar": h(d[1]),
                               "fp12": [h(x) for x in d[2:]]}

    # --- EIP-196 / EIP-197 vectors (examples/reth_bn128.rs)
    k["eip196_add"] = {"source": "examples/reth_bn128.rs:229-243",
                       "input": hexblob("examples/reth_bn128.rs", 230, 237),
                       "expected": hexblob("examples/reth_bn128.rs", 238, 243)}
    k["eip196_mul"] = {"source": "examples/reth_bn128.rs:311-324",
                       "input": hexblob("examples/reth_bn128.rs", 312, 318),
                       "expected": hexblob("examples/reth_bn128.rs", 319, 324)}
    k["eip197_pair"] = {"source": "examples/reth_bn128.rs:389-416",
                        "input": hexblob("examples/reth_bn128.rs", 390, 405),
                        "expected": "00" * 31 + "01"}
    assert len(k["eip196_add"]["input"]) == 256 and len(k["eip196_add"]["expected"]) == 128
    assert len(k["eip196_mul"]["input"]) == 192 and len(k["eip196_mul"]["expected"]) == 128
    assert len(k["eip197_pair"]["input"]) == 2 * 384

    # --- SvdW constants (src/svdw.rs:285-296)
    hx = re.findall(r'"([0-9a-f]{64})"', lines("src/svdw.rs", 285, 297))
    k["svdw_constants"] = {"source": "src/svdw.rs:285-296", "z": "0x1", "c1": "0x4",
                           "c2": "0x" + hx[0], "c3": "0x" + hx[1], "c4": "0x" + hx[2]}

    # --- RFC 9380 expand_message_xmd SHA-256 vectors (src/hasher.rs:372-388, 430-470)
    def xmd_map(lo, hi):
        txt = lines("src/hasher.rs", lo, hi)
        pairs = re.findall(r'm\.insert\("([^"]*)",\s*"([0-9a-f ]+)"\)', txt, re.S)
        return [[a, b] for a, b in pairs if " " not in b]

    dst_long = re.search(r'let dst = b"([^"]+)"', lines("src/hasher.rs", 455, 460)).group(1)
    k["xmd_sha256_short"] = {"source": "src/hasher.rs:367-376,434-438", "dst": "QUUX-V01-CS02-with-expander-SHA256-128",
                             "len_in_bytes": 32, "vectors": xmd_map(367, 377)}
    k["xmd_sha256_long_dst"] = {"source": "src/hasher.rs:378-388,456-458", "dst": dst_long, "len_in_bytes": 32,
                                "vectors": xmd_map(378, 389)}
    assert len(k["xmd_sha256_short"]["vectors"]) == 3 and len(k["xmd_sha256_long_dst"]["vectors"]) == 3

    # --- RFC 9380 expand_message_xof SHAKE128 vectors (src/hasher.rs:345-366, 393-428)
    xof_map = xmd_map

    xof_dst_long = re.search(r'let dst = b"([^"]+)"', lines("src/hasher.rs", 413, 418)).group(1)
    k["xof_shake128_short"] = {"source": "src/hasher.rs:345-355,394-398", "dst": "QUUX-V01-CS02-with-expander-SHAKE128",
                               "len_in_bytes": 32, "vectors": xof_map(345, 356)}
    k["xof_shake128_long_dst"] = {"source": "src/hasher.rs:357-366,413-417", "dst": xof_dst_long, "len_in_bytes": 32,
                                  "vectors": xof_map(357, 367)}
    assert len(k["xof_shake128_short"]["vectors"]) == 3 and len(k["xof_shake128_long_dst"]["vectors"]) == 3
    assert len(xof_dst_long) > 255

    # --- SvdW map vectors from the reference's own Sage specification (src/sage_reference/svdw.sage:1-137), evaluated
    #     by its plain-Python port oracle/svdw_sage.py (Sage is not in the image; the port restates only the field type).
    #     The constants the port DERIVES (find_z_svdw, c1..c4) must equal the Rust literals extracted above.
    import random
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from oracle import svdw_sage

    assert os.path.exists(os.path.join(REF, "src", "sage_reference", "svdw.sage"))
    s = svdw_sage.bn254_g1_svdw()
    c = k["svdw_constants"]
    assert (int(s.Z), int(s.c1), int(s.c2), int(s.c3), int(s.c4)) == tuple(int(c[n], 16) for n in ("z", "c1", "c2", "c3", "c4"))
    rng = random.Random(20261017)
    us = [0, 1, 2, svdw_sage.P - 1] + [int(z) for z in s.undefs] + [rng.randrange(svdw_sage.P) for _ in range(12)]
    k["svdw_map_vectors"] = {
        "source": "src/sage_reference/svdw.sage:87-137 (generic_svdw.map_to_point) via oracle/svdw_sage.py",
        "provenance": "second, independent pin of the SvdW map: the reference's Sage specification ported to plain "
                      "Python (Z searched by find_z_svdw, c1..c4 derived), cross-checked on 10^4 random inputs against "
                      "the restatement of src/svdw.rs by tests/test_svdw_second_pin.py",
        "cases": [["0x%x" % u, "0x%x" % s.map_to_point(u)[0], "0x%x" % s.map_to_point(u)[1]] for u in us]}

    with open(OUT, "w") as f:
        json.dump(k, f, indent=1, sort_keys=True)
    print("wrote", OUT, "with", len(k), "entries")


if __name__ == "__main__":
    main()
