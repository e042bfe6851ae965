"""CPU tier: static consistency of the Rust boundary (ffi/).  There is no Rust toolchain in the image, so nothing here
compiles Rust; what CAN be checked is checked: the fork-side wrapper (ffi/sylow-batch/src/batch.rs) defines every name
its `pub use` list exports, only calls symbols that sylow-cuda-sys declares, and passes each as many arguments as the
declaration (generated from include/sylow_b200.h) takes."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BATCH = os.path.join(ROOT, "ffi", "sylow-batch", "src", "batch.rs")
SYS = os.path.join(ROOT, "ffi", "sylow-cuda-sys", "src", "lib.rs")


def _strip_comments(src: str) -> str:
    return re.sub(r"//[^\n]*", "", src)


def _split_top_level(args: str):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([{<":
            depth += 1
        elif ch in ")]}>":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [a.strip() for a in out]


def _call_args(src: str, start: int):
    """arguments of the call whose opening parenthesis is at src[start]"""
    depth, i = 0, start
    while True:
        if src[i] == "(":
            depth += 1
        elif src[i] == ")":
            depth -= 1
            if depth == 0:
                return _split_top_level(src[start + 1:i])
        i += 1


def _sys_decls():
    src = _strip_comments(open(SYS).read())
    fns = {}
    for m in re.finditer(r"pub fn (sylow_b200_\w+)\s*\(", src):
        fns[m.group(1)] = len(_call_args(src, m.end() - 1))
    consts = set(re.findall(r"pub const (SYLOW_B200_\w+)", src))
    return fns, consts


def test_pub_use_list_is_defined():
    raw = open(BATCH).read()
    m = re.search(r"pub use crate::batch::\{([^}]*)\}", raw.replace("//!", ""))
    assert m, "batch.rs must document the `pub use` list lib.rs adds"
    names = [n.strip() for n in m.group(1).replace("\n", " ").split(",") if n.strip()]
    assert len(names) >= 12
    code = _strip_comments(raw)
    for n in names:
        assert re.search(r"pub (fn|struct|enum) %s\b" % re.escape(n), code), "batch.rs exports `%s` but does not define it" % n
    # the entry points BASELINE's north_star and VERDICT r1 name
    for need in ("pairing_batch", "verify_batch", "g1_mul_batch", "g2_mul_batch", "pairing_check_batch", "sign_batch",
                 "verify_each", "final_exponentiation_batch", "glued_pairing_batch", "Engine", "BatchError"):
        assert need in names, need


def test_sys_symbols_exist_and_arities_match():
    fns, consts = _sys_decls()
    assert len(fns) >= 55
    code = _strip_comments(open(BATCH).read())
    used = set()
    for m in re.finditer(r"sys::(sylow_b200_\w+)\s*\(", code):
        name = m.group(1)
        used.add(name)
        assert name in fns, "batch.rs calls sys::%s, which sylow-cuda-sys does not declare" % name
        got = len(_call_args(code, m.end() - 1))
        assert got == fns[name], "sys::%s takes %d arguments, batch.rs passes %d" % (name, fns[name], got)
    for c in set(re.findall(r"sys::(SYLOW_B200_\w+)", code)):
        assert c in consts, c
    # one context for all GPUs: the wrapper must go through the in-library multi-device path
    assert "sylow_b200_create_multi" in used and "sylow_b200_destroy" in used
    assert "std::thread" not in code, "slicing over GPUs lives behind the C ABI, not in the wrapper"


def test_error_enum_carries_cuda_code_and_maps_every_status():
    code = _strip_comments(open(BATCH).read())
    assert re.search(r"Cuda\s*\{\s*code:\s*i32\s*\}", code)
    _, consts = _sys_decls()
    for c in consts:
        if c.startswith("SYLOW_B200_ERR_") and c != "SYLOW_B200_ERR_ARG":
            assert "sys::%s" % c in code, "status %s is not mapped" % c


def test_integration_doc_names_the_wrapper_functions():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = _strip_comments(open(BATCH).read())
    for fn in re.findall(r"pub fn (\w+)\(e: &mut Engine", code):
        assert fn in doc, "INTEGRATION.md does not mention %s" % fn
