"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/integrals_b200.h declares, and fails loudly (no fallback) when there is no GPU."""
import os
import re

import numpy as np
import pytest

import integrals_b200 as ib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "integrals_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ib200_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported():
    L = ib.load_library()
    syms = _header_symbols()
    assert len(syms) >= 17
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(ib.SYMBOLS) == syms, "python binding list and header disagree"


def test_version_and_family_table():
    L = ib.load_library()
    assert L.ib200_version() == 100
    assert L.ib200_family_nparam(ib.FAMILY_IDS["sinxp"], 1) == 1
    assert L.ib200_family_nparam(ib.FAMILY_IDS["sinxp"], 2) == -1      # 1-D only
    assert L.ib200_family_nparam(ib.FAMILY_IDS["genz_gaussian"], 4) == 8
    assert L.ib200_family_nparam(ib.FAMILY_IDS["gausspoly"], 2) == 7
    assert L.ib200_family_nparam(99, 1) == -1
    assert L.ib200_family_nparam(1, 9) == -1


def test_family_table_matches_oracle():
    import oracle as O
    L = ib.load_library()
    for name, fid in ib.FAMILY_IDS.items():
        for d in range(1, 9):
            got = L.ib200_family_nparam(fid, d)
            if got >= 0:
                assert got == O.family_nparam(fid, d), (name, d)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ib.IB200Error, match="no CUDA device|no CPU fallback"):
        ib.Engine(0)


def test_ctypes_signatures_match_header():
    """every prototype in include/integrals_b200.h has the same number of arguments as the ctypes binding declares, and
    pointer / integer / floating arguments sit at the same positions"""
    import ctypes as C
    import re
    import integrals_b200 as ib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "include", "integrals_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    L = ib.load_library()
    checked = 0
    for m in re.finditer(r"\b(?:int32_t|int64_t|double|void|const char \*)\s*\*?\s*(ib200_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        assert len(params) == len(fn.argtypes), (name, params, fn.argtypes)
        for par, ct in zip(params, fn.argtypes):
            is_ptr = "*" in par
            if is_ptr:
                assert ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "contents") or ct.__name__.startswith("LP_"), (name, par, ct)
            elif par.startswith(("double", "float")):
                assert ct in (C.c_double, C.c_float), (name, par, ct)
            else:
                assert ct in (C.c_int32, C.c_int64, C.c_int), (name, par, ct)
        checked += 1
    assert checked >= 15


def test_device_rule_tables_equal_the_pinned_tables():
    """the Gauss-Kronrod (7,15) tables compiled into the CUDA kernels (csrc/quadgk.cu, quadgk_vec.cu, hcubature_impl.cuh) and into the
    oracle are the same doubles as kronrod.kronrod(7) -- the tables pinned against QuadGK's published example and SciPy's table
    (tests/test_oracle_pins.py); and every kernel file spells the Genz-Malik weights with the same integers (Genz & Malik 1980)"""
    import re
    from integrals_b200.kronrod import kronrod
    x, w, wg = kronrod(7)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    num = r"[-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?"

    def tables(path, prefix, qual=r"(?:__constant__|static const) double"):
        src = open(os.path.join(root, path)).read()
        out = []
        for name in ("x", "w", "wg"):
            m = re.search(qual + r"\s+" + prefix + name + r"\[\d+\]\s*=\s*\{([^}]*)\}", src, re.I)
            assert m, (path, prefix + name)
            out.append(np.array([float(v) for v in re.findall(num, m.group(1))]))
        return out
    for path, prefix in (("integrals.jl_b200/csrc/quadgk.cu", "c_gk_"), ("integrals.jl_b200/csrc/quadgk_vec.cu", "v_gk_"),
                         ("integrals.jl_b200/csrc/hcubature_impl.cuh", "c_"), ("oracle/oracle.c", "GK_")):
        tx, tw, twg = tables(path, prefix)
        assert np.array_equal(tx, x) and np.array_equal(tw, w) and np.array_equal(twg, wg), path
    gm = ["12824 - 9120 * D + 400 * D * D", "980 / 6561.0", "1820 - 400 * D", "200 / 19683.0", "6859 / 19683.0",
          "729 - 950 * D + 50 * D * D", "245 / 486.0", "265 - 100 * D", "25 / 729.0"]
    for path in ("hcubature_impl.cuh", "hcubature_pair.cuh", "hcubature_vec.cuh", "hcub_rounds.cuh", "batch_bridge.cu"):
        src = open(os.path.join(root, "integrals.jl_b200", "csrc", path)).read()
        for term in gm:
            assert term in src, (path, term)
    osrc = open(os.path.join(root, "oracle", "oracle.c")).read()
    for term in gm:
        assert term.replace("D", "n") in osrc, term


def test_every_entry_point_rejects_a_null_context():
    """error behaviour of the boundary (SURVEY 8(b) "Errors"): no entry point dereferences a NULL context -- each returns
    IB200_ERR_INVALID and leaves a message naming itself for ib200_last_error(NULL).  (Argument validation only: nothing here needs a GPU.)"""
    import ctypes as C
    L = ib.load_library()
    no_ctx = {"ib200_version", "ib200_family_nparam", "ib200_last_error", "ib200_comm_unique_id", "ib200_create", "ib200_destroy",
              "ib200_kernel_launches", "ib200_last_kernel_ms",
              "ib200_integrand_compile", "ib200_integrand_release"}      # context-free: NVRTC needs no device (tests/test_user_integrand.py)
    checked = 0
    for name in ib.SYMBOLS:
        if name in no_ctx:
            continue
        fn = getattr(L, name)
        args = [0.0 if t in (C.c_double, C.c_float) else (0 if t in (C.c_int32, C.c_int64) else None) for t in fn.argtypes]
        assert fn(*args) == -1, name
        msg = L.ib200_last_error(None).decode()
        assert msg.startswith(name if name != "ib200_sampled_f32" else "ib200_sampled") and "NULL" in msg, (name, msg)
        checked += 1
    assert checked >= 20
    assert L.ib200_kernel_launches(None) == 0
    L.ib200_destroy(None)                                   # a no-op


def test_julia_binding_files_match_the_document():
    """julia/algorithms_b200.jl and julia/ext/IntegralsB200Ext.jl are INTEGRATION.md's code blocks written out as source, and every
    `ccall((:name, lib), ...)` in them names a symbol the header declares, with the argument count of its prototype"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("extract_julia_binding", os.path.join(ROOT, "scripts", "extract_julia_binding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for rel, src in mod.render().items():
        assert open(os.path.join(ROOT, rel)).read() == src, f"{rel} is stale: run scripts/extract_julia_binding.py"
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "integrals_b200.h")).read(), flags=re.S)
    nargs, kinds = {}, {}
    for m in re.finditer(r"\b(ib200_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        a = m.group(2).strip()
        nargs[m.group(1)] = 0 if a in ("", "void") else a.count(",") + 1
        kinds[m.group(1)] = "" if a in ("", "void") else "".join(
            "p" if "*" in t else ("f" if re.search(r"\b(double|float)\b", t) else "i") for t in a.split(","))
    jl = open(os.path.join(ROOT, "julia", "ext", "IntegralsB200Ext.jl")).read()
    jl = "\n".join(line for line in jl.splitlines() if not line.lstrip().startswith("#"))      # comment lines sketch calls with "..."
    seen = 0
    for m in re.finditer(r"ccall\(\(:(ib200_\w+), lib\),\s*\w+,\s*\(", jl):
        name = m.group(1)
        assert name in nargs, name
        i = m.end(); depth = 1; j = i
        while depth:                                       # the argument-type tuple
            depth += {"(": 1, ")": -1}.get(jl[j], 0); j += 1
        types = re.sub(r"#=.*?=#", "", jl[i:j - 1], flags=re.S)
        tl = [t.strip() for t in types.split(",") if t.strip()]
        assert len(tl) == nargs[name], (name, len(tl), nargs[name])
        got = "".join("p" if t.startswith(("Ptr{", "Ref{", "Cstring")) else ("f" if t in ("Float64", "Float32") else "i") for t in tl)
        assert got == kinds[name], (name, got, kinds[name])       # pointers, integers and floats at the same positions
        seen += 1
    assert seen >= 10


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "src")), reason="the reference tree is only present in the build container")
def test_julia_extension_uses_only_names_the_reference_defines():
    """grep-level check of the Julia extension against /root/reference/src: every `Integrals.<name>` it extends or calls is defined there,
    every field it reads off IntegralCache / SampledIntegralCache / the problem types exists in that struct, the sampled algorithms get a
    find_weights method (src/sampled.jl:128-133 calls it inside init) and the solution is built from build_problem(cache)
    (SampledIntegralCache has no `prob` field, src/common.jl:264-274)."""
    src = ""
    for dp, _, fs in os.walk(os.path.join(REFERENCE, "src")):
        for f in fs:
            if f.endswith(".jl"):
                src += open(os.path.join(dp, f)).read() + "\n"
    jl = open(os.path.join(ROOT, "julia", "ext", "IntegralsB200Ext.jl")).read()
    code = "\n".join(line.split(" #")[0] for line in jl.splitlines() if not line.lstrip().startswith("#"))
    # functions of the reference the extension extends / calls
    for name in sorted(set(re.findall(r"\bIntegrals\.(\w+)", code))):
        assert re.search(r"\bfunction\s+" + name + r"\b|^\s*" + name + r"\(.*\)\s*=", src, flags=re.M) or re.search(r"\b(struct|abstract type)\s+" + name + r"\b", src), name
    for name in ("init_cacheval", "__solvebp_call", "__solve", "find_weights", "build_problem", "dimension"):
        assert re.search(r"\bIntegrals\." + name + r"\b", code), f"the extension never touches Integrals.{name}"

    def fields(struct):
        m = re.search(r"mutable struct " + struct + r"\b[^\n]*\n(.*?)\nend", src, flags=re.S)
        assert m, struct
        return set(re.findall(r"^\s*(\w+)::", m.group(1), flags=re.M))
    sampled_fields, cache_fields = fields("SampledIntegralCache"), fields("IntegralCache")
    assert {"y", "x", "dim", "isfresh", "cacheval"} <= sampled_fields and "prob" not in sampled_fields
    # the sampled method only reads fields SampledIntegralCache has
    m = re.search(r"function Integrals\.__solvebp_call\(cache::SampledIntegralCache.*?\nend\n", code, flags=re.S)
    assert m
    used = set(re.findall(r"\bcache\.(\w+)", m.group(0)))
    assert used and used <= sampled_fields, used - sampled_fields
    # ... and the IntegralCache methods only fields IntegralCache has
    rest = code.replace(m.group(0), "")
    used = set(re.findall(r"\bcache\.(\w+)", rest))
    assert used and used <= cache_fields, used - cache_fields
    # the defects of the first version stay fixed: weights through find_weights, solution through build_problem, the grid copy rooted
    assert "Integrals.find_weights(x::AbstractVector, alg::Union{TrapezoidalB200, SimpsonsB200})" in code
    assert "cache.prob" not in code and "pointer(collect(" not in code
    assert re.search(r"GC\.@preserve y out xs", code)
    # every reference type / function imported by name exists
    imp = re.search(r"using Integrals:(.*?)\nconst lib", jl, flags=re.S).group(1)
    new_names = {"AbstractB200Algorithm", "QuadGKB200", "HCubatureB200", "QuadratureRuleB200", "GaussLegendreB200", "TrapezoidalB200",
                 "SimpsonsB200", "family_id", "VectorOf"}                         # defined by julia/algorithms_b200.jl
    alg_src = open(os.path.join(ROOT, "julia", "algorithms_b200.jl")).read()
    for name in [n.strip() for n in imp.replace("\n", " ").split(",") if n.strip()]:
        if name in new_names:
            assert re.search(r"\b(struct|abstract type)\s+" + name + r"\b|^" + name + r"\(", alg_src, flags=re.M), name
        else:
            assert re.search(r"\b(struct|function|abstract type)\s+" + name + r"\b", src), name


def test_every_option_key_the_library_accepts_is_documented_in_the_header():
    """ib200_set_option: the keys handled in csrc/ctx.cu and the list in include/integrals_b200.h stay in step"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "integrals.jl_b200", "csrc", "ctx.cu")).read()
    body = src[src.index("ib200_set_option(ib200_ctx *ctx"):]
    body = body[:body.index("unknown key")]
    keys = set(re.findall(r'k == "([a-z_0-9]+)"', body))
    assert {"hcub_mode", "quadgk_mode", "hcub_compact", "hcub_compact_level", "single_sub_boxes"} <= keys
    header = open(os.path.join(root, "include", "integrals_b200.h")).read()
    missing = sorted(k for k in keys if f'"{k}"' not in header)
    assert not missing, missing
