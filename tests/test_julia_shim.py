"""The Julia binding (stheno.jl_b200/julia/SthenoB200.jl) cannot be executed here (no `julia`
binary in the image).  What can be checked mechanically, and is: every `ccall` names a symbol the
header declares and the built library exports, passes the declared number of arguments with
compatible C types, and every `struct Sb*` mirrors its C twin field for field."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "include", "stheno_b200.h")).read()
JL = open(os.path.join(ROOT, "stheno.jl_b200", "julia", "SthenoB200.jl")).read()

OPAQUE = {"sb_ctx", "sb_factor", "sb_vfe"}
STRUCTS = {"sb_array": "SbArray", "sb_term": "SbTerm", "sb_block": "SbBlock", "sb_covspec": "SbCovSpec",
           "sb_noise": "SbNoise", "sb_timings": "SbTimings"}


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", " ", s)


def c_kind(t):
    """C parameter / field type -> abstract kind."""
    t = t.replace("const", " ").strip()
    stars = t.count("*")
    base = t.replace("*", " ").split()[0]
    if stars == 0:
        return {"int32_t": "i32", "int64_t": "i64", "double": "f64"}[base]
    if base == "char":
        return "cstring"
    if base == "void" or base in OPAQUE:
        return "handle**" if stars == 2 else "void*"
    if base in STRUCTS:
        return f"struct*:{STRUCTS[base]}"
    return {"double": "f64*", "int64_t": "i64*", "int32_t": "i32*"}[base]


def jl_kind(t):
    t = t.strip()
    m = {"Int32": "i32", "Int64": "i64", "Float64": "f64", "Cdouble": "f64", "Cstring": "cstring"}
    if t in m:
        return m[t]
    mm = re.fullmatch(r"(Ptr|Ref)\{(.+)\}", t)
    assert mm, f"unrecognised Julia ccall type {t!r}"
    inner = mm.group(2).strip()
    if inner == "Cvoid":
        return "void*"
    if inner == "Ptr{Cvoid}":
        return "handle**"
    if inner in STRUCTS.values():
        return f"struct*:{inner}"
    return {"Float64": "f64*", "Int64": "i64*", "Int32": "i32*"}[inner]


def compatible(ck, jk):
    if ck == jk:
        return True
    # a typed Julia pointer may be passed where C takes void*, and vice versa for raw buffers
    if ck == "void*" and jk in ("void*", "f64*", "i64*", "i32*"):
        return True
    if jk == "void*" and ck in ("f64*",):
        return True
    return False


def header_protos():
    src = strip_comments(HDR)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int32_t|int64_t)\s+(sb_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        if args in ("void", ""):
            kinds = []
        else:
            kinds = []
            for a in args.split(","):
                a = " ".join(a.split())
                a = re.sub(r"\b\w+$", "", a).strip() if not a.endswith("*") else a  # drop the parameter name
                kinds.append(c_kind(a))
        protos[name] = ("cstring" if "char" in ret else c_kind(ret), kinds)
    return protos


def split_top(s):
    """split on commas not nested in {} or ()"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def julia_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*LIB\),\s*(\w+),\s*\(", JL):
        name, ret = m.group(1), m.group(2)
        i = m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(JL[j], 0)
            j += 1
        argt = JL[i:j - 1]
        types = [t for t in split_top(argt) if t]
        # actual arguments: up to the matching close of the ccall
        k = j
        if JL[k] == ")":          # no arguments: ccall((:f, LIB), Ret, ())
            actual = []
        else:
            assert JL[k] == ",", (name, JL[k - 20:k + 20])
            depth, e = 1, k
            while depth:
                e += 1
                depth += {"(": 1, ")": -1}.get(JL[e], 0)
            actual = split_top(JL[k + 1:e])
        calls.append((name, ret, types, actual))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_protos()
    assert len(protos) >= 30
    calls = julia_ccalls()
    assert len(calls) >= 20
    for name, ret, types, actual in calls:
        assert name in protos, f"ccall to undeclared symbol {name}"
        cret, ckinds = protos[name]
        assert jl_kind(ret) == cret, (name, ret, cret)
        assert len(types) == len(ckinds), f"{name}: Julia passes {len(types)} argument types, C declares {len(ckinds)}"
        assert len(actual) == len(types), f"{name}: {len(actual)} actual arguments for {len(types)} types"
        for pos, (jt, ck) in enumerate(zip(types, ckinds)):
            assert compatible(ck, jl_kind(jt)), f"{name} arg {pos}: Julia {jt} vs C {ck}"


def test_shim_covers_the_overridden_entry_points():
    used = {c[0] for c in julia_ccalls()}
    need = {"sb_ctx_create", "sb_ctx_destroy", "sb_last_error", "sb_cov_dense", "sb_cov_diag", "sb_factor_create",
            "sb_factor_destroy", "sb_logpdf", "sb_factor_set_data", "sb_factor_alpha", "sb_factor_set_alpha",
            "sb_predict", "sb_predict_cov", "sb_predict_factor", "sb_rand", "sb_vfe_create", "sb_vfe_predict",
            "sb_vfe_predict_cov", "sb_vfe_destroy", "sb_ctx_timings", "sb_ctx_set_option", "sb_logpdf_grad",
            "sb_factor_export_size", "sb_factor_export", "sb_factor_import"}
    assert need <= used, need - used
    # the methods a Stheno user calls on this path
    for sig in ["logpdf(fx::B200Finite", "posterior(fx::B200Finite", "elbo(v::AbstractGPs.VFE", "posterior(v::AbstractGPs.VFE",
                "logpdf(f::B200Sparse", "posterior(f::B200Sparse", "marginals(fx::B200Finite", "cov(fp::B200PosteriorGP",
                "rand(rng::AbstractRNG, fx::B200PostFinite", "AbstractVector{<:Tuple{Symbol,Any}}"]:
        assert sig in JL, sig


def c_struct_fields(cname):
    src = strip_comments(HDR)
    m = re.search(r"typedef\s+struct\s*\{([^}]*)\}\s*" + cname + r"\s*;", src, flags=re.S)
    assert m, cname
    fields = []
    for decl in m.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        names = [n.strip() for n in decl.split(",")]
        first = names[0]
        mm = re.match(r"(.*?)(\w+)$", first)
        typ, n0 = mm.group(1).strip(), mm.group(2)
        for nm in [n0] + names[1:]:
            fields.append((nm.strip(), c_kind(typ)))
    return fields


def jl_struct_fields(jname):
    m = re.search(r"struct\s+" + jname + r"\b(.*?)\bend\b", JL, flags=re.S)
    assert m, jname
    fields = []
    for part in re.split(r"[;\n]", m.group(1)):
        part = part.strip()
        if "::" in part:
            nm, t = part.split("::")
            fields.append((nm.strip(), jl_kind(t.strip())))
    return fields


def test_struct_layouts_match():
    for cname, jname in STRUCTS.items():
        cf, jf = c_struct_fields(cname), jl_struct_fields(jname)
        assert [n for n, _ in cf] == [n for n, _ in jf], (cname, cf, jf)
        for (n, ck), (_, jk) in zip(cf, jf):
            assert compatible(ck, jk) or (ck.startswith("struct*") and jk.startswith("struct*")), (cname, n, ck, jk)


def test_diag_spec_lists_only_paired_blocks():
    """Round-1 bug: the diag spec listed every (i, j) block; var of a multi-block input was clobbered."""
    assert "which === :diag && j != i && continue" in JL
    assert JL.count("which=:diag") >= 3
