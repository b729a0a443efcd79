"""julia/MamboB200.jl cannot be executed here (no `julia`), so its `ccall` sites and struct mirrors are checked textually against
include/mambo_csa.h: every called symbol is declared, the argument count matches the prototype, every argument has the same C
category (pointer / int / unsigned / double / long long), the return type matches, and the Julia struct mirrors list the header's
fields in the header's order with matching types."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mambo_csa.h")
SHIM = os.path.join(ROOT, "julia", "MamboB200.jl")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _c_category(t):
    t = t.strip()
    if "*" in t:
        return "ptr"
    t = re.sub(r"\bconst\b", "", t).strip()
    t = re.sub(r"\s+[A-Za-z_][A-Za-z_0-9]*$", "", t).strip() if " " in t else t       # drop the parameter name
    return {"int": "int", "unsigned": "uint", "unsigned int": "uint", "double": "double", "long long": "longlong", "void": "void"}[t]


def _jl_category(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring",):
        return "ptr"
    return {"Cint": "int", "Cuint": "uint", "Cdouble": "double", "Clonglong": "longlong", "Cvoid": "void"}[t]


def header_prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(const char\s*\*|int|void|double)\s+(mambo_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        cats = [] if args in ("", "void") else [_c_category(a) for a in _split_top(args)]
        protos[name] = ("ptr" if "*" in ret else _c_category(ret), cats)
    return protos


def shim_calls():
    src = open(SHIM).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(mambo_[a-z_0-9]+), LIB\),\s*([A-Za-z{}]+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        sig = src[m.end():i - 1]
        calls.append((m.group(1), _jl_category(m.group(2)), [_jl_category(a) for a in _split_top(sig)] if sig.strip() else []))
    return calls


def test_every_ccall_matches_its_prototype():
    protos = header_prototypes()
    calls = shim_calls()
    assert len(calls) >= 18
    for name, ret, cats in calls:
        assert name in protos, f"{name} is called from the Julia shim but not declared in include/mambo_csa.h"
        pret, pcats = protos[name]
        assert ret == pret, f"{name}: return type {ret} vs {pret}"
        assert cats == pcats, f"{name}: argument categories {cats} in the shim vs {pcats} in the header"


def _header_struct(name):
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct " + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        typ, names = decl.split(" ", 1)
        fields += [(n.strip(), {"int": "Cint", "double": "Cdouble", "unsigned": "Cuint"}[typ]) for n in names.split(",")]
    return fields


def _shim_struct(name):
    src = open(SHIM).read()
    body = re.search(r"^struct " + name + r"\b[^\n]*\n(.*?)^end", src, flags=re.S | re.M).group(1)
    fields = []
    for part in re.split(r"[;\n]", body):
        part = part.split("#")[0].strip()
        if part:
            n, t = part.split("::")
            fields.append((n.strip(), t.strip()))
    return fields


def test_struct_mirrors_follow_the_header():
    for jl, c in (("OptOptions", "mambo_opt_options_t"), ("OptResult", "mambo_opt_result_t"), ("DfStats", "mambo_df_stats_t")):
        assert _shim_struct(jl) == _header_struct(c), (jl, c)
