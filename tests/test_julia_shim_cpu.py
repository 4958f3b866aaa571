"""The Julia shim cannot be executed here (no Julia toolchain).  What CAN be checked without it: every `ccall` names a
symbol that include/azb200.h declares and libazb200.so exports, passes exactly as many arguments as the C prototype has
parameters and declares as many argument types; every C struct mirrors the header field for field; block openers and
`end`s balance; every module the shim uses is imported."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "alphazero.jl_b200", "julia", "AlphaZeroB200.jl")
HDR = os.path.join(ROOT, "include", "azb200.h")


def _strip(src):
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)       # string literals
    return "\n".join(l.split("#", 1)[0] for l in src.splitlines())


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def _balanced(s, start):
    """index just after the parenthesis that closes the one at s[start]"""
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise AssertionError("unbalanced parentheses")


def _prototypes():
    h = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int32_t|int64_t|const char\*)\s+(az_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    return protos, h


def test_ccalls_match_the_header_and_the_library():
    import ctypes
    import _pkg
    _pkg.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "alphazero.jl_b200", "libazb200.so"))
    protos, _ = _prototypes()
    src = _strip(open(SHIM).read())
    calls = list(re.finditer(r"ccall\(\(:(\w+), LIB\)", src))
    assert len(calls) >= 20
    for m in calls:
        name = m.group(1)
        assert name in protos, name + " is not declared in include/azb200.h"
        assert hasattr(lib, name), name + " is not exported by libazb200.so"
        body = src[m.start() + len("ccall"):_balanced(src, m.start() + len("ccall"))][1:-1]
        parts = _split_top(body)          # (:sym, LIB), rettype, (argtypes...), args...
        argtypes = _split_top(parts[2].strip()[1:-1])
        argtypes = [a for a in argtypes if a]
        assert len(argtypes) == protos[name], (name, argtypes, protos[name])
        assert len(parts) - 3 == protos[name], (name, parts[3:], protos[name])


def test_c_structs_mirror_the_header():
    _, h = _prototypes()
    src = open(SHIM).read()

    def c_fields(name):
        body = dict((n, b) for b, n in re.findall(r"typedef struct \{([^{}]*)\}\s*(\w+)\s*;", h))[name]
        return [re.sub(r"\[.*?\]", "", f.strip().split()[-1]) for f in body.split(";") if f.strip()]

    def jl_fields(name):
        body = re.search(r"struct " + name + r"\b(.*?)\nend", src, flags=re.S).group(1)
        return [re.match(r"\s*(\w+)::", l).group(1) for l in body.splitlines() if re.match(r"\s*\w+::", l)]
    assert jl_fields("CMctsParams") == c_fields("az_mcts_params")
    assert jl_fields("CSimParams") == c_fields("az_sim_params")
    assert jl_fields("CResNetHP") == c_fields("az_resnet_hp")
    assert jl_fields("CSimpleNetHP") == c_fields("az_simplenet_hp")


def test_blocks_balance_and_modules_are_imported():
    src = _strip(open(SHIM).read())
    # keywords at bracket depth 0 only: `for` / `if` inside [...] or (...) are comprehensions / generators, `end` inside
    # [...] would be an index (the shim avoids it)
    depth_at, d = [], 0
    for ch in src:
        depth_at.append(d)
        if ch in "([{":
            d += 1
        elif ch in ")]}":
            d -= 1
    assert d == 0, "unbalanced brackets"
    toks = [m.group(1) for m in re.finditer(r"(?<![\w.:@])(module|function|struct|if|for|while|let|begin|do|try|quote|macro|end)(?![\w!?])", src)
            if depth_at[m.start()] == 0]
    opens = sum(1 for t in toks if t != "end")
    assert opens == toks.count("end"), (opens, toks.count("end"))
    for mod in ("Flux", "JSON3", "Distributed"):
        assert re.search(r"^\s*(import|using)\s+" + mod + r"\b", open(SHIM).read(), flags=re.M), mod + " is used but not imported"
    assert "Examples" in re.search(r"using AlphaZero:(.*?)\n\S", open(SHIM).read(), flags=re.S).group(1)
    # the seam methods exist for simulate and simulate_distributed
    assert "function AlphaZero.simulate_distributed(simulator::Simulator, gspec::$S, p::SimParams; game_simulated)" in open(SHIM).read()
    assert "function AlphaZero.simulate(simulator::Simulator, gspec::$S, p::SimParams; game_simulated)" in open(SHIM).read()
