"""Static consistency checks that need no GPU: most product code below the C ABI only ever runs on a GPU box, so a
misspelt name or a drifted argument list would otherwise surface there for the first time."""
import ast
import builtins
import os
import re
import symtable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files():
    out = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for top in ("lasr_b200", "oracle", "profiles"):
        for d, _, fs in os.walk(os.path.join(ROOT, top)):
            if "__pycache__" in d or os.sep + "_ref" in d:
                continue
            out += [os.path.join(d, f) for f in fs if f.endswith(".py")]
    return out


def test_no_undefined_names_in_product_bench_and_oracle_code():
    bad = []
    for path in _py_files():
        src = open(path).read()
        top = symtable.symtable(src, path, "exec")
        known = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
        for node in ast.walk(ast.parse(src)):
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                known |= {(a.asname or a.name).split(".")[0] for a in node.names}

        def globals_assigned(t):
            for s in t.get_symbols():
                if s.is_global() and s.is_assigned():
                    known.add(s.get_name())
            for c in t.get_children():
                globals_assigned(c)
        globals_assigned(top)

        def visit(t):
            for s in t.get_symbols():
                n = s.get_name()
                if s.is_referenced() and s.is_global() and n not in known and not hasattr(builtins, n) \
                        and n not in ("__file__", "__name__", "__doc__", "__class__"):
                    bad.append((os.path.relpath(path, ROOT), t.get_name(), t.get_lineno(), n))
            for c in t.get_children():
                visit(c)
        visit(top)
    assert not bad, bad


def _header_arities():
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lasr_b200.h")).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"\b(lasr_b200_\w+)\s*\(([^)]*)\)\s*;", src):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def _header_kinds():
    """name -> parameter kinds ('ptr' | 'int' | 'float' | 'size') of every prototype in include/lasr_b200.h."""
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lasr_b200.h")).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)

    def kind(p):
        if "*" in p:
            return "ptr"
        t = re.sub(r"\b\w+$", "", p.strip()).replace("const", "").strip() or p.strip()
        return {"int": "int", "int32_t": "int", "float": "float", "size_t": "size", "cudaStream_t": "ptr"}[t]
    out = {}
    for m in re.finditer(r"\b(lasr_b200_\w+)\s*\(([^)]*)\)\s*;", src):
        params = m.group(2).strip()
        out[m.group(1)] = [] if params in ("", "void") else [kind(p) for p in params.split(",")]
    return out


def test_ctypes_signatures_follow_the_header_parameter_by_parameter():
    """Every prototype of include/lasr_b200.h against the argtypes lasr_b200/_cabi.py declares for it: same number of
    parameters, and a pointer / int / float / size_t in the same positions."""
    import ctypes
    from lasr_b200 import _cabi
    L = _cabi.lib()
    assert {k: len(v) for k, v in _header_kinds().items()} == _header_arities()

    def ckind(t):
        if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") or (hasattr(t, "_type_") and not isinstance(t._type_, str)):
            return "ptr"
        return {ctypes.c_int: "int", ctypes.c_float: "float", ctypes.c_size_t: "size"}[t]
    kinds = _header_kinds()
    assert set(kinds) == set(_cabi.EXPORTED_SYMBOLS)
    for name, want in kinds.items():
        fn = getattr(L, name)
        have = [ckind(t) for t in (fn.argtypes or [])]
        assert have == want, (name, "header", want, "ctypes", have)


def test_native_calls_pass_as_many_arguments_as_the_entry_point_takes():
    """`_cabi.native_call("name", ...)` sites (the stream is appended by native_call) and `_cabi.<attr>` references."""
    from lasr_b200 import _cabi
    L = _cabi.lib()
    sites = 0
    for path in _py_files():
        if os.sep + "lasr_b200" + os.sep not in path:
            continue
        for node in ast.walk(ast.parse(open(path).read())):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id == "_cabi":
                assert hasattr(_cabi, node.attr), (path, node.lineno, node.attr)
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in ("native_call", "meshreg_call") \
                    and node.args and isinstance(node.args[0], ast.Constant):
                fn = getattr(L, "lasr_b200_" + node.args[0].value)
                assert not any(isinstance(a, ast.Starred) for a in node.args)
                assert len(fn.argtypes) == len(node.args), (path, node.lineno, node.args[0].value)
                sites += 1
    assert sites >= 11
