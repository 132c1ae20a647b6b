"""Static check: every `call("gcc_...", ...)` in the Python host code passes exactly the number of arguments the
header declares (ctypes would only report a mismatch at run time on the GPU box)."""
import ast
import os

from gcc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _calls(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and node.args and isinstance(node.args[0], ast.Constant) \
                and isinstance(node.args[0].value, str) and node.args[0].value.startswith("gcc_"):
            fname = getattr(node.func, "id", getattr(node.func, "attr", ""))
            if fname == "call":
                yield node.args[0].value, len(node.args) - 1, node.lineno


def test_call_sites_match_header():
    protos = _lib.parse_header()
    files = [os.path.join(ROOT, "gcc_b200", f) for f in os.listdir(os.path.join(ROOT, "gcc_b200")) if f.endswith(".py")]
    files += [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    files += [os.path.join(ROOT, "scripts", f) for f in os.listdir(os.path.join(ROOT, "scripts")) if f.endswith(".py")]
    seen = 0
    for path in files:
        for name, nargs, line in _calls(path):
            assert name in protos, "%s:%d calls undeclared %s" % (path, line, name)
            assert nargs == len(protos[name][1]), "%s:%d %s passes %d args, header declares %d" % (
                path, line, name, nargs, len(protos[name][1]))
            seen += 1
    assert seen > 40
