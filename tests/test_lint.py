"""Undefined-name lint over the Python sources (pyflakes is not in the image).  Round 1's GPU suite
went red on a `res[K]` typo that only a GPU run would have executed; this catches that class on CPU."""
import ast
import builtins
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files():
    out = []
    for top in ("tests", "gtn_b200", "oracle", "scripts"):
        for d, _, fs in os.walk(os.path.join(ROOT, top)):
            if "__pycache__" in d:
                continue
            out += [os.path.join(d, f) for f in fs if f.endswith(".py")]
    out += [os.path.join(ROOT, f) for f in ("bench.py", "__graft_entry__.py")]
    return sorted(out)


class _Scopes(ast.NodeVisitor):
    """Collects every name bound anywhere in a function (or module / class) body, then checks that
    every Name load resolves in some enclosing scope or builtins.  Flow-insensitive on purpose."""

    def __init__(self):
        self.stack = [set(dir(builtins)) | {"__file__", "__name__", "__doc__"}]
        self.errors = []

    @staticmethod
    def _bound(node):
        names = set()

        def targets(t):
            for n in ast.walk(t):
                if isinstance(n, ast.Name):
                    names.add(n.id)

        class V(ast.NodeVisitor):
            def visit_FunctionDef(s, n):
                names.add(n.name)

            visit_AsyncFunctionDef = visit_FunctionDef

            def visit_ClassDef(s, n):
                names.add(n.name)

            def visit_Lambda(s, n):
                pass

            def visit_Import(s, n):
                for a in n.names:
                    names.add((a.asname or a.name).split(".")[0])

            def visit_ImportFrom(s, n):
                for a in n.names:
                    names.add(a.asname or a.name)

            def visit_Global(s, n):
                names.update(n.names)

            visit_Nonlocal = visit_Global

            def visit_Name(s, n):
                if isinstance(n.ctx, (ast.Store, ast.Del)):
                    names.add(n.id)

            def visit_ExceptHandler(s, n):
                if n.name:
                    names.add(n.name)
                s.generic_visit(n)

            def visit_MatchAs(s, n):
                if n.name:
                    names.add(n.name)
                s.generic_visit(n)

            # comprehension targets live in their own scope but treating them as bound here is a
            # harmless over-approximation
        v = V()
        body = node.body if isinstance(node.body, list) else [node.body]
        for st in body:
            v.visit(st)
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            a = node.args
            for arg in a.posonlyargs + a.args + a.kwonlyargs + [a.vararg, a.kwarg]:
                if arg is not None:
                    names.add(arg.arg)
        return names

    def _scope(self, node):
        self.stack.append(self._bound(node))
        self.generic_visit(node)
        self.stack.pop()

    def visit_Module(self, node):
        self._scope(node)

    def visit_FunctionDef(self, node):
        for d in node.decorator_list + node.args.defaults + [k for k in node.args.kw_defaults if k]:
            self.visit(d)
        self._scope(node)

    visit_AsyncFunctionDef = visit_FunctionDef

    def visit_Lambda(self, node):
        self._scope(node)

    def visit_ClassDef(self, node):
        self._scope(node)

    def visit_Name(self, node):
        if isinstance(node.ctx, ast.Load) and not any(node.id in s for s in self.stack):
            self.errors.append((node.lineno, node.id))


@pytest.mark.parametrize("path", _py_files(), ids=lambda p: os.path.relpath(p, ROOT))
def test_no_undefined_names(path):
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    if any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree)):
        pytest.skip("star import")
    v = _Scopes()
    v.visit(tree)
    assert not v.errors, ["%s:%d undefined name %r" % (os.path.relpath(path, ROOT), ln, nm) for ln, nm in v.errors]
