"""AST docstring linter for nn.Module classes (reference torchrec/linter/module_linter.py): every public module class
needs a docstring; when ``__init__`` takes arguments the docstring should have an ``Args`` section; ``forward`` with a
non-trivial signature should be documented (``Returns`` / ``Example``). Usage: ``python -m torchrec_b200.linter.module_linter file.py``."""
import ast
import json
import sys
from typing import Any, Dict, List

MAX_NUM_ARGS_IN_MODULE_CTOR = 5


def print_error_message(python_path: str, node: ast.AST, name: str, message: str, severity: str = "warning") -> Dict[str, Any]:
    lint_item = {"path": python_path, "line": getattr(node, "lineno", 0), "char": getattr(node, "col_offset", 0) + 1, "severity": severity, "name": name, "description": message}
    print(json.dumps(lint_item))
    return lint_item


def get_function_args(node: ast.FunctionDef) -> List[str]:
    a = node.args
    names = [x.arg for x in a.posonlyargs + a.args + a.kwonlyargs]
    if a.vararg:
        names.append(a.vararg.arg)
    if a.kwarg:
        names.append(a.kwarg.arg)
    return [n for n in names if n not in ("self", "cls")]


def check_class_definition(python_path: str, node: ast.ClassDef) -> List[Dict[str, Any]]:
    issues: List[Dict[str, Any]] = []
    is_module = any((isinstance(b, ast.Attribute) and b.attr == "Module") or (isinstance(b, ast.Name) and b.id == "Module") for b in node.bases)
    if not is_module or node.name.startswith("_"):
        return issues
    doc = ast.get_docstring(node)
    if doc is None:
        issues.append(print_error_message(python_path, node, "docstring-missing", f"Module `{node.name}` has no docstring"))
        return issues
    funcs = {f.name: f for f in node.body if isinstance(f, ast.FunctionDef)}
    init = funcs.get("__init__")
    if init is not None:
        args = get_function_args(init)
        if len(args) > MAX_NUM_ARGS_IN_MODULE_CTOR:
            issues.append(print_error_message(python_path, init, "too-many-ctor-args", f"Module `{node.name}` constructor takes {len(args)} arguments (> {MAX_NUM_ARGS_IN_MODULE_CTOR}); consider a config object"))
        if args and "Args:" not in doc and "Args\n" not in doc:
            issues.append(print_error_message(python_path, init, "args-section-missing", f"Docstring of `{node.name}` has no `Args:` section although __init__ takes {args}"))
        for a in args:
            if ("Args:" in doc) and a not in doc and not a.startswith("_"):
                issues.append(print_error_message(python_path, init, "arg-undocumented", f"Argument `{a}` of `{node.name}.__init__` is not mentioned in the docstring"))
    fwd = funcs.get("forward")
    if fwd is not None and get_function_args(fwd) and "Example" not in doc and "Returns" not in doc and "->" not in doc:
        issues.append(print_error_message(python_path, fwd, "forward-undocumented", f"Docstring of `{node.name}` documents neither the forward contract (`Returns:`) nor an `Example::`"))
    return issues


def linter_one_file(python_path: str) -> List[Dict[str, Any]]:
    with open(python_path, "r") as f:
        src = f.read()
    try:
        tree = ast.parse(src)
    except SyntaxError as e:
        return [print_error_message(python_path, ast.Module(body=[], type_ignores=[]), "syntax-error", str(e), "error")]
    issues: List[Dict[str, Any]] = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef):
            issues.extend(check_class_definition(python_path, node))
    return issues


if __name__ == "__main__":
    for p in sys.argv[1:]:
        linter_one_file(p)
