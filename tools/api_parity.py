"""API parity audit against the reference tree: for every reference module ``torchrec/<path>.py`` with a counterpart here
(``torchrec_b200/<path>.py``; ``distributed/`` maps to ``parallel/``) report

* public top-level names (classes, functions, CamelCase aliases) of the reference module that the counterpart neither defines nor imports,
* reference modules without a counterpart file (import aliases registered at runtime are not visible to this static check),
* constructor / function parameters of same-named public classes / functions that the counterpart does not accept
  (a ``*args`` / ``**kwargs`` counterpart is taken to accept everything).

Static (``ast`` only - nothing of either tree is imported). Usage::

    python tools/api_parity.py [--reference /root/reference/torchrec] [--names] [--signatures] [--modules]

It lists candidates, not verdicts: type variables, private helper classes of the reference's architecture (per-sharding-type
awaitables, FBGEMM wrappers) show up as "missing" although nothing user-visible depends on them."""
import argparse
import ast
import os
from typing import Dict, List, Optional, Set, Tuple


def _parse(path: str) -> Optional[ast.Module]:
    try:
        with open(path) as f:
            return ast.parse(f.read())
    except (OSError, SyntaxError):
        return None


def public_names(tree: ast.Module) -> Set[str]:
    out: Set[str] = set()
    for n in tree.body:
        if isinstance(n, (ast.ClassDef, ast.FunctionDef, ast.AsyncFunctionDef)) and not n.name.startswith("_"):
            out.add(n.name)
        elif isinstance(n, ast.Assign):
            for t in n.targets:
                if isinstance(t, ast.Name) and not t.id.startswith("_") and t.id[0].isupper() and not t.id.isupper():
                    out.add(t.id)
    return out


def available_names(tree: ast.Module) -> Tuple[Set[str], bool]:
    """Names importable from a module (defined, assigned or imported anywhere at module level incl. if / try blocks) and whether it star-imports."""
    out: Set[str] = set()
    star = False
    for n in ast.walk(tree):
        if isinstance(n, (ast.ClassDef, ast.FunctionDef, ast.AsyncFunctionDef)):
            out.add(n.name)
        elif isinstance(n, ast.Assign):
            for t in n.targets:
                if isinstance(t, ast.Name):
                    out.add(t.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            for a in n.names:
                if a.name == "*":
                    star = True
                else:
                    out.add((a.asname or a.name).split(".")[0])
    return out, star


def signatures(tree: ast.Module) -> Dict[str, Tuple[List[str], bool]]:
    out: Dict[str, Tuple[List[str], bool]] = {}
    for n in tree.body:
        if isinstance(n, ast.ClassDef):
            for m in n.body:
                if isinstance(m, ast.FunctionDef) and m.name == "__init__":
                    a = m.args
                    out[n.name] = ([x.arg for x in a.args[1:]] + [x.arg for x in a.kwonlyargs], a.vararg is not None or a.kwarg is not None)
        elif isinstance(n, ast.FunctionDef) and not n.name.startswith("_"):
            a = n.args
            out[n.name + "()"] = ([x.arg for x in a.args] + [x.arg for x in a.kwonlyargs], a.vararg is not None or a.kwarg is not None)
    return out


def counterpart(rel: str, mine: str) -> Optional[str]:
    for cand in (os.path.join(mine, rel), os.path.join(mine, rel.replace("distributed/", "parallel/", 1))):
        if os.path.exists(cand):
            return cand
    return None


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference/torchrec")
    ap.add_argument("--mine", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "torchrec_b200"))
    ap.add_argument("--names", action="store_true")
    ap.add_argument("--signatures", action="store_true")
    ap.add_argument("--modules", action="store_true")
    args = ap.parse_args()
    if not (args.names or args.signatures or args.modules):
        args.names = args.signatures = args.modules = True
    missing_modules: List[Tuple[str, int]] = []
    n_names = n_sigs = 0
    for dp, _, files in os.walk(args.reference):
        if "/tests" in dp or "fb" in dp.split(os.sep):
            continue
        for f in sorted(files):
            if not f.endswith(".py") or f.startswith("test_") or f == "__init__.py":
                continue
            rp = os.path.join(dp, f)
            rel = os.path.relpath(rp, args.reference)
            rtree = _parse(rp)
            if rtree is None or not public_names(rtree):
                continue
            mp = counterpart(rel, args.mine)
            if mp is None:
                missing_modules.append((rel, len(public_names(rtree))))
                continue
            mtree = _parse(mp)
            if mtree is None:
                continue
            if args.names:
                have, star = available_names(mtree)
                miss = sorted(public_names(rtree) - have)
                if miss:
                    n_names += 0 if star else len(miss)
                    print(f"[names]{' (star import: unverified)' if star else ''} {rel}: {miss}")
            if args.signatures:
                rs, ms = signatures(rtree), signatures(mtree)
                for name, (rargs, _) in rs.items():
                    if name in ms:
                        margs, mvar = ms[name]
                        miss_args = [a for a in rargs if a not in margs]
                        if miss_args and not mvar:
                            n_sigs += 1
                            print(f"[signature] {rel}: {name}: missing parameters {miss_args}")
    if args.modules:
        for rel, n in missing_modules:
            print(f"[module] {rel} ({n} public names) has no counterpart file")
    print(f"summary: {n_names} names, {n_sigs} signatures, {len(missing_modules)} modules flagged")


if __name__ == "__main__":
    main()
