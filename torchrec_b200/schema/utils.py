"""Backward-compatibility check of public signatures (reference torchrec/schema/utils.py:38)."""
import inspect
from typing import Any


def is_signature_compatible(previous_signature: inspect.Signature, current_signature: inspect.Signature) -> bool:
    """True when every call valid for ``previous_signature`` is valid for ``current_signature``:
    positional parameters keep their order, names and defaults; new parameters have defaults; *args/**kwargs are kept;
    the return annotation is unchanged."""
    prev, cur = list(previous_signature.parameters.values()), list(current_signature.parameters.values())
    P = inspect.Parameter
    expected_kinds = (P.POSITIONAL_ONLY, P.POSITIONAL_OR_KEYWORD)
    cur_by_name = {p.name: p for p in cur}
    cur_has_var_pos = any(p.kind == P.VAR_POSITIONAL for p in cur)
    cur_has_var_kw = any(p.kind == P.VAR_KEYWORD for p in cur)
    i = 0
    for pp in prev:
        if pp.kind in expected_kinds:
            if i >= len(cur) or cur[i].kind not in expected_kinds:
                return False
            cp = cur[i]
            if cp.name != pp.name and pp.kind == P.POSITIONAL_OR_KEYWORD:
                return False
            if pp.default is not P.empty and cp.default is P.empty:
                return False
            if not _annot_ok(pp.annotation, cp.annotation):
                return False
            i += 1
        elif pp.kind == P.VAR_POSITIONAL:
            if not cur_has_var_pos:
                return False
        elif pp.kind == P.KEYWORD_ONLY:
            cp = cur_by_name.get(pp.name)
            if cp is None:
                if not cur_has_var_kw:
                    return False
            else:
                if pp.default is not P.empty and cp.default is P.empty:
                    return False
                if not _annot_ok(pp.annotation, cp.annotation):
                    return False
        elif pp.kind == P.VAR_KEYWORD:
            if not cur_has_var_kw:
                return False
    prev_names = {p.name for p in prev}
    for cp in cur:
        if cp.name not in prev_names and cp.kind in (P.POSITIONAL_ONLY, P.POSITIONAL_OR_KEYWORD, P.KEYWORD_ONLY) and cp.default is P.empty:
            return False
    return _annot_ok(previous_signature.return_annotation, current_signature.return_annotation)


def _annot_ok(prev: Any, cur: Any) -> bool:
    if prev is inspect.Signature.empty or prev is inspect.Parameter.empty:
        return True
    return str(prev) == str(cur) or prev == cur
