"""Import shim for the unmodified reference (test / benchmark tooling only; never imported by the product).

Looks for the reference at $STR2STR_REFERENCE, then /root/reference (build container), then the staged copy under
baseline/_ref/ (tools/stage_reference.py; what the GPU box has).

The reference (/root/reference) needs `dm-tree`, hydra and lightning at import time; none is installed.
Two shims are enough for the hot-path modules: a file-backed `tree.map_structure`
and an empty `src.utils` package so `src.utils.tensor_utils` loads without `src/utils/__init__.py`.
"""
import importlib.machinery
import os
import sys
import tempfile
import types

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def find_reference():
    for cand in (os.environ.get("STR2STR_REFERENCE"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "src", "models", "net")):
            return cand
    return None


REF = find_reference()


def available() -> bool:
    return REF is not None


def install():
    if REF is None:
        raise RuntimeError("reference sources not found (neither /root/reference nor baseline/_ref: run tools/stage_reference.py in the build container)")
    shim_dir = os.path.join(tempfile.gettempdir(), "str2str_refshim")
    os.makedirs(shim_dir, exist_ok=True)
    with open(os.path.join(shim_dir, "tree.py"), "w") as f:
        f.write(
            "def map_structure(fn, *s):\n"
            "    a = s[0]\n"
            "    if isinstance(a, dict):\n"
            "        return {k: map_structure(fn, *[x[k] for x in s]) for k in a}\n"
            "    if isinstance(a, (list, tuple)):\n"
            "        return type(a)(map_structure(fn, *xs) for xs in zip(*s))\n"
            "    return fn(*s)\n"
        )
    for p in (shim_dir, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import src  # noqa: F401  (reference's empty top-level package)

    if "src.utils" not in sys.modules:
        pkg = types.ModuleType("src.utils")
        pkg.__path__ = [os.path.join(REF, "src", "utils")]
        pkg.__spec__ = importlib.machinery.ModuleSpec("src.utils", None, is_package=True)
        sys.modules["src.utils"] = pkg
