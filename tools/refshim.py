"""Import shim for the read-only reference checkout (test tooling only; never imported by the product).

The reference (/root/reference) needs `dm-tree`, hydra and lightning at import time; none is installed.
Two shims are enough for the hot-path modules: a file-backed `tree.map_structure`
and an empty `src.utils` package so `src.utils.tensor_utils` loads without `src/utils/__init__.py`.
"""
import importlib.machinery
import os
import sys
import tempfile
import types

REF = os.environ.get("STR2STR_REFERENCE", "/root/reference")


def install():
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference checkout not found at {REF}")
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
