"""Stage the UNMODIFIED reference sources of the hot path under the git-ignored baseline/_ref/ so that they travel to the
GPU box with the repo snapshot (gpurun ships git-ignored files; /root/reference itself does not exist there).

    python tools/stage_reference.py        (build container only; __graft_entry__.build() runs it when /root/reference exists)

Nothing staged here is product source or part of the repository history: bench.py --impl reference and the `ref_gpu`
record import it (through tools/refshim.py) to time the reference itself, and for nothing else.  Files are copied byte for
byte; a manifest with their sha256 is written beside them so that a reader can check they are unmodified.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SRC = os.environ.get("STR2STR_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
# the python packages the hot path imports + the data file residue_constants reads at import
KEEP = ("src/__init__.py", "src/common", "src/models", "src/utils/tensor_utils.py", "LICENSE")


def stage() -> str:
    if not os.path.isdir(SRC):
        raise RuntimeError(f"{SRC} not found: the reference can only be staged in the build container")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for item in KEEP:
        s = os.path.join(SRC, item)
        if not os.path.exists(s):
            continue
        files = [s] if os.path.isfile(s) else [os.path.join(d, f) for d, _, fs in os.walk(s) for f in fs if not f.endswith(".pyc")]
        for f in files:
            relp = os.path.relpath(f, SRC)
            out = os.path.join(DST, relp)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(f, out)
            manifest[relp] = hashlib.sha256(open(f, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    print(stage(), file=sys.stderr)
