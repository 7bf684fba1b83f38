"""Recipe for oracle/_ref: a runnable copy of the UNMODIFIED reference modules of the hot path.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).  The reference is a Python package that lives at
/root/reference in the build container only; the GPU box has no /root/reference.  This script imports the reference
through oracle/ref_harness.py (two in-memory import stubs, no reference file is modified), records exactly which
`fast3r.*` source files the hot path (`Fast3R`, `inference`) pulls in, and copies those files verbatim to
`oracle/_ref/fast3r/...` together with a manifest of sha256 digests.  `oracle/_ref/` is git-ignored (reference
sources never enter the history) but travels with the gpurun snapshot, so that on the GPU box
`bench.py --impl reference` times the reference's OWN `inference()` on the host cores (kind "reference") and the
reference model can be run on the B200 as the library bar.

Run in the build container:  python oracle/make_ref.py      (also called by __graft_entry__.build())
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def build(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(SRC, "fast3r")):
        if os.path.isdir(os.path.join(DST, "fast3r")):
            return DST  # GPU box: use the prebuilt copy
        raise RuntimeError(f"{SRC} not present and no prebuilt oracle/_ref")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["FAST3R_REFERENCE_ROOT"] = SRC
    from oracle.ref_harness import import_reference
    import_reference()
    files = sorted({m.__file__ for n, m in list(sys.modules.items())
                    if n.split(".")[0] == "fast3r" and getattr(m, "__file__", None)
                    and os.path.abspath(m.__file__).startswith(SRC + os.sep)})
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for f in files:
        rel = os.path.relpath(f, SRC)
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(f, out)
        manifest[rel] = hashlib.sha256(open(f, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(files)} reference files copied ({sum(os.path.getsize(f) for f in files)} bytes)")
    return DST


if __name__ == "__main__":
    build()
