#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules of the hot path into oracle/_ref/ (git-ignored, shipped to the GPU box with
the snapshot like the built .so) so that `bench.py --impl reference` can time the real reference on the box's host
cores (cpu_baseline.kind = "reference").

    python oracle/make_ref.py            # needs /root/reference (or $PM_REFERENCE); no-op message otherwise

The reference is pure Python: "building" it means byte-compiling models/{emage,camn,disco}_audio/*.py where they
lie under /root/reference into sourceless .pyc files (py_compile; the GPU box runs the same image, hence the same
interpreter) - the analogue of compiling a C reference into oracle/_ref/*.so.  No reference SOURCE is copied
anywhere.  oracle/_ref/MANIFEST.json records the SHA-256 of every source that was compiled and of every .pyc; the
loader (oracle/ref_loader.py) refuses a tree whose hashes do not match.  Nothing staged here is tracked by git and
nothing of it is imported by the product (tests/test_boundary.py).  TEST / MEASUREMENT INFRASTRUCTURE ONLY.
"""
import hashlib
import json
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ("emage_audio", "camn_audio", "disco_audio")


def stage(ref_root=None, quiet=False):
    ref_root = ref_root or os.environ.get("PM_REFERENCE", "/root/reference")
    src_models = os.path.join(ref_root, "models")
    if not os.path.isdir(src_models):
        if not quiet:
            print(f"{ref_root} not present: oracle/_ref left as it is", file=sys.stderr)
        return None
    manifest = {}
    shutil.rmtree(os.path.join(DEST, "models"), ignore_errors=True)
    for pkg in PACKAGES:
        src, dst = os.path.join(src_models, pkg), os.path.join(DEST, "models", pkg)
        os.makedirs(dst, exist_ok=True)
        for name in sorted(os.listdir(src)):
            if not name.endswith(".py"):
                continue
            out = os.path.join(dst, name + "c")                      # sourceless import layout: module.pyc next to nothing
            py_compile.compile(os.path.join(src, name), cfile=out, dfile=f"models/{pkg}/{name}", doraise=True)
            with open(os.path.join(src, name), "rb") as f:
                src_hash = hashlib.sha256(f.read()).hexdigest()
            with open(out, "rb") as f:
                manifest[f"models/{pkg}/{name}c"] = {"sha256": hashlib.sha256(f.read()).hexdigest(), "source_sha256": src_hash}
    # the reference's `models` is a namespace package; this repo has a regular `models/` shim package, which would win
    # the import no matter the path order - an empty marker of our own makes the staged tree a regular package too
    open(os.path.join(DEST, "models", "__init__.py"), "w").close()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref_root, "python": sys.version.split()[0], "files": manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print(f"staged {len(manifest)} reference files into {DEST}")
    return DEST


if __name__ == "__main__":
    stage()
