"""Stage the UNMODIFIED reference package for the reference arm of bench.py.

    python oracle/make_ref.py        # needs /root/reference; writes oracle/_ref/ (git-ignored, travels with gpurun)

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is copying its `sres`
package -- byte for byte, .py files only -- next to a manifest of SHA-256 digests.  `/root/reference` does not exist
on the GPU box; `oracle/_ref/` does (built artefacts are not gpurun-ignored), which lets `bench.py --impl reference`
and the `cpu_baseline` leg time the reference's OWN RCAN / downsample / l2loss / torch.optim.Adam code path on the
box's host cores instead of the oracle port.  Nothing under oracle/_ref/ is committed and nothing in the product
imports it (tests/test_abi_cpu.py::test_no_cpu_fallback).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/sres"
DST = os.path.join(HERE, "_ref")


def stage(src: str = SRC, dst: str = DST) -> str:
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} not found: the reference can only be staged in the build container")
    pkg = os.path.join(dst, "sres")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    manifest = {}
    for dirpath, dirnames, files in os.walk(src):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, src)
        out = os.path.join(pkg, rel) if rel != "." else pkg
        os.makedirs(out, exist_ok=True)
        for f in sorted(files):
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(dirpath, f), os.path.join(out, f))
                with open(os.path.join(dirpath, f), "rb") as fh:
                    manifest[os.path.normpath(os.path.join("sres", rel, f))] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump(dict(source=src, files=manifest), fh, indent=1, sort_keys=True)
    return dst


if __name__ == "__main__":
    print(stage(*sys.argv[1:3]))
