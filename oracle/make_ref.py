"""Recipe for oracle/_ref: the UNMODIFIED reference, staged so that it can travel to the GPU box.

    python oracle/make_ref.py            (authoring container only: it reads /root/reference)

The reference (edwardyang12/Dex-NERF, nerf-pytorch/) is pure Python, so "building" it means copying its
package `nerf/` where it lies under /root/reference into oracle/_ref/nerf/ byte for byte (outputs only into
oracle/_ref/, which is git-ignored - the sources never enter the repository's history - but not
gpurun-ignored, so the copy ships to the GPU box like the built libdexnerf.so).  Next to it go the three
import shims of tests/golden/_shims (torchsearchsorted -> torch.searchsorted, imageio, matplotlib: modules the
reference imports that are absent from this image) and a MANIFEST with the sha256 of every copied file, so that
`bench.py --impl reference` can prove on the box that it is timing the reference's own bytes.

oracle/ref_runner.py (committed, ours) drives the staged reference: the two one-line repaired forwards of
SURVEY.md section 8a-3 are subclasses defined THERE, never edits of the copied files.

This is test / measurement infrastructure: only tests/, __graft_entry__.build() and bench.py's reference legs
touch it."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/nerf-pytorch"
DST = os.path.join(HERE, "_ref")
SHIMS = os.path.join(HERE, "..", "tests", "golden", "_shims")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build(force=False):
    """Stage the reference.  Returns the destination, or None when /root/reference is absent (GPU box: the
    prebuilt copy is used as shipped)."""
    if not os.path.isdir(os.path.join(SRC, "nerf")):
        return None
    manifest_path = os.path.join(DST, "MANIFEST.json")
    if os.path.exists(manifest_path) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.join(DST, "nerf"))
    manifest = {"source": SRC, "files": {}}
    for name in sorted(os.listdir(os.path.join(SRC, "nerf"))):
        if not name.endswith(".py"):
            continue
        src = os.path.join(SRC, "nerf", name)
        shutil.copyfile(src, os.path.join(DST, "nerf", name))
        manifest["files"]["nerf/" + name] = sha256(src)
    shutil.copytree(SHIMS, os.path.join(DST, "_shims"))
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return DST


def verify():
    """True when oracle/_ref holds exactly the bytes its MANIFEST records."""
    manifest_path = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(manifest_path):
        return False
    manifest = json.load(open(manifest_path))
    return all(os.path.exists(os.path.join(DST, rel)) and sha256(os.path.join(DST, rel)) == digest
               for rel, digest in manifest["files"].items())


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else "no /root/reference here: nothing staged")
