"""Recipe for ``oracle/_ref/``: packs the reference's OWN, unmodified Python package (``/root/reference/mbrl/**/*.py``, plus
the settings yamls under ``experiments/``) into
``oracle/_ref/reference_mbrl.zip`` so that ``bench.py --impl reference`` can time the reference itself on the GPU box, where
``/root/reference`` does not exist (the archive is git-ignored -- no reference source enters the history -- but not
gpurun-ignored, so it travels with the snapshot like the built ``.so``).  Imported through ``zipimport`` behind the stub
modules of ``oracle/ref_shims.py``.  Run by ``__graft_entry__.build()`` wherever the reference tree is mounted.

BASELINE INFRASTRUCTURE: only ``bench.py --impl reference`` / its ``cpu_baseline`` leg and the tests read the archive."""
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "reference_mbrl.zip")


def build(reference_root="/root/reference"):
    src = os.path.join(reference_root, "mbrl")
    if not os.path.isdir(src):
        return None  # GPU box: the prebuilt archive (if any) is used as it is
    files = []
    for root, _, names in os.walk(src):
        files += [os.path.join(root, n) for n in names if n.endswith(".py")]
    # the settings files the boundary tests build the drop-ins from (yaml field names are the reference's API)
    for root, _, names in os.walk(os.path.join(reference_root, "experiments")):
        files += [os.path.join(root, n) for n in names if n.endswith(".yaml")]
    files.sort()
    newest = max(os.path.getmtime(f) for f in files)
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with zipfile.ZipFile(OUT + ".tmp", "w", zipfile.ZIP_DEFLATED) as z:
        for f in files:
            z.write(f, os.path.relpath(f, reference_root))
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(*sys.argv[1:]))
