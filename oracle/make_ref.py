"""TEST INFRASTRUCTURE - recipe that stages the UNMODIFIED reference under oracle/_ref/ so it can travel to the GPU box.

The reference (wubowen416/Speech-driven-Gesture-Generation-...) is pure Python with no setup.py / pyproject.toml, and
/root/reference does not exist on the GPU box.  This script copies, byte for byte, the files its sampler imports
(`models/`, `utils/json_config.py`, `utils/string_parser.py`, the two shipped configs) from where they lie into
`oracle/_ref/` (git-ignored, NOT gpurun-ignored: it ships with the snapshot like a built .so), adds the one stub the
import chain needs (`fasttext`, imported by models/modules/ha2g/model/vocab.py:5 and never used on the sampling path)
and writes a manifest of sha256 digests so that `oracle/ref_runner.py` can prove the staged files are unmodified.

    python oracle/make_ref.py            # build container only (needs /root/reference)

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may execute anything under oracle/; the product never does.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("GD_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
COPY = ["models", "utils/json_config.py", "utils/string_parser.py", "configs/beat-ours.json", "configs/tedexp-ours.json"]
STUB = '"""stub: models/modules/ha2g/model/vocab.py imports fasttext at module level; the sampling path never calls it."""\n'


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def build(force=False):
    """-> path of the staged reference, or None if the reference sources are not present (GPU box: use what shipped)."""
    manifest_path = os.path.join(DST, "MANIFEST.json")
    if not os.path.isdir(SRC):
        return DST if os.path.exists(manifest_path) else None
    if os.path.exists(manifest_path) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for rel in COPY:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
    for root, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, DST)] = _sha(p)
    with open(os.path.join(DST, "fasttext.py"), "w") as f:
        f.write(STUB)
    json.dump({"source": SRC, "files": manifest, "stubs": ["fasttext.py"]}, open(manifest_path, "w"), indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out if out else f"{SRC} not present and no staged copy: nothing to do")
