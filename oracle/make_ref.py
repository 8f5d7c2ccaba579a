"""Recipe for oracle/_ref: the reference's OWN pure-Python modules of the hot path, copied verbatim from the read-only
tree where it lies (/root/reference) so that they travel to the GPU box with the snapshot.

TEST INFRASTRUCTURE ONLY (checker / CPU baseline, never the product). oracle/_ref/ is git-ignored -- reference sources are
never committed -- but not gpurun-ignored. `__graft_entry__.build()` runs this when the reference tree is present; on the
GPU box the already-copied files are used as they are. Nothing here is edited: files are byte-for-byte copies
(checksummed in MANIFEST.json); the only files written by us are the two empty package markers the reference's imports
need (`data/__init__.py` replaces the reference's data/__init__.py, whose dataset imports need pc_util / plyfile / h5py).

    python oracle/make_ref.py [--src /root/reference]
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# reference files on the path SURVEY.md section 8(a) names (+ DeIT.py, star-imported by both model files)
FILES = ["models/__init__.py", "models/vit_3d_2d_pretrain.py", "models/embed_layer_3d_modality.py", "models/DeIT.py",
         "models/3DViT/model.py", "data/pointnet_util.py", "utils/binvox_rw.py"]
OURS = {"data/__init__.py": "", "utils/__init__.py": ""}  # empty package markers (not reference code)


def make(src="/root/reference"):
    if not os.path.isfile(os.path.join(src, "models", "vit_3d_2d_pretrain.py")):
        return False
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    for rel, text in OURS.items():
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        with open(d, "w") as f:
            f.write(text)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "/root/reference"
    print("oracle/_ref written" if make(src) else "reference tree not found: nothing written")
