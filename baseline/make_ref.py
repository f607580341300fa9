"""Stage the UPSTREAM reference's hot-path files under baseline/_ref/ so that they travel to the GPU box.

    python baseline/make_ref.py            (also called by __graft_entry__.build() when /root/reference exists)

baseline/_ref/ is git-ignored (never part of this repo's history) but NOT gpurun-ignored, so `bench.py --impl reference`,
the `gpu_eager_reference` leg of the bench line and tests/test_gpu_dropin_literal.py can run the reference's OWN code
on the box: models/pointnet_extrusion.py + models/pointnet_util.py (backbone), losses.py, data_utils.py,
global_variables.py, utils.py, the training script whose loop body (train_Point2Cyl_without_sketch.py:244-353) is
exec'd, and IGR/{network,sampler,general}.py.  The files are copied byte for byte (UNMODIFIED; sha256 recorded in
MANIFEST.json); oracle/ref_shim.py supplies the import stubs (chamferdist, h5py, trimesh, ... and torch.symeig).
`pip install /root/reference` is not applicable: the reference has no setup.py / pyproject (flat scripts).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models/pointnet_extrusion.py", "models/pointnet_util.py", "losses.py", "data_utils.py", "global_variables.py",
         "utils.py", "train_Point2Cyl_without_sketch.py", "train_Point2Cyl.py", "IGR/network.py", "IGR/sampler.py",
         "IGR/general.py", "LICENSE"]


def make(src_root: str = "/root/reference") -> str:
    if not os.path.isfile(os.path.join(src_root, "models", "pointnet_util.py")):
        raise FileNotFoundError(f"no reference tree at {src_root}")
    manifest = {}
    for rel in FILES:
        src = os.path.join(src_root, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": src_root, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    return DST


def root():
    """Where the reference's files can be loaded from on this machine, or None."""
    for p in (os.environ.get("P2C_REFERENCE_ROOT"), DST, "/root/reference"):
        if p and os.path.isfile(os.path.join(p, "models", "pointnet_util.py")):
            return p
    return None


if __name__ == "__main__":
    print("staged", make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
