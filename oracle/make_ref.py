"""Recipe for oracle/_ref/: vendors the UNMODIFIED reference modules of the hot path (byte-for-byte
copies of /root/reference/barf/*.py) so that the GPU box — which has no /root/reference — can time the
reference's own CPU implementation beside the CUDA path (`bench.py --impl reference`, `cpu_baseline`).

TEST / BASELINE INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (no reference source enters the
history) but not gpurun-ignored (it travels with the snapshot).  `__graft_entry__.build()` runs this
when /root/reference is present.  Nothing in the product package imports it.

    python oracle/make_ref.py            # -> oracle/_ref/barf/*.py + MANIFEST.json (sha256 per file)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("NERF_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")

# the modules BarfModel.training_step touches (barf/run_barf.py:151-196 builds exactly these)
FILES = {
    "barf": ["model_barf.py", "model_camera_calibration.py", "model_camera_extrinsics.py", "model_interpolation.py",
             "model_interpolation_architecture.py", "positional_encodings.py", "magic.py", "data_module.py",
             "dataset.py"],
}


def main() -> int:
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "barf")):
        print(f"make_ref: no reference under {REFERENCE_ROOT}; keeping oracle/_ref as it is")
        return 0
    manifest = {}
    for variant, names in FILES.items():
        dst_dir = os.path.join(DEST, variant)
        os.makedirs(dst_dir, exist_ok=True)
        for n in names:
            src = os.path.join(REFERENCE_ROOT, variant, n)
            dst = os.path.join(dst_dir, n)
            shutil.copyfile(src, dst)
            manifest[f"{variant}/{n}"] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REFERENCE_ROOT, "sha256": manifest}, f, indent=1, sort_keys=True)
    print(f"make_ref: {len(manifest)} unmodified reference files -> {DEST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
