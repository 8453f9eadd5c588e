"""Recipe for ``oracle/_ref`` (TEST INFRASTRUCTURE ONLY): the reference's own two source files of the hot path, copied
verbatim from the read-only reference tree into ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored, so it travels to the
GPU box where /root/reference does not exist).  Nothing is copied into tracked paths.

    python -m oracle.build_ref            # run in the authoring container; __graft_entry__.build() calls it too

``bench.py --impl reference`` then executes the reference's classes themselves (``GAN_final.py``'s
``CasNetGenerator`` / ``Discriminator`` / ``GAN.training_step`` through ``oracle/ref_shim.py``: stubs for the absent
monai / pytorch_lightning / itk imports, the restated MONAI 0.4.0 UNet for the un-vendored third-party class) and
reports ``cpu_baseline.kind = "reference"``; without the copy it falls back to the oracle restatement (``"port"``).
"""
import os
import shutil
import sys

from .ref_shim import REF_COPY_ROOT, REFERENCE_ROOT

FILES = ("code/GAN/GAN_final.py", "test_runs/GAN.py")


def build(verbose=True):
    if not os.path.isdir(REFERENCE_ROOT):
        if verbose:
            print(f"oracle/_ref: {REFERENCE_ROOT} absent (GPU box?) -- keeping whatever copy travelled with the snapshot")
        return False
    for rel in FILES:
        src, dst = os.path.join(REFERENCE_ROOT, rel), os.path.join(REF_COPY_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    if verbose:
        print(f"oracle/_ref: copied {', '.join(FILES)}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
