"""Materialise ``oracle/_ref/`` -- the reference's OWN CPU implementation of the deformation half of the path -- so that
it travels to the GPU box (where /root/reference does not exist).

TEST / BASELINE INFRASTRUCTURE ONLY.  Called by ``__graft_entry__.build()`` in the build container.  The four files the
reference's ``scene.deformation.deform_network`` needs are copied UNMODIFIED from where they lie under /root/reference
into ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored: like a built .so it ships with the snapshot but never enters the
history).  ``oracle/ref_loader.py`` imports the module from /root/reference when present, else from here; it is used by
``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference(deform)+port(rasterizer)") and ``gpu_eager_baseline``.

The rasterizer half cannot be materialised: submodules/depth-diff-gaussian-rasterization is an empty directory in the
reference tree, so there is nothing to copy or compile -- ``oracle/raster_ref.c`` stays a "port".
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("G4D_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("scene/deformation.py", "scene/hexplane.py", "scene/grid.py", "utils/graphics_utils.py",
         "utils/loss_utils.py", "scene/regulation.py")


def materialise(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref is complete (freshly copied or already there)."""
    if not os.path.isfile(os.path.join(REF_SRC, FILES[0])):
        return all(os.path.isfile(os.path.join(REF_DST, f)) for f in FILES)
    for f in FILES:
        src, dst = os.path.join(REF_SRC, f), os.path.join(REF_DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            shutil.copyfile(src, dst)
            if verbose:
                print("oracle/_ref <-", src)
    with open(os.path.join(REF_DST, "README"), "w") as fh:
        fh.write("Unmodified copies of %s from %s, made by oracle/make_ref.py at build() time.\n"
                 "git-ignored; baseline infrastructure only (see oracle/make_ref.py).\n" % (", ".join(FILES), REF_SRC))
    return True


if __name__ == "__main__":
    print("complete" if materialise(verbose=True) else "reference tree absent and oracle/_ref incomplete")
