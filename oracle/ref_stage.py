"""Stages the reference's own modules for the score path into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference (MrGreyPanda/NEM-MCMC-optimization) is pure Python: there is nothing to compile, but its
functions can be timed and used as the checker on the GPU box only if its files travel with the snapshot.
`stage()` copies nem.py and utils.py byte for byte from the read-only reference mount into oracle/_ref/
(git-ignored like a built .so, NOT gpurun-ignored, never committed: the repository history holds no
reference source).  `__graft_entry__.build()` calls it whenever /root/reference is present; on the GPU box
the staged files are simply used as they arrived.  Only nem.py (score tables + score, :25-86) and utils.py
(ancestor :37-54, data synthesis :5-35) are on the path; nem_order_mcmc.py (wandb import, broken driver) is not.
"""
from __future__ import annotations

import os
import shutil

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REF_SRC = os.environ.get("NEM_REFERENCE_ROOT", "/root/reference")
FILES = ("nem.py", "utils.py")


def staged() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def stage(verbose: bool = False) -> bool:
    """-> True if oracle/_ref holds the reference modules afterwards."""
    if all(os.path.exists(os.path.join(REF_SRC, f)) for f in FILES):
        os.makedirs(REF_DIR, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(REF_DIR, f))
        with open(os.path.join(REF_DIR, "STAGED_FROM"), "w") as fh:
            fh.write(f"{REF_SRC}: {', '.join(FILES)} (unmodified copies, staged by oracle/ref_stage.py; not part of the repository)\n")
        if verbose:
            print(f"staged {FILES} from {REF_SRC} into {REF_DIR}")
    return staged()


if __name__ == "__main__":
    print("staged" if stage(verbose=True) else f"{REF_SRC} not mounted and {REF_DIR} empty: reference arm falls back to the port")
