"""Vendors the reference's own Python sources of the hot path into oracle/_ref/ (git-ignored, NOT gpurun-ignored), so that
the CPU arm of bench.py can execute the REFERENCE's code on the GPU box, where /root/reference does not exist.

    python oracle/build_ref.py          (also called by __graft_entry__.build() when /root/reference is present)

Nothing is copied into the repository's history: oracle/_ref/ is listed in .gitignore, and nothing in the product
(dfa-nerf_b200/) ever imports it -- only bench.py's CPU legs and the tests do, through oracle/ref_arm.py.
"""
import os
import shutil

REF = '/root/reference/NeRFs/DFANeRF'
FILES = ('run_nerf_helpers.py', 'decoder.py', 'run_nerf_com_trainExpLater.py', 'load_audface.py')
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def build_ref():
    """Returns True when oracle/_ref/ holds the reference sources afterwards."""
    if os.path.isdir(REF):
        os.makedirs(DST, exist_ok=True)
        for f in FILES:
            shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


if __name__ == '__main__':
    print('oracle/_ref:', 'ready' if build_ref() else 'unavailable (no /root/reference here)')
