"""Recipe for oracle/_ref/: the reference's OWN files for the -icl path, copied unmodified from /root/reference where they
lie (the reference is pure Python: there is nothing to compile).  oracle/_ref/ is git-ignored — nothing of it enters the
history — but it is not gpurun-ignored, so the real reference travels to the GPU box and bench.py can time a sample of
chromosome pairs through `wgdi.collinearity.collinearity(options, points).run()` (wgdi/collinearity.py:42-70) there.

    python oracle/make_ref.py            # run in the build container; a no-op where /root/reference does not exist

Test infrastructure / CPU baseline only: nothing under wgdi_b200/ imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
FILES = ("__init__.py", "collinearity.py", "run_colliearity.py", "base.py")     # the four files SURVEY.md 8(a) cites


def make_ref(verbose=True):
    src = os.path.join(REFERENCE, "wgdi")
    dst = os.path.join(HERE, "_ref", "wgdi")
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/make_ref.py: {src} not present; keeping whatever oracle/_ref holds")
        return os.path.isfile(os.path.join(dst, "collinearity.py"))
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    if verbose:
        print(f"oracle/_ref/wgdi: {', '.join(FILES)} copied unmodified from {src}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
