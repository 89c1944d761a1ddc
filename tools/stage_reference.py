"""Stage the UNMODIFIED reference files the GPU box needs under baseline/_ref/ (git-ignored, but it travels with the gpurun snapshot).

    python tools/stage_reference.py            # no-op when /root/reference is absent

/root/reference does not exist on the GPU box.  Two things there want the reference's own code, byte for byte:
  * tests/test_reference_scripts.py runs the reference's inference demo `test_experiment.py` UNCHANGED against the drop-in module
    (north_star: "so test.py and experiments/*_experiment.py run unchanged");
  * `bench.py --impl reference` / `gpu_baseline` time the reference's own `models/hit_sir_pro.py` (kind: "reference").
Nothing under baseline/_ref is product source, nothing there is ever committed (.gitignore), and the product never imports it.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/hit_sir_pro.py", "utils/utils.py", "utils/arch_util.py", "test_experiment.py"]


def stage(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(REF, "models")):
        return False
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    for pkg in ("models", "utils"):                     # the reference tree relies on namespace packages; keep it that way
        init = os.path.join(REF, pkg, "__init__.py")
        if os.path.exists(init):
            shutil.copyfile(init, os.path.join(DST, pkg, "__init__.py"))
    if verbose:
        print(f"staged {len(FILES)} reference files under {DST}")
    return True


def reference_root():
    """Directory holding the reference's models/ and utils/ (the live tree here, the staged copy on the GPU box), or None."""
    for root in (REF, DST):
        if os.path.exists(os.path.join(root, "models", "hit_sir_pro.py")):
            return root
    return None


if __name__ == "__main__":
    sys.exit(0 if stage() or True else 1)
