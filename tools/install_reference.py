"""'Installs' the unmodified reference for bench.py's reference arm: the reference is a script tree without packaging
metadata (pip has nothing to install), so the hot path's package -- models/ and configs/mine.yml -- is copied as is
into baseline/_ref/ (git-ignored, travels to the GPU box with the snapshot; never part of the repo's history).
Run by __graft_entry__.build() whenever /root/reference (or $EVC_REF) is present."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install(src=None, quiet=False):
    src = src or os.environ.get("EVC_REF", "/root/reference")
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(src, "models")):
        if not quiet:
            print(f"[install_reference] {src}/models not found: nothing installed")
        return False
    os.makedirs(dst, exist_ok=True)
    ign = shutil.ignore_patterns("__pycache__", "*.pyc", "weights", "fvd", "*.pth", "*.pt")
    for sub in ("models", "configs"):
        d = os.path.join(dst, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(src, sub), d, ignore=ign)
    with open(os.path.join(dst, "INSTALLED_FROM"), "w") as f:
        f.write(f"{src}\nunmodified copy of models/ (minus weights/, fvd/) and configs/ for bench.py --impl reference\n")
    if not quiet:
        print(f"[install_reference] {src} -> {dst}")
    return True


if __name__ == "__main__":
    install(sys.argv[1] if len(sys.argv) > 1 else None)
