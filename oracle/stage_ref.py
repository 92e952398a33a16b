"""Stage the UNMODIFIED reference into oracle/_ref/ (git-ignored, travels to the GPU box).

TEST / BENCH INFRASTRUCTURE ONLY -- nothing under oracle/ is imported by the product package.

The reference (`/root/reference`) is a pure-Python project whose `lib/` does not import as shipped
(SURVEY.md R8): `lib/convsubsampling.py:3` imports `lib.conformer`, a file that only exists under
`unused_lib/`.  Staging = copy `lib/` and add the two files its stray import needs, exactly as
SURVEY.md Appendix D describes.  No reference source is copied into the git history.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B2C_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(force: bool = False) -> bool:
    """Returns True when oracle/_ref/lib is importable afterwards."""
    if os.path.isdir(os.path.join(DST, "lib")) and not force:
        return True
    if not os.path.isdir(os.path.join(REF, "lib")):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(REF, "lib"), os.path.join(DST, "lib"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("conformer.py", "relativepositionalembeddings.py"):      # R8 shim
        shutil.copy(os.path.join(REF, "unused_lib", f), os.path.join(DST, "lib", f))
    for d in ("", "standard", "finetuning"):
        init = os.path.join(DST, "lib", d, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    return True


if __name__ == "__main__":
    ok = stage(force="--force" in sys.argv)
    print("oracle/_ref staged" if ok else "reference not present; oracle/_ref not staged")
