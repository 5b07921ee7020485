"""Place the UNMODIFIED reference where `bench.py --impl reference` can import it on the GPU box.

The reference (pure Python, absolute ``TeXOCR.*`` imports, no setup.py) is copied verbatim from ``/root/reference``
to ``baseline/_ref/TeXOCR/`` -- git-ignored (never part of this repo's history) but shipped to the GPU box with the
working tree.  Run in the build container only (``__graft_entry__.build()`` calls it when ``/root/reference`` exists);
on the GPU box the copy that travelled is used as it is.  Nothing under ``texocr_b200/`` imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref", "TeXOCR")


def install(force: bool = False) -> str:
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else ""
    if os.path.isdir(DST) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    return DST


def import_reference():
    """``import TeXOCR.model`` from baseline/_ref; raises ImportError when the copy is absent."""
    root = os.path.join(HERE, "_ref")
    if not os.path.isdir(os.path.join(root, "TeXOCR", "model")):
        raise ImportError(f"{root}/TeXOCR not present (run baseline/install_ref.py where /root/reference exists)")
    if root not in sys.path:
        sys.path.insert(0, root)
    import TeXOCR.model as M      # noqa
    return M


if __name__ == "__main__":
    print(install(force="--force" in sys.argv) or "reference not available here")
