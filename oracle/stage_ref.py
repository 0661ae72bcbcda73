"""TEST INFRASTRUCTURE -- stages the UNMODIFIED reference so that its real CPU path can be timed on the GPU box.

The reference (SchlutowSM2Group/BLDFM) is pure Python; `/root/reference` exists only in the build container.
`stage()` copies its package directory `src/bldfm` byte for byte, together with the two import shims of
SURVEY.md Appendix B (`abltk` logger/paths stub, `pyfftw` -> scipy.fft; tests/golden/shims), into the
git-ignored `oracle/_ref/`, which travels to the GPU box like a built `.so`.  Nothing under `oracle/_ref/` is
part of the repository's history or of the product; `bench.py --impl reference` and the `cpu_baseline` leg
execute it through `oracle/reference_runner.py`, nothing else does.
"""
from __future__ import annotations

import shutil
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/src/bldfm")
SHIMS = HERE.parent / "tests" / "golden" / "shims"
DEST = HERE / "_ref"


def available() -> bool:
    return (DEST / "src" / "bldfm" / "solver.py").exists() and (DEST / "shims" / "abltk").exists()


def stage(force: bool = False) -> bool:
    """Copy the reference package + shims into oracle/_ref (no-op when the reference is not present here)."""
    if not REF_SRC.exists():
        return available()
    if available() and not force:
        return True
    if DEST.exists():
        shutil.rmtree(DEST)
    (DEST / "src").mkdir(parents=True)
    shutil.copytree(REF_SRC, DEST / "src" / "bldfm", ignore=shutil.ignore_patterns("__pycache__"))
    shutil.copytree(SHIMS, DEST / "shims", ignore=shutil.ignore_patterns("__pycache__", "*.pkl"))
    (DEST / "numba_cache").mkdir()
    return True


if __name__ == "__main__":
    print("staged" if stage(force=True) else "reference not available", DEST)
