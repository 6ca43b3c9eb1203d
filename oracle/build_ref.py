"""Recipe for oracle/_ref/: the UNMODIFIED reference modules that bench.py's CPU arm times
(`cpu_baseline.kind = "reference"`, `--impl reference`).

The reference is pure Python (RVMC.py, RVMC_bias_corr.py): nothing is compiled.  The two files are
copied byte for byte from /root/reference into the git-ignored oracle/_ref/ (never into history, never
into the product package); it travels to the GPU box with the repo snapshot, where /root/reference does
not exist.  They import under the two shim packages of oracle/shims/ (astropy.constants.G pinned to
6.6743e-11, an empty matplotlib.pyplot) -- SURVEY Appendix B.  Run by __graft_entry__.build(); a no-op
when /root/reference is absent (then an existing oracle/_ref/ is used as is, and without one the CPU arm
falls back to the oracle port, kind "port").

    python oracle/build_ref.py
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
FILES = ("RVMC.py", "RVMC_bias_corr.py")


def build_ref():
    """-> path of oracle/_ref if the reference modules are there (copied now or earlier), else None."""
    if os.path.isdir(REF):
        os.makedirs(OUT, exist_ok=True)
        for name in FILES:
            src, dst = os.path.join(REF, name), os.path.join(OUT, name)
            if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
                shutil.copyfile(src, dst)
    return OUT if all(os.path.exists(os.path.join(OUT, n)) for n in FILES) else None


def load_reference():
    """(RVMC module, RVMC_bias_corr module) of the unmodified reference, or None when oracle/_ref is absent."""
    import sys
    import warnings
    if not all(os.path.exists(os.path.join(OUT, n)) for n in FILES):
        return None
    for p in (OUT, os.path.join(HERE, "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")           # "\o" SyntaxWarnings of RVMC.py:291,318 on Python 3.12
        import RVMC
        import RVMC_bias_corr
    return RVMC, RVMC_bias_corr


if __name__ == "__main__":
    print(build_ref())
