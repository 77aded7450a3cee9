"""TEST INFRASTRUCTURE ONLY — compiles the reference's pure-Python modules that sit around the
rasterizer (gaussian_renderer/__init__.py, scene/{gaussian_model,cameras,deform_model}.py,
utils/*.py, arguments/__init__.py and the rasterizer package's own __init__.py) to bytecode, from
the sources where they lie under /root/reference, into oracle/_ref/pyref/ (git-ignored, travels to
the GPU box with the other checker builds).  No source is copied.  tests/test_dropin_render.py
imports the reference's render() from there when /root/reference itself is absent (the GPU box).

    python oracle/build_pyref.py [/root/reference]
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "pyref")
RAST_INIT = "submodules/diff-gaussian-rasterization-w-tof/diff_gaussian_rasterization_w_tof/__init__.py"


def main(ref="/root/reference"):
    jobs = [("gaussian_renderer/__init__.py", "gaussian_renderer/__init__.pyc"),
            ("arguments/__init__.py", "arguments/__init__.pyc"),
            (RAST_INIT, "ref_rasterizer_surface/__init__.pyc")]
    for d in ("scene", "utils"):
        for fn in sorted(os.listdir(os.path.join(ref, d))):
            if fn.endswith(".py") and fn != "__init__.py":
                jobs.append((f"{d}/{fn}", f"{d}/{fn}c"))
    n = 0
    for src, dst in jobs:
        s, t = os.path.join(ref, src), os.path.join(OUT, dst)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(t), exist_ok=True)
        py_compile.compile(s, cfile=t, dfile=src, doraise=True)
        n += 1
    # one archive next to the tree: file-sync tools commonly skip *.pyc, an opaque .tar travels
    import tarfile
    with tarfile.open(OUT + ".tar", "w") as tf:
        tf.add(OUT, arcname="pyref")
    print(f"compiled {n} reference modules to {OUT} (+ {OUT}.tar)")


if __name__ == "__main__":
    main(*sys.argv[1:2])
