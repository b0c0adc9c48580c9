"""TEST INFRASTRUCTURE — builds the reference's own Python package twice from /root/reference (read-only;
a copy under /tmp is compiled, nothing is copied into the repo's history):

  baseline/_ref/       : the UNMODIFIED reference, `pip install --target` (stock C sources, OpenMP)
  baseline/_ref_b200/  : the same Cython wrapper modules (poismf/cfuns_{double,float}.pyx + poismf_c_wrapper.pxi,
                         untouched) compiled against poismf_b200/host/poismf_host.c INSTEAD of src/poismf.c,
                         src/pred.c, src/topN.c, src/nonnegcg.c, src/tnc.c and linked to libpoismf_b200.so —
                         exactly the change INTEGRATION.md §2 describes; the package's __init__.py is the
                         reference's own file, taken from the stock install.

Both directories are git-ignored and travel to the GPU box with the snapshot.  tests/test_boundary.py imports
the reference's PoisMF class from each and compares fits, predictions, top-N lists and transforms.

    python scripts/build_wrapper.py [--force]
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PMF_REFERENCE", "/root/reference")
STOCK = os.path.join(ROOT, "baseline", "_ref")
OURS = os.path.join(ROOT, "baseline", "_ref_b200")
TMP = "/tmp/poismf_ref_build"

SETUP_B200 = r'''
import numpy, os
from setuptools import setup, Extension
from Cython.Build import cythonize
B200 = {root!r}
def ext(name, pyx, macros):
    return Extension(name, [pyx, os.path.join(B200, "poismf_b200", "host", "poismf_host.c")],
                     include_dirs=[numpy.get_include(), "src", os.path.join(B200, "include")],
                     define_macros=[("_FOR_PYTHON", None), ("NDEBUG", None)] + macros,
                     extra_compile_args=["-O2", "-std=c99"],
                     libraries=["poismf_b200"], library_dirs=[os.path.join(B200, "poismf_b200")],
                     extra_link_args=["-Wl,-rpath,$ORIGIN/../../../poismf_b200"])
setup(name="poismf_b200_wrapper", packages=[],
      ext_modules=cythonize([ext("poismf.c_funs_double", "poismf/cfuns_double.pyx", []),
                             ext("poismf.c_funs_float", "poismf/cfuns_float.pyx", [("USE_FLOAT", None)])],
                            language_level=2))
'''


def run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("failed: " + " ".join(cmd) + "\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return r


def main(force=False):
    if not os.path.isdir(REF):
        print("reference sources not present: nothing to build (prebuilt directories are used as they are)")
        return
    env = dict(os.environ, CC="/usr/bin/gcc", LDSHARED="/usr/bin/gcc -shared", DONT_SET_MARCH="1",
               CFLAGS="-march=x86-64-v3")
    if force or not os.path.isdir(os.path.join(STOCK, "poismf")):
        shutil.rmtree(TMP, ignore_errors=True)
        shutil.copytree(REF, TMP)
        shutil.rmtree(STOCK, ignore_errors=True)
        run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
             "--find-links", "/opt/wheelhouse", "--target", STOCK, TMP], env=env)
    if force or not os.path.isdir(os.path.join(OURS, "poismf")):
        from poismf_b200.build import build
        build()
        if not os.path.isdir(TMP):
            shutil.copytree(REF, TMP)
        with open(os.path.join(TMP, "setup_b200.py"), "w") as f:
            f.write(SETUP_B200.format(root=ROOT))
        shutil.rmtree(OURS, ignore_errors=True)
        os.makedirs(os.path.join(OURS, "poismf"))
        run([sys.executable, "setup_b200.py", "build_ext", "--build-lib", OURS, "--build-temp", os.path.join(TMP, "_b200_tmp")],
            cwd=TMP, env=env)
        shutil.copy(os.path.join(STOCK, "poismf", "__init__.py"), os.path.join(OURS, "poismf", "__init__.py"))
    print("stock :", sorted(os.listdir(os.path.join(STOCK, "poismf"))))
    print("b200  :", sorted(os.listdir(os.path.join(OURS, "poismf"))))


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main(force="--force" in sys.argv)
