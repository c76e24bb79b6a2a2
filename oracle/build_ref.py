"""Build recipe for oracle/_ref: compile the reference's Cython kernels where they lie.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product package
(dynetlsm_b200); only tests/, __graft_entry__.smoke()/build() and bench.py's CPU legs use it.

What it does: cythonizes the four ``.pyx`` files of the reference straight from
``/root/reference/dynetlsm`` (static_network_fast.pyx, directed_likelihoods_fast.pyx,
gaussian_likelihood_fast.pyx; flags equivalent to the reference's setup.py:114-125, -O3 -fPIC)
and writes ONLY build outputs (generated .c, .o, .so) into ``oracle/_ref/``.  No reference
source is copied into the repository.  ``oracle/_ref/`` is git-ignored but travels to the GPU
box, where the compiled kernels back ``bench.py --impl reference`` and the oracle cross-checks.

Run:  python oracle/build_ref.py            (no-op if /root/reference is absent)
"""
import os
import sys
import glob

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DYNETLSM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ["static_network_fast", "directed_likelihoods_fast", "gaussian_likelihood_fast",
           "forecast"]  # forecast.pyx is off-path; hdp_lpcm.py:18 imports it


def have_ref_build():
    return all(glob.glob(os.path.join(OUT, m + ".*.so")) for m in MODULES)


def build(force=False):
    src_dir = os.path.join(REF, "dynetlsm")
    if not os.path.isdir(src_dir):
        return have_ref_build()
    if have_ref_build() and not force:
        return True
    import numpy
    from setuptools import Extension
    from setuptools.dist import Distribution
    from Cython.Build import cythonize

    os.makedirs(OUT, exist_ok=True)
    exts = [
        Extension(
            m,
            sources=[os.path.join(src_dir, m + ".pyx")],
            include_dirs=[numpy.get_include()],
            extra_compile_args=["-O3", "-fPIC", "-w"],
            define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
        )
        for m in MODULES
    ]
    exts = cythonize(exts, build_dir=os.path.join(OUT, "build"), quiet=True,
                     compiler_directives={"language_level": 3})
    dist = Distribution({"ext_modules": exts})
    cmd = dist.get_command_obj("build_ext")
    cmd.build_lib = OUT
    cmd.build_temp = os.path.join(OUT, "build")
    cmd.ensure_finalized()
    cmd.run()
    return have_ref_build()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "unavailable (no reference sources and no prebuilt .so)")
