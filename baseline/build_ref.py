"""Installs the UNMODIFIED reference (selene-sdk 0.6.0, /root/reference) into baseline/_ref/ for the
``bench.py --impl reference`` arm.  Authoring container only (the GPU box receives baseline/_ref with the
repository snapshot; it is git-ignored, not gpurun-ignored).

    python baseline/build_ref.py

Recipe (the one the base contract prescribes, with the two documented allowances):
  * /root/reference is read-only and the build writes Cython output next to the sources, so the install runs
    from a scratch copy under /tmp;
  * pyfaidx, pytabix, h5py and ruamel.yaml are not in the offline wheelhouse, so dependency resolution is
    skipped (``--no-deps``).  At run time those four imports are satisfied by the pure-python stand-ins of
    baseline/shims/ (FASTA slicing, interval lookup, stubs: no arithmetic of the path).
The reference's example model files (``models/*.py``, user-supplied architecture files in Selene's design,
loaded by path) are copied next to the package as baseline/_ref/selene_models/.
"""
import os
import shutil
import subprocess
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SCRATCH = "/tmp/selene_ref_src"


def main():
    if not os.path.isdir(REF):
        print("baseline/build_ref: %s not present, keeping whatever is in %s" % (REF, OUT))
        return 0
    if os.path.isdir(os.path.join(OUT, "selene_sdk")) and os.path.isdir(os.path.join(OUT, "selene_models")):
        print("baseline/build_ref: %s already installed" % OUT)
        return 0
    shutil.rmtree(SCRATCH, ignore_errors=True)
    shutil.copytree(REF, SCRATCH, ignore=shutil.ignore_patterns(".git"))
    shutil.rmtree(OUT, ignore_errors=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", OUT, SCRATCH]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout[-3000:])
        return r.returncode
    shutil.copytree(os.path.join(REF, "models"), os.path.join(OUT, "selene_models"))
    print("baseline/build_ref: installed selene-sdk into", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
