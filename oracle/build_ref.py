"""Compiles the reference's two native modules FROM WHERE THEY LIE under /root/reference into
oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun):

    selene_sdk/sequences/_sequence.pyx          -> oracle/_ref/_sequence.<abi>.so
    selene_sdk/targets/_genomic_features.pyx    -> oracle/_ref/_genomic_features.<abi>.so

No reference source is copied into the repository: Cython writes its generated C into oracle/_ref/
as a build intermediate and it is deleted after gcc has run.  These binaries are test
infrastructure: tests/test_oracle.py uses them to validate oracle/selene_oracle.py, and bench.py's
CPU legs may time them (cpu_baseline.kind == "reference" for the stages they cover).
"""
import os
import subprocess
import sys
import sysconfig

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SOURCES = {
    "_sequence": "selene_sdk/sequences/_sequence.pyx",
    "_genomic_features": "selene_sdk/targets/_genomic_features.pyx",
}


def main():
    import numpy as np
    if not os.path.isdir(REF):
        print("build_ref: %s not present, keeping whatever is in %s" % (REF, OUT))
        return 0
    os.makedirs(OUT, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    inc = sysconfig.get_paths()["include"]
    for name, rel in SOURCES.items():
        src = os.path.join(REF, rel)
        c_file = os.path.join(OUT, name + ".c")
        so = os.path.join(OUT, name + suffix)
        subprocess.run([sys.executable, "-m", "cython", "-3", src, "-o", c_file], check=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-w", "-I", inc, "-I", np.get_include(), c_file, "-o", so],
                       check=True)
        os.remove(c_file)
        print("build_ref: built", so)
    with open(os.path.join(OUT, "__init__.py"), "w") as fh:
        fh.write("# reference-native modules compiled by oracle/build_ref.py (test infrastructure)\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
