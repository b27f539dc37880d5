"""TEST INFRASTRUCTURE (authoring container only): builds the reference's own Python extension
`whatshap.core` from /root/reference in a scratch directory OUTSIDE this repository so that tests can
drive the real Cython classes.  Nothing is copied into the repo and the build does not travel to the
GPU box (a Python reference cannot travel; its outputs are the golden vectors under tests/golden/).

Recipe: copy src/ whatshap/ setup.py, stub _version.py, skip the polyphase solver
import (needs `pulp`), build_ext --inplace with the distribution compiler (/usr/bin/g++; /opt/gcc links
libstdc++ statically and the module crashes).  Result here: OK, about 45 s, 5 extensions.
"""
import os
import shutil
import subprocess
import sys

REF = os.environ.get("WHATSHAP_REF", "/root/reference")
OUT = os.environ.get("WHMEC_PYREF_DIR", "/tmp/whmec_pyref")


def build() -> str:
    """Returns the directory to put on sys.path, or '' when the reference tree is not available."""
    marker = os.path.join(OUT, "whatshap", ".built")
    if os.path.exists(marker):
        return OUT
    if not os.path.isdir(os.path.join(REF, "src")):
        return ""
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    for item in ("src", "whatshap", "setup.py"):
        src = os.path.join(REF, item)
        dst = os.path.join(OUT, item)
        shutil.copytree(src, dst) if os.path.isdir(src) else shutil.copy(src, dst)
    for root, dirs, files in os.walk(OUT):
        for name in dirs + files:
            os.chmod(os.path.join(root, name), 0o755)
    with open(os.path.join(OUT, "whatshap", "_version.py"), "w") as f:
        f.write('version = "0+oracle"\n')
    init = os.path.join(OUT, "whatshap", "__init__.py")
    text = open(init).read().replace("    import whatshap.polyphase.solver  # noqa", "    pass")
    open(init, "w").write(text)
    env = dict(os.environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++", CFLAGS="-O2", CXXFLAGS="-O2")
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace", "-j", "8"], cwd=OUT, env=env, check=True,
                   capture_output=True)
    open(marker, "w").write("ok\n")
    return OUT


if __name__ == "__main__":
    print(build() or "reference tree not found")
