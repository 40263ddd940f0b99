"""Place an UNMODIFIED copy of the reference package under baseline/_ref (git-ignored).

The documented route -- ``pip install --no-index --no-build-isolation --find-links
/opt/wheelhouse --target baseline/_ref /root/reference`` -- fails in this image because the
reference's build backend (poetry-core, pyproject.toml:31-33) is not in the wheelhouse.  The
package is pure Python, so the install is a file copy of ``exauq/``; nothing is edited.  Its
dependency ``mogp-emulator`` cannot be installed (no network): ``oracle/mogp_emulator``
stands in for it in tests and in the CPU-baseline arm only.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("EXAUQ_REFERENCE", "/root/reference")


def install(verbose: bool = True) -> bool:
    if os.path.isdir(os.path.join(DEST, "exauq")):
        return True
    if not os.path.isdir(os.path.join(SRC, "exauq")):
        if verbose:
            print(f"reference not found at {SRC}; baseline/_ref not created")
        return False
    os.makedirs(DEST, exist_ok=True)
    pip = subprocess.run(
        [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
         "--find-links", "/opt/wheelhouse", "--target", DEST, SRC],
        capture_output=True, text=True)
    if pip.returncode == 0 and os.path.isdir(os.path.join(DEST, "exauq")):
        if verbose:
            print("installed with pip")
        return True
    shutil.copytree(os.path.join(SRC, "exauq"), os.path.join(DEST, "exauq"),
                    ignore=shutil.ignore_patterns("__pycache__"))
    if verbose:
        print("pip could not build the reference (no poetry-core); copied the pure-Python package instead")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
