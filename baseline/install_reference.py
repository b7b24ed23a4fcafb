"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with gpurun).

    python baseline/install_reference.py

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>` - the contract's
   install.  The reference's own setup.py lists `packages=["desed_task"]` only (setup.py:9), so the wheel carries
   `desed_task/data_augm.py` and NONE of the sub-packages (`nnet`, `utils`, `dataio`, `evaluation`); the reference's recipes
   work around that with `pip install -e` (recipes/dcase2023_task4_baseline/conda_create_environment.sh:10).
2. The equivalent of that editable install for a directory that has to travel: the sub-package directories are added to
   baseline/_ref/desed_task/ from the same checkout, byte for byte.  Nothing under baseline/_ref is tracked by git.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")


def main():
    if not os.path.isdir(os.path.join(REF, "desed_task")):
        print("no reference checkout at %s: nothing to install" % REF)
        return 1
    tmp = tempfile.mkdtemp(prefix="refcopy_")
    src = os.path.join(tmp, "reference")
    shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "data", "PSDS_Eval"))      # /root/reference is read-only
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", DST, src], capture_output=True, text=True)
    print(r.stdout[-600:], r.stderr[-600:])
    if r.returncode != 0:
        return r.returncode
    pkg = os.path.join(DST, "desed_task")
    for sub in ("nnet", "utils", "dataio", "evaluation"):
        shutil.copytree(os.path.join(REF, "desed_task", sub), os.path.join(pkg, sub), dirs_exist_ok=True,
                        ignore=shutil.ignore_patterns("__pycache__"))
    # the recipe configs the benchmark's reference legs read (confs only)
    for rec in ("dcase2023_task4_baseline", "dcase2024_task4_baseline"):
        shutil.copytree(os.path.join(REF, "recipes", rec, "confs"), os.path.join(DST, "recipes", rec, "confs"),
                        dirs_exist_ok=True)
    bad = [f for f in ("data_augm.py", "nnet/CRNN.py", "nnet/CNN.py", "nnet/RNN.py", "utils/scaler.py")
           if not filecmp.cmp(os.path.join(REF, "desed_task", f), os.path.join(pkg, f), shallow=False)]
    assert not bad, bad
    shutil.rmtree(tmp, ignore_errors=True)
    print("installed the unmodified reference into", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
