"""Recipe that "installs" the reference for the baseline arms: copies the UNMODIFIED reference sources that the train-step
path and its driver need from /root/reference into the git-ignored oracle/_ref/fmov_pose/ (it travels to the GPU box with
the snapshot like the built .so; it never enters the repo's history), with a manifest of SHA-256 sums.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python with no setup.py, so there is nothing
to pip-install or compile: this copy IS the install.  Consumers: `bench.py --impl reference` / `cpu_baseline`
(`kind: "reference"`), the PyTorch-on-B200 arm, and tests/test_gpu_zzzz_exp_runner.py (the reference's own exp_runner.py
driving this package's models).  Nothing under fmov_pose_b200/ may import it.

    python oracle/build_ref.py          # in the build container, where /root/reference exists
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "fmov_pose")
WANT = ["exp_runner.py", "models", "utils", "confs", "LICENSE"]


def build(ref=None, quiet=False):
    ref = ref or os.environ.get("FMOV_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(ref, "exp_runner.py")):
        if not quiet:
            print(f"build_ref: {ref} not found (GPU box?): keeping whatever oracle/_ref already holds")
        return os.path.exists(os.path.join(DST, "exp_runner.py"))
    if os.path.exists(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for name in WANT:
        src = os.path.join(ref, name)
        if os.path.isdir(src):
            for root, _, files in os.walk(src):
                for f in files:
                    if f.endswith((".pyc",)) or "__pycache__" in root:
                        continue
                    rel = os.path.relpath(os.path.join(root, f), ref)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel))
                    manifest[rel] = hashlib.sha256(open(os.path.join(root, f), "rb").read()).hexdigest()
        elif os.path.exists(src):
            shutil.copyfile(src, os.path.join(DST, name))
            manifest[name] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": ref, "files": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print(f"build_ref: {len(manifest)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
