"""ORACLE (test infrastructure): vendor the reference's own model code into `oracle/_ref` so it can run on the GPU box.

    python oracle/build_ref.py            # in the build container (needs /root/reference, read-only)

`/root/reference` does not exist on the GPU box, and the reference is pure Python: the part of its package that holds
the path of SURVEY.md section 8a — `metamorph/{__init__,constants,mm_utils}.py` and `metamorph/model/**` — is copied
verbatim, file by file, into `oracle/_ref/metamorph/`. `oracle/_ref/` is listed in .gitignore (reference sources never
enter this repository's history) and NOT in .gpurunignore, so the copy travels with the snapshot exactly like the built
`_C.so`. Its arithmetic lives in `transformers` / `torch`, which the image provides (5.5 / 2.11 instead of the pinned
4.45 / 2.2: SURVEY.md section 8c lists the drift; `oracle/ref_model.py` holds the two shims).
`oracle/ref_bench.py` times it on the host cores (bench.py `--impl reference` and the `cpu_baseline` legs).
`metamorph/train/` is not copied: it cannot be imported offline and is not on the path.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["metamorph/__init__.py", "metamorph/constants.py", "metamorph/mm_utils.py"]
TREES = ["metamorph/model"]


def main(src=SRC, dst=DST) -> int:
    if not os.path.isdir(os.path.join(src, "metamorph")):
        print(f"build_ref: {src} not present (GPU box?) - keeping whatever {dst} holds", file=sys.stderr)
        return 0 if os.path.isdir(os.path.join(dst, "metamorph")) else 1
    files = list(FILES)
    for tree in TREES:
        for root, _, names in os.walk(os.path.join(src, tree)):
            for n in sorted(names):
                if n.endswith(".py"):
                    files.append(os.path.relpath(os.path.join(root, n), src))
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    manifest = {}
    for rel in sorted(files):
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
        with open(out, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()[:16]
    head = None
    try:
        with open(os.path.join(src, ".git", "HEAD")) as fh:
            head = fh.read().strip()
    except OSError:
        pass
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "git_head": head, "files": manifest,
                   "note": "verbatim copies; never committed (oracle/_ref/ is git-ignored)"}, fh, indent=1)
    print(f"build_ref: {len(manifest)} files -> {dst}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
