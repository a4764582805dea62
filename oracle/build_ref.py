"""ORACLE — TEST INFRASTRUCTURE ONLY.  Recipe for oracle/_ref: the UNMODIFIED reference modules of the hot path, made importable
on a box without /root/reference.

The reference is pure Python, so "building" it means staging the three modules the path lives in, byte for
byte, from where they lie under /root/reference/src into oracle/_ref/ — a git-ignored directory (never part of the history;
reference sources are not copied into the repo) that travels to the GPU box with the repo snapshot like the built .so files:

    dynamics/gnn/model.py        DynamicsPredictor                      (model.py:63-313)
    dynamics/dataset/graph.py    construct_edges_from_states[_batch]    (graph.py:38-156)
    dynamics/utils.py            pad_torch, truncate_graph              (utils.py:37-46, 127-137)

plus two stub packages for imports those modules make at top level but never use on this path and that the image lacks
(`dgl.geometry.farthest_point_sampler`, `moviepy.editor`; graph.py:5, utils.py:6,8), and a MANIFEST.json with the sha256 of
every staged file.  `__graft_entry__.build()` runs this when /root/reference is present; the GPU box only uses what was staged.

    python oracle/build_ref.py            # stage (idempotent)
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")
FILES = ["dynamics/gnn/model.py", "dynamics/dataset/graph.py", "dynamics/utils.py"]
STUBS = {
    "dgl/__init__.py": "# stub: dgl is absent from the image and unused on the dynamics path\n",
    "dgl/geometry.py": "def farthest_point_sampler(*a, **k):\n    raise RuntimeError('dgl stub: farthest_point_sampler is not on the dynamics path')\n",
    "moviepy/__init__.py": "# stub: moviepy is absent from the image and unused on the dynamics path\n",
    "moviepy/editor.py": "# stub\n",
}


def available() -> bool:
    return os.path.exists(os.path.join(OUT, "MANIFEST.json"))


def build(force: bool = False) -> bool:
    """Stages oracle/_ref from /root/reference/src.  Returns True when oracle/_ref is usable afterwards."""
    if not os.path.isdir(REF_SRC):
        return available()
    if available() and not force:
        man = json.load(open(os.path.join(OUT, "MANIFEST.json")))
        if all(_sha(os.path.join(REF_SRC, f)) == man["files"].get(f) for f in FILES):
            return True
    shutil.rmtree(OUT, ignore_errors=True)
    man = {"source": REF_SRC, "files": {}}
    for f in FILES:
        dst = os.path.join(OUT, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, f), dst)
        man["files"][f] = _sha(dst)
    for f, text in STUBS.items():
        dst = os.path.join(OUT, "_stubs", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, "w").write(text)
    json.dump(man, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    return True


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


if __name__ == "__main__":
    print("oracle/_ref staged" if build(force=True) else "no /root/reference here and nothing staged")
