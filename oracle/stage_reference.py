"""TEST INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference sources into oracle/_ref/reference/.

    python oracle/stage_reference.py            # copy + verify against oracle/reference_manifest.json
    python oracle/stage_reference.py --manifest # (re)write the manifest from /root/reference

Why: /root/reference does not exist on the GPU box.  oracle/_ref/ is git-ignored (the reference's text never
enters this repository's history) but it is NOT gpurun-ignored, so whatever this recipe puts there travels with
the snapshot, like a built .so.  Two consumers, both checkers / baselines, never the product:
  * tests/test_gpu_reference_text.py executes the reference's own `class cchess_main` text (main.py:1118-1554)
    over cchess_zero_b200's GameBoard / MCTS_tree / policy_value_network on the GPU;
  * bench.py --impl reference (and the cpu_baseline leg) runs the reference's own cchess_main.selfplay() with
    search_threads=16 on the host cores (oracle/ref_cpu_arm.py).
The manifest (sha256 per file, committed) is what proves that the staged text is the unmodified reference.
__graft_entry__.build() calls stage() whenever /root/reference is present."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CCHESS_REFERENCE_DIR", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
MANIFEST = os.path.join(HERE, "reference_manifest.json")
# main.py star-imports the two network modules at line 18-19 (they only touch TensorFlow inside their constructors)
FILES = ["main.py", "policy_value_network.py", "policy_value_network_gpus.py"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def write_manifest():
    m = {f: _sha(os.path.join(SRC, f)) for f in FILES}
    with open(MANIFEST, "w") as f:
        json.dump(m, f, indent=1, sort_keys=True)
        f.write("\n")
    return m


def manifest():
    with open(MANIFEST) as f:
        return json.load(f)


def staged_dir():
    """Directory holding a verified copy of the reference (the live tree here, the staged copy on the GPU box); None if neither."""
    for d in (SRC, DST):
        if all(os.path.isfile(os.path.join(d, f)) for f in FILES):
            return d
    return None


def verify(d=None):
    """True when every file under `d` hashes to the committed manifest (= byte-identical to the reference)."""
    d = d or staged_dir()
    if d is None or not os.path.isfile(MANIFEST):
        return False
    m = manifest()
    return all(_sha(os.path.join(d, f)) == m[f] for f in FILES)


def stage():
    if not os.path.isdir(SRC):
        return None
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    if not verify(DST):
        raise RuntimeError("staged reference does not match oracle/reference_manifest.json")
    return DST


if __name__ == "__main__":
    if "--manifest" in sys.argv:
        print(json.dumps(write_manifest(), indent=1))
    print(stage())
