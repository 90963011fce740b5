"""Recipe for `oracle/_ref/`: the reference's OWN Python modules for the hot path, unmodified, so that the GPU box
(which has no /root/reference) can run them -- as the reference arm of bench.py (`--impl reference`,
`cpu_baseline.kind == "reference"`) and as a second checker beside the oracle's restatement.

    python oracle/build_ref.py            # in the build container; __graft_entry__.build() calls it too

Nothing is edited: the files are copied byte for byte from /root/reference into oracle/_ref/ (git-ignored, so the
reference's sources never enter this repository's history; not gpurun-ignored, so the directory travels to the GPU box
like a built .so) and their SHA-256 digests are written to oracle/_ref/MANIFEST.json.  The modules import Pyro 1.3.0,
keras and matplotlib, none of which is installed in this image: they run against `oracle/pyro_shim` (a restatement of the
four Pyro entry points the hot path uses; plotting and dataset downloads become no-ops), exactly as
tests/golden/make_golden.py runs them to produce the golden vectors.

This is test / measurement infrastructure: only tests/, bench.py's reference arm and cpu_baseline leg may import it."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("RBNN_REFERENCE_DIR", "/root/reference")
DEST = os.path.join(HERE, "_ref")
# the hot path and what it imports
MODULES = ["model_nn.py", "model_bnn.py", "model_ensemble.py", "lossGradients.py", "adversarialAttacks.py", "utils.py",
           "savedir.py", "grid_search_halfMoons.py"]


def build(verbose=True):
    """Returns True when oracle/_ref holds the reference's modules (freshly copied, or already there)."""
    if not os.path.isdir(REFERENCE):
        ok = os.path.exists(os.path.join(DEST, "MANIFEST.json"))
        if verbose:
            print("oracle/_ref: %s is not present here; %s" % (REFERENCE, "using the copy that travelled with the tree"
                                                               if ok else "no reference copy available"))
        return ok
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for name in MODULES:
        src = os.path.join(REFERENCE, name)
        if not os.path.exists(src):
            continue
        shutil.copyfile(src, os.path.join(DEST, name))
        with open(src, "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REFERENCE, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: copied %d reference modules from %s" % (len(manifest), REFERENCE))
    return True


def import_reference():
    """(pyro shim, model_bnn, lossGradients, adversarialAttacks) of the reference copy, or None when oracle/_ref is absent.
    The shim and the copy are put at the FRONT of sys.path: never call this from product code."""
    if not os.path.exists(os.path.join(DEST, "MANIFEST.json")):
        return None
    for p in (DEST, os.path.join(HERE, "pyro_shim")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for clash in ("lossGradients", "adversarialAttacks", "model_bnn", "model_nn", "model_ensemble", "utils", "savedir"):
        mod = sys.modules.get(clash)
        if mod is not None and not getattr(mod, "__file__", "").startswith(DEST):
            del sys.modules[clash]
    import pyro
    import model_bnn
    import lossGradients
    import adversarialAttacks
    return pyro, model_bnn, lossGradients, adversarialAttacks


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
