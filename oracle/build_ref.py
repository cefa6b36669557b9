"""oracle/_ref: the UNMODIFIED reference module for the CPU arm of bench.py.

TEST / BENCH INFRASTRUCTURE.  The reference is pure Python with no build step, so "building" it is
copying SonicSim-SonicSet/SonicSim_moving.py byte for byte out of /root/reference into oracle/_ref/
(git-ignored, NOT gpurun-ignored: it travels to the GPU box like the built .so files) and writing the
three-line stand-in for `SonicSim_rir`, which the module imports for type names only
(SonicSim_moving.py:12; the real one needs habitat_sim / magnum).  No reference source is committed.

    python oracle/build_ref.py            # authoring container; __graft_entry__.build() calls it

`load()` imports the copy (under a private module name, so it never shadows the drop-in).
"""
import hashlib
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/SonicSim-SonicSet/SonicSim_moving.py"
REF_DIR = os.path.join(HERE, "_ref")
REF_DST = os.path.join(REF_DIR, "SonicSim_moving.py")
STUB = ("# stand-in written by oracle/build_ref.py: SonicSim_moving.py:12 imports these names for annotations only\n"
        "class Receiver: pass\nclass Source: pass\nclass Scene: pass\n")


def build(verbose=True):
    """Copy the reference module into oracle/_ref/.  Returns True when oracle/_ref is usable afterwards."""
    if os.path.isfile(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        shutil.copyfile(REF_SRC, REF_DST)
        with open(os.path.join(REF_DIR, "SonicSim_rir.py"), "w") as f:
            f.write(STUB)
        with open(os.path.join(REF_DIR, "SHA256"), "w") as f:
            f.write(hashlib.sha256(open(REF_DST, "rb").read()).hexdigest() + "  SonicSim_moving.py\n")
        if verbose:
            print("oracle/_ref: copied", REF_SRC)
    elif verbose:
        print("oracle/_ref: /root/reference absent, keeping what is there (%s)" % ("present" if available() else "nothing"))
    return available()


def available():
    return os.path.isfile(REF_DST) and os.path.isfile(os.path.join(REF_DIR, "SonicSim_rir.py"))


_mod = None


def load():
    """The unmodified reference `SonicSim_moving` module from oracle/_ref (None if it was never built)."""
    global _mod
    if _mod is not None or not available():
        return _mod
    saved = sys.modules.get("SonicSim_rir")
    stub = types.ModuleType("SonicSim_rir")
    exec(open(os.path.join(REF_DIR, "SonicSim_rir.py")).read(), stub.__dict__)
    sys.modules["SonicSim_rir"] = stub
    try:
        spec = importlib.util.spec_from_file_location("_sonicsim_ref_moving", REF_DST)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["SonicSim_rir"] = saved
        else:
            sys.modules.pop("SonicSim_rir", None)
    _mod = mod
    return mod


if __name__ == "__main__":
    ok = build()
    print("usable:", ok)
