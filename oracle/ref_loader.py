"""Import the UNMODIFIED reference modules from /root/reference (authoring container only).

TEST INFRASTRUCTURE - not a product path.  Only `oracle/make_golden.py` and the
`not gpu` tests that are explicitly skipped when /root/reference is absent may
use this.  Nothing that runs on the GPU box imports it.

`SonicSim_moving` does `from SonicSim_rir import Receiver, Source, Scene`
(SonicSim-SonicSet/SonicSim_moving.py:12) only for type names, and
`SonicSim_rir` needs habitat_sim / magnum / matplotlib which are not installable
here.  We inject stub modules for those names so that the hot-path functions
(SonicSim_moving.py:15-125, SonicSim_audio.py:17-47) import and run unmodified.
"""
import importlib
import os
import sys
import types

REF_DIR = "/root/reference/SonicSim-SonicSet"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "SonicSim_moving.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load(want_audio: bool = False):
    """Returns the reference `SonicSim_moving` module (and `SonicSim_audio` if asked)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_DIR)
    saved = {k: sys.modules.get(k) for k in
             ("SonicSim_rir", "SonicSim_moving", "SonicSim_audio", "pyloudnorm",
              "matplotlib", "matplotlib.pyplot")}
    try:
        _stub("SonicSim_rir", Receiver=object, Source=object, Scene=object,
              render_rir_parallel=None)
        if want_audio:
            if "matplotlib" not in sys.modules or sys.modules["matplotlib"] is None:
                mpl = _stub("matplotlib")
                mpl.pyplot = _stub("matplotlib.pyplot")
            try:
                import pyloudnorm  # noqa: F401
            except Exception:
                _stub("pyloudnorm")
        sys.path.insert(0, REF_DIR)
        for k in ("SonicSim_moving", "SonicSim_audio"):
            sys.modules.pop(k, None)
        moving = importlib.import_module("SonicSim_moving")
        audio = importlib.import_module("SonicSim_audio") if want_audio else None
    finally:
        if REF_DIR in sys.path:
            sys.path.remove(REF_DIR)
        # do not leave reference/stub modules importable under the drop-in names
        for k in ("SonicSim_rir", "SonicSim_moving", "SonicSim_audio"):
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            elif k in ("pyloudnorm", "matplotlib", "matplotlib.pyplot"):
                sys.modules.pop(k, None)
    return (moving, audio) if want_audio else moving
