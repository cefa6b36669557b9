"""sonicsim_b200 - B200-native moving-source acoustic renderer (hot path of JusperLee/SonicSim).

Drop-in modules:  sonicsim_b200.SonicSim_moving, sonicsim_b200.SonicSim_audio
Batch API:        sonicsim_b200.render (Renderer, convolve_moving, render_scene)
C ABI:            include/sonicsim_b200.h  (sonicsim_b200/csrc/libsonicsim_b200.so)
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "SonicSim_moving", "render"]
__version__ = "0.1.0"


def install_dropin():
    """Make `import SonicSim_moving` / `import SonicSim_audio` resolve to this package's modules
    (what SonicSet.py:19-20 imports).  See INTEGRATION.md."""
    import sys
    from . import SonicSim_moving
    sys.modules["SonicSim_moving"] = SonicSim_moving
    try:
        from . import SonicSim_audio
        sys.modules["SonicSim_audio"] = SonicSim_audio
    except ImportError:
        pass
