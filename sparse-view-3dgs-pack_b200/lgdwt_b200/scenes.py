"""Synthetic scenes and cameras — re-export of the numpy-only top-level module `lgdwt_scenes` (kept outside this
package so that a process which must not load the CUDA library, e.g. bench.py's reference arm, can still import the
generators)."""
from lgdwt_scenes import *  # noqa: F401,F403
from lgdwt_scenes import Camera, Scene, _projection, _random_unit_quaternions, _sigmoid  # noqa: F401
