"""pointcloudmatters_b200 -- B200-native (sm_100a) hot path of HaoyiZhu/PointCloudMatters.

Scope: the point-cloud behaviour-cloning training step.  The package holds
the CUDA kernels + C ABI (`csrc/`, `libpcm_b200.so`, `include/pcm_b200.h`) and the host-side
mirror of the reference interfaces for that path (`pointops`, encoder / policy modules, the
training step).  There is no CPU fallback: importing the kernel-backed modules without the built
shared library raises.
"""
__version__ = "0.1.0"


def install_as_pointops() -> None:
    """Register the drop-in package under the reference's import name (`import pointops`)."""
    import sys

    from . import pointops as _p

    sys.modules["pointops"] = _p
