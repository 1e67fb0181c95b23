"""raytrace3_b200 — B200-native SPPM render hot path behind the reference's
Scene/Obj/Light/Camera/RayTracer/SPPMRayTracer API (wondergo2017/raytrace3).

The compute lives in csrc/ (hand-written CUDA for sm_100a, C ABI in include/sppm_b200.h);
this package is the Python flavour of the host-side mirror used by tests and bench.py."""
from . import abi  # noqa: F401
from .api import (Context, RayTracer, SPPMRayTracer, SppmError, photon_calls, photon_range, rounds_per_call,  # noqa: F401
                  row_range)
from .scene import (Bezier3Obj, Camera, Light, Material, Obj, PlaneObj, Scene, SphereObj, SpotLight,  # noqa: F401
                    Texture, make_scene, reference_camera)
