"""B200-native batched Schwarzschild null-geodesic tracer (drop-in for the curvedpy call of
bldevries/blackhole_geodesic_calculator's render engines).  See DESIGN.md."""
from .api import (CAPTURED, ESCAPED, LAMBDA_EXHAUSTED, MISSED_SPHERE, START_INSIDE_HOLE, STATUS_NAMES, STEP_FAILED,
                  generate_rays, make_camera, make_params, trace, trace_camera, trace_device)

__all__ = ["trace", "trace_device", "trace_camera", "generate_rays", "make_camera", "make_params", "MISSED_SPHERE", "ESCAPED", "CAPTURED", "START_INSIDE_HOLE",
           "LAMBDA_EXHAUSTED", "STEP_FAILED", "STATUS_NAMES"]
__version__ = "0.2.0"
