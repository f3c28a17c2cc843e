"""B200-native batched Schwarzschild null-geodesic tracer (drop-in for the curvedpy call of
bldevries/blackhole_geodesic_calculator's render engines).  See DESIGN.md."""
from .api import (CAPTURED, ESCAPED, LAMBDA_EXHAUSTED, START_INSIDE_HOLE, STATUS_NAMES, STEP_FAILED, make_params,
                  trace, trace_device)

__all__ = ["trace", "trace_device", "make_params", "ESCAPED", "CAPTURED", "START_INSIDE_HOLE",
           "LAMBDA_EXHAUSTED", "STEP_FAILED", "STATUS_NAMES"]
__version__ = "0.1.0"
