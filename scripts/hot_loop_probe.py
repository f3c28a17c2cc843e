"""Pure hot-loop throughput of the trace kernel: every lane integrates one long ray in lockstep (max_step-limited, no
events, no refills), so service time and lane idling vanish and what remains is the RK45 attempt itself.

Prints SM cycles per warp-attempt per scheduler; the FP64-pipe floor is 2 x (FP64 instructions per attempt).
"""
import json, sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api

def main():
    dev = torch.device("cuda", 0)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    out = {}
    for mode in ("parity", "plane"):
        for rays_per_lane in (1,):
            n = sms * 512 * rays_per_lane
            rng = np.random.default_rng(1)
            d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
            pos = d * 100.0 + rng.normal(size=(n, 3))
            dirs = d + 0.3 * rng.normal(size=(n, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
            p = torch.from_numpy(pos).to(dev); k = torch.from_numpy(dirs).to(dev)
            kw = dict(M=1.0, r_sphere=np.inf, rtol=1e-3, atol=1e-6, max_step=0.5, lambda_max=400.0, mode=mode,
                      return_counters=True)
            res = api.trace(p, k, **kw)
            torch.cuda.synchronize()
            att = res[3][0].double().mean().item()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); api.trace(p, k, **kw); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = min(ts) * 1e-3
            clk = 1.965e9
            out[mode] = {"rays": n, "attempts_per_ray": att, "ms": t * 1e3,
                         "cycles_per_warp_attempt_per_scheduler": t * clk / (att * rays_per_lane * 4)}
    print(json.dumps(out))

if __name__ == "__main__":
    main()
