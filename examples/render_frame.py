"""Render one Schwarzschild frame without Blender: camera rays generated on the device, traced to the sphere exit with
the in-flight accretion-disk event, shaded on the host with a procedural sky and disk - the three consumers of the
reference's engines (background lookup RRE.py:366-378, black shadow LIM.py:308-309, disk LIM.py:413-438) in ~60 lines.

    python examples/render_frame.py out.png [width] [spp] [rtol]

rtol defaults to 1e-6 (atol = rtol / 1000) for clean checker edges; the reference's own default, rtol = 1e-3, leaves
exit directions uncertain by a few 1e-3 rad, visible as ragged edges in its renders as well.
"""
import os
import struct
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blackhole_geodesic_calculator_b200 import api, raygen  # noqa: E402


def write_png(path, rgb):
    h, w, _ = rgb.shape
    raw = b"".join(b"\x00" + rgb[y].tobytes() for y in range(h))
    chunk = lambda tag, data: struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "frame.png"
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    rtol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-6
    h, M, R = w, 1.0, 60.0
    cam_pos = (120.0, -80.0, 18.0)
    cam = api.make_camera(cam_pos, raygen.look_at_rotation(cam_pos), w, h, 0.26, 0.26, seed=42, jitter="philox")
    n = spp * w * h
    t0 = time.perf_counter()
    pos, d, hit = api.generate_rays(cam, n, R)                       # device tensors, loop order s -> y -> x
    res = api.trace(pos, d, M=M, r_sphere=R, rtol=rtol, atol=rtol * 1e-3, image_width=w, disk=(6.0, 22.0))
    exit_pos, exit_dir, status, disk_xy = (t.cpu().numpy() for t in res)   # device tensors in, device tensors out
    dt = time.perf_counter() - t0
    # sky: equirectangular checkerboard tinted by direction, as background_hit would sample a texture
    phi = np.arctan2(exit_dir[:, 1], exit_dir[:, 0])
    theta = np.arccos(np.clip(exit_dir[:, 2], -1, 1))
    check = ((np.floor(phi / np.pi * 18) + np.floor(theta / np.pi * 18)) % 2).astype(np.float64)
    sky = np.stack([0.15 + 0.55 * check * (0.5 + 0.5 * np.cos(phi)), 0.15 + 0.45 * check,
                    0.25 + 0.55 * check * (0.5 + 0.5 * np.sin(phi))], axis=1)
    rgb = np.where((status == 0)[:, None], sky, 0.0)                 # captured rays stay black
    # disk: first crossing of z = 0 inside the annulus; brightness falls with radius, spokes show the lensing
    on_disk = np.isfinite(disk_xy[:, 0])
    rr = np.hypot(disk_xy[on_disk, 0], disk_xy[on_disk, 1])
    az = np.arctan2(disk_xy[on_disk, 1], disk_xy[on_disk, 0])
    glow = (6.0 / rr) ** 1.5 * (0.75 + 0.25 * np.sign(np.sin(12 * az)))
    rgb[on_disk] = np.stack([np.minimum(1.0, 1.6 * glow), np.minimum(1.0, 0.9 * glow), 0.35 * glow], axis=1)
    img = rgb.reshape(spp, h, w, 3).mean(axis=0)
    write_png(out, (np.clip(img, 0, 1) ** (1 / 2.2) * 255).astype(np.uint8))
    print(f"{out}: {w}x{h} x {spp} spp = {n} rays generated + traced + copied to the host in {dt * 1e3:.1f} ms; "
          f"captured {100 * (status == 1).mean():.2f} %, on disk {100 * on_disk.mean():.2f} %")


if __name__ == "__main__":
    main()
