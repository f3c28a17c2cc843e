"""CPU tests: host-side ray generation (the caller side of the path, RelativisticRenderEngine.py:185-230)."""
import math
import random

import numpy as np

from blackhole_geodesic_calculator_b200 import raygen


def reference_loop(width, height, samples, fov, rot, seed):
    """The reference's generator restated literally as the s -> y -> x Python loop (RRE.py:185-230)."""
    aspect = height / width
    dy = aspect / height
    dx = 1 / width
    random.seed(seed)
    out = []
    for s in range(samples):
        for y in range(height):
            for x in range(width):
                xr = fov * (x - int(width / 2)) / width
                yr = fov * (y - int(height / 2)) / height * aspect
                v = np.array([xr + dx * (random.random() - 0.5), yr + dy * (random.random() - 0.5), -1.0])
                v = rot @ v
                out.append(v / np.linalg.norm(v))
    return np.array(out)


def test_mt19937_stream_and_loop_order():
    rot = raygen.look_at_rotation(raygen.CFG_CAMERA_POS)
    a = reference_loop(12, 8, 2, 0.6, rot, 42)
    b = raygen.camera_rays(12, 8, 2, 0.6, 0.6, rot, 42, "mt19937")
    assert np.abs(a - b).max() < 1e-15
    # a slice of the stream equals the slice of the whole
    c = raygen.camera_rays(12, 8, 2, 0.6, 0.6, rot, 42, "mt19937", first_ray=37, n_rays=50)
    assert np.array_equal(c, b[37:87])


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    r = raygen.philox4x32_10(np.array([0], dtype=np.uint64), 0)
    assert [int(v[0]) for v in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_philox_slices_are_counter_based():
    a = raygen.camera_rays(32, 16, 2, jitter="philox")
    b = raygen.camera_rays(32, 16, 2, jitter="philox", first_ray=100, n_rays=300)
    assert np.array_equal(a[100:400], b)


def test_look_at_and_euler():
    rot = raygen.look_at_rotation((120.0, -80.0, 40.0))
    assert np.allclose(rot @ rot.T, np.eye(3), atol=1e-14) and np.isclose(np.linalg.det(rot), 1.0)
    fwd = rot @ np.array([0, 0, -1.0])
    assert np.allclose(fwd, -np.array([120.0, -80.0, 40.0]) / np.linalg.norm([120.0, -80.0, 40.0]))
    e = raygen.euler_xyz_rotation(0.3, -0.2, 1.1)
    assert np.allclose(e @ e.T, np.eye(3), atol=1e-14)
    assert np.allclose(raygen.euler_xyz_rotation(0, 0, math.pi / 2) @ [1, 0, 0], [0, 1, 0], atol=1e-15)


def test_sphere_entry_and_config_bundle():
    pos, d = raygen.config_bundle(64, 64, 1)
    assert pos.shape == (4096, 3)
    assert np.allclose(np.linalg.norm(pos, axis=1), 60.0, atol=1e-11)
    assert (np.sum(pos * d, axis=1) < 0).all()  # pointing inward
    p, hit = raygen.sphere_entry((200.0, 0, 0), np.array([[0, 1.0, 0], [-1.0, 0, 0]]), 60.0)
    assert list(hit) == [False, True] and np.allclose(p[1], [60.0, 0, 0])


def test_near_critical_bundle_has_requested_b():
    pos, d, b = raygen.near_critical_bundle(500)
    got = raygen.conserved_impact_parameter(pos, d, 1.0)
    assert np.abs(got - b).max() < 1e-12 and b.min() >= 5.0 and b.max() <= 5.4
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-14)


def test_conversions_round_trip_and_match_the_oracle_entry_map():
    """adapters.Conversions (curvedpy's name, RRE.py:289-291): xyz -> sph -> xyz is the identity, and the spherical
    components are the ones the oracle starts its integration from."""
    from blackhole_geodesic_calculator_b200 import adapters
    from oracle import schwarzschild_ref as R
    rng = np.random.default_rng(3)
    conv = adapters.Conversions()
    for _ in range(50):
        x, k = rng.normal(size=3) * 20.0, rng.normal(size=3)
        xs, ks = conv.convert_xyz_to_sph(x, k)
        x2, k2 = conv.convert_sph_to_xyz(xs, ks)
        assert np.allclose(x2, x, rtol=1e-13, atol=1e-13) and np.allclose(k2, k, rtol=1e-12, atol=1e-13)
        if hasattr(R, "xyz_to_sph"):
            xo, ko = R.xyz_to_sph(x, k)
            assert np.allclose(xs, xo, rtol=1e-13) and np.allclose(ks, ko, rtol=1e-12, atol=1e-14)
