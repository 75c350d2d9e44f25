"""The reference's narrowphase tests (/root/reference/eggshell/collision.cc:527-809) restated
against the oracle with seeded numpy randomness and smaller trip counts."""
import numpy as np

from oracle import pyoracle as O


def _unit(v):
    return v / np.linalg.norm(v)


def _random_box(rng, axis1=None):
    """SetRandomBox, collision.cc:497-520."""
    c = rng.uniform(-1, 1, 3) * 0.5
    a0 = _unit(axis1 if axis1 is not None else rng.uniform(-1, 1, 3))
    a1 = rng.uniform(-1, 1, 3)
    a1 = _unit(a1 - (a0 @ a1) * a0)
    a2 = np.cross(a0, a1)
    return c, np.stack([a0, a1, a2], axis=1), rng.uniform(0, 1, 3)


def _quat_box(rng):
    q = _unit(rng.uniform(0, 1, 4) - 0.5)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return rng.uniform(-1, 1, 3), R, np.abs(rng.uniform(-1, 1, 3))


def test_line_closest_approach():                      # collision.cc:527-549
    rng = np.random.default_rng(10)
    for _ in range(1000):
        pa, pb = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
        ua, ub = _unit(rng.uniform(-1, 1, 3)), _unit(rng.uniform(-1, 1, 3))
        al, be = O.line_closest_approach(pa, ua, pb, ub)
        d = (pa + al * ua) - (pb + be * ub)
        assert abs(ua @ d) < 1e-9 and abs(ub @ d) < 1e-9
        for fa, fb in ((1.01, 1), (0.99, 1), (1, 1.01), (1, 0.99)):
            assert np.linalg.norm((pa + al * fa * ua) - (pb + be * fb * ub)) > np.linalg.norm(d)


def test_intersect_line_segment_and_line():            # collision.cc:551-577
    rng = np.random.default_rng(11)
    for _ in range(5000):
        p1, p2, n = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        d = rng.uniform(0, 1) - 0.5
        hit, p = O.intersect_line_segment_and_line(p1, p2, n, d)
        v1, v2 = n @ p1 + d, n @ p2 + d
        if hit:
            assert (v1 > 0 and v2 < 0) or (v1 < 0 and v2 > 0)
            assert abs(n @ p + d) < 1e-9
            assert np.linalg.norm(_unit(p - p1) + _unit(p - p2)) < 1e-6
        else:
            assert (v1 >= 0 and v2 >= 0) or (v1 <= 0 and v2 <= 0)


def test_clip_polygon_by_half_space():                 # collision.cc:579-633
    rng = np.random.default_rng(12)
    for i in range(5000):
        if i <= 2:
            poly = np.array([[0, 0], [0, 1], [1, 1], [1, 0]], dtype=float)
            n, d = np.array([0.0, 1.0]), -(i - 1) * 1e-12
        else:
            poly = rng.uniform(0, 1, (rng.integers(3, 8), 2)) - 0.5
            n, d = rng.uniform(-1, 1, 2), rng.uniform(0, 1) * 2 - 1
        new = O.clip_polygon(poly, n, d)
        assert (len(new) > 0) == bool(np.any(poly @ n + d > 0))
        if len(new):
            assert 3 <= len(new) <= 2 * len(poly)
            assert np.all(new @ n + d >= -1e-12)
            assert np.all(np.linalg.norm(new - np.roll(new, -1, axis=0), axis=1) > 1e-12)
        if i <= 2:
            assert len(new) == 4
            if i < 2:
                assert np.array_equal(new, poly)


def test_intersect_box_and_rectangle():                # collision.cc:635-688
    rng = np.random.default_rng(13)
    for i in range(4000):
        cB, RB, hB = _random_box(rng)
        cR, RR, hR = _random_box(rng, RB[:, 0] if (i & 1) == 0 else None)
        if (i & 2) == 0:
            RR = np.stack([RR[:, 1], RR[:, 0], -RR[:, 2]], axis=1)
        if (i & 15) == 0:
            i1 = rng.integers(0, 3); i2 = rng.integers(0, 2)
            if i2 == i1:
                i2 += 1
            RR = np.stack([RB[:, i1], RB[:, i2], np.cross(RB[:, i1], RB[:, i2])], axis=1)
        hR = hR.copy(); hR[2] = 0
        poly = O.intersect_box_rect(cB, RB, hB, cR, RR, hR)
        assert (len(poly) == 0) == O.boxes_separated(cB, RB, hB, cR, RR, hR)
        assert len(poly) == 0 or len(poly) >= 3
        for pt in poly:
            q1 = cR + RR @ np.array([pt[0], pt[1], 0.0])
            q2 = RB.T @ (q1 - cB)
            assert np.all(np.abs(q2) < hB + 1e-9)
            if not np.any(np.abs(q2) > hB - 1e-9):
                assert abs(abs(pt[0]) - hR[0]) < 1e-9 and abs(abs(pt[1]) - hR[1]) < 1e-9


def _face_pseudo_distance(c, R, h, p):                 # collision.cc:484-489
    return np.max(np.abs(R.T @ (p - c)) / h) - 1


def test_collide_boxes():                              # collision.cc:690-807
    rng = np.random.default_rng(14)
    codes = set()
    for it in range(6000):
        c1, R1, h1 = _quat_box(rng)
        c2, R2, h2 = _quat_box(rng)
        aligned = False
        if it % 5 == 0:
            i = it // 5
            c2, R2, h2 = _random_box(rng, R1[:, 0])
            if (i & 3) == 1:
                R2 = np.stack([R2[:, 1], R2[:, 0], -R2[:, 2]], axis=1)
            elif (i & 3) == 2:
                i1 = rng.integers(0, 3); i2 = rng.integers(0, 2)
                if i2 == i1:
                    i2 += 1
                R2 = np.stack([R1[:, i1], R1[:, i2], np.cross(R1[:, i1], R1[:, i2])], axis=1)
                aligned = True
        sep = O.boxes_separated(c1, R1, h1, c2, R2, h2)
        hit, code, depth, axis, cts = O.collide_boxes(c1, R1, h1, c2, R2, h2)
        assert sep == (not hit) and (not hit) == (code == 0) and (not hit) == (len(cts) == 0)
        if not hit:
            continue
        codes.add(code)
        assert abs(np.linalg.norm(axis) - 1) < 1e-9 and depth >= -1e-9
        h99, _, _, _, c99 = O.collide_boxes(c1 - 0.99 * depth * axis, R1, h1, c2, R2, h2)
        assert h99 and np.all(c99[:, 3:6] @ axis > 0)
        h101 = O.collide_boxes(c1 - 1.01 * depth * axis, R1, h1, c2, R2, h2)[0]
        assert not h101
        assert np.all(cts[:, 6] >= -1e-9)
        assert np.all(np.abs(np.linalg.norm(cts[:, 3:6], axis=1) - 1) < 1e-9)
        if 1 <= code <= 3:
            for ct in cts:
                assert abs(_face_pseudo_distance(c2, R2, h2, ct[:3])) < 1e-9
                assert abs(_face_pseudo_distance(c1, R1, h1, ct[:3] + ct[3:6] * ct[6])) < 1e-9
        elif 4 <= code <= 6:
            for ct in cts:
                assert abs(_face_pseudo_distance(c1, R1, h1, ct[:3])) < 1e-9
                assert abs(_face_pseudo_distance(c2, R2, h2, ct[:3] - ct[3:6] * ct[6])) < 1e-9
        elif 7 <= code <= 15:
            assert len(cts) == 1 and np.array_equal(cts[0, 3:6], axis)
        elif code == 16:
            assert len(cts) == 1 and np.array_equal(cts[0, :3], c2)
        if aligned and 1 <= code <= 6:
            assert len(cts) == 4
    assert {1, 2, 3, 4, 5, 6}.issubset(codes) and any(7 <= c <= 15 for c in codes)


def test_collide_box_and_ground_order_and_threshold():  # collision.cc:408-432
    R = np.eye(3)
    side = np.array([0.3, 0.3, 0.3])
    # resting exactly on the ground: v.z == 0 is NOT a contact (strict <)
    assert len(O.collide_box_ground([0, 0, 0.15], R, side)) == 0
    c = O.collide_box_ground([0, 0, 0.149], R, side)
    assert len(c) == 4
    # nested x,y,z order with z innermost: the four z=-1 vertices in (x,y) = (-,-),(-,+),(+,-),(+,+)
    assert np.allclose(c[:, 0], [-0.15, -0.15, 0.15, 0.15]) and np.allclose(c[:, 1], [-0.15, 0.15, -0.15, 0.15])
    assert np.allclose(c[:, 3:6], [0, 0, 1]) and np.allclose(c[:, 6], 0.001)
    assert len(O.collide_box_ground([0, 0, -1], R, side)) == 8
