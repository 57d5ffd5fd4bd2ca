"""Size-independent properties of the input-side geometry (hypothesis): the contour resampling always returns exactly the
requested number of points on the polygon's own edges, clockwise, starting at the vertex nearest the top-centre of the
extent; every flip is an involution; resize / flip commute with the ground-truth boxes the way the head assumes."""
import numpy as np
from hypothesis import given, settings, strategies as st

from lsnet_b200.datasets import PolygonMasks, contour
from lsnet_b200.datasets.transforms import bbox_flip, extreme_flip, keypoint_flip


@st.composite
def rings(draw, min_n=3, max_n=500):
    """Star-shaped (hence simple) rings of n vertices around a random centre, either orientation."""
    n = draw(st.integers(min_n, max_n))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.RandomState(seed)
    th = np.sort(rng.rand(n)) * 2 * np.pi + np.arange(n) * 1e-9          # strictly increasing angles
    if draw(st.booleans()):
        th = th[::-1]
    r = 5 + 60 * rng.rand(n)
    cx, cy = rng.uniform(80, 400, 2)
    return np.stack([cx + r * np.cos(th), cy + r * np.sin(th)], 1)


@settings(max_examples=60, deadline=None)
@given(rings(), st.sampled_from([8, 36, 128]), st.sampled_from([1, 4, 10]))
def test_unify_polygons_invariants(poly, num_points, spline):
    box = np.array([poly[:, 0].min(), poly[:, 1].min(), poly[:, 0].max(), poly[:, 1].max()])
    out = contour.unify_polygons([poly.reshape(-1).tolist()], box, num_points, spline)
    assert len(out) == 1 and out[0].shape == (2 * num_points,)
    ring = out[0].reshape(-1, 2)
    # on the polygon's extent (resampled points lie on its edges; the box fallback for filtered polygons is the extent)
    assert (ring[:, 0] >= box[0] - 1e-6).all() and (ring[:, 0] <= box[2] + 1e-6).all()
    assert (ring[:, 1] >= box[1] - 1e-6).all() and (ring[:, 1] <= box[3] + 1e-6).all()
    assert contour.signed_area(ring) <= 1e-9                         # never counter-clockwise (shapely's sense)
    tcx, tcy = (ring[:, 0].min() + ring[:, 0].max()) / 2, ring[:, 1].min()
    d = (ring[:, 0] - tcx) ** 2 + (ring[:, 1] - tcy) ** 2
    assert d[0] <= d.min() + 1e-9                                    # starts nearest the top-centre of its own extent


@settings(max_examples=60, deadline=None)
@given(rings(), st.sampled_from([16, 360, 1280]))
def test_uniformsample_count_and_edges(poly, n):
    s = contour.uniformsample(poly, n)
    assert s.shape == (n, 2)
    if len(poly) > n:                                                # a subsequence of the original vertices, in order
        idx = [int(np.where((poly == p).all(1))[0][0]) for p in s]
        assert idx == sorted(idx) and len(set(idx)) == n
    else:                                                            # every original vertex survives (weight 0 of its edge)
        assert all((np.abs(s - p).sum(1) < 1e-9).any() for p in poly)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from(['horizontal', 'vertical']))
def test_flips_are_involutions(seed, direction):
    rng = np.random.RandomState(seed)
    shape = (int(rng.randint(50, 900)), int(rng.randint(50, 1400)), 3)
    g = int(rng.randint(0, 6))
    b = rng.rand(g, 4).astype(np.float32) * 50
    e = rng.rand(g, 10).astype(np.float32) * 50
    k = rng.rand(g, 51).astype(np.float32) * 50
    for fn, a in ((bbox_flip, b), (extreme_flip, e), (keypoint_flip, k)):
        once = fn(a, shape, direction)
        assert once.shape == a.shape and np.allclose(fn(once, shape, direction), a, atol=1e-3)
    polys = [[rng.rand(2 * int(rng.randint(3, 40))) * 50] for _ in range(g)]
    pm = PolygonMasks(polys, shape[0], shape[1])
    for keep in (False, True):
        back = pm.flip(direction, keep).flip(direction, keep)
        assert all(np.allclose(a[0], c[0]) for a, c in zip(pm.masks, back.masks))
    if g:                                                            # a flip reverses the orientation unless the ring is re-ordered
        p = polys[0][0].reshape(-1, 2)
        a0 = contour.signed_area(p)
        assert np.sign(contour.signed_area(pm.flip(direction, False).masks[0][0].reshape(-1, 2))) == -np.sign(a0)
        assert np.sign(contour.signed_area(pm.flip(direction, True).masks[0][0].reshape(-1, 2))) == np.sign(a0)
        assert np.allclose(pm.flip('horizontal', True).masks[0][0][1], p[0, 1])     # ... and keeps the start vertex


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_extreme_points_follow_their_box_through_a_flip(seed):
    """The four extreme points (top, left, bottom, right) of a box stay on the flipped box's matching sides."""
    rng = np.random.RandomState(seed)
    W = 640
    x1, y1 = rng.uniform(0, 300, 2)
    x2, y2 = x1 + rng.uniform(5, 300), y1 + rng.uniform(5, 300)
    u = rng.rand(4)
    e = np.array([[x1 + u[0] * (x2 - x1), y1, x1, y1 + u[1] * (y2 - y1), x1 + u[2] * (x2 - x1), y2, x2, y1 + u[3] * (y2 - y1),
                   (x1 + x2) / 2, (y1 + y2) / 2]], np.float32)
    b = np.array([[x1, y1, x2, y2]], np.float32)
    fe, fb = extreme_flip(e, (480, W, 3), 'horizontal')[0], bbox_flip(b, (480, W, 3), 'horizontal')[0]
    assert abs(fe[1] - fb[1]) < 1e-3 and abs(fe[5] - fb[3]) < 1e-3           # top / bottom points on the top / bottom sides
    assert abs(fe[2] - fb[0]) < 1e-3 and abs(fe[6] - fb[2]) < 1e-3           # the LEFT point lies on the flipped box's left side
    assert abs(fe[8] - (fb[0] + fb[2]) / 2) < 1e-3                           # centre stays the centre


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 5), st.integers(1, 4), st.integers(0, 7))
def test_group_sampler_partitions(seed, spg, world, epoch):
    """Both samplers, any group sizes / batch size / world size: every rank draws the same number of samples, every
    per-GPU batch comes from ONE aspect-ratio group, and together the ranks cover the whole dataset each epoch."""
    import types
    from lsnet_b200.datasets import DistributedGroupSampler, GroupSampler
    rng = np.random.RandomState(seed)
    n = int(rng.randint(1, 60))
    flag = (rng.rand(n) < rng.rand()).astype(np.uint8)
    ds = types.SimpleNamespace(flag=flag)
    per_rank = []
    for rank in range(world):
        s = DistributedGroupSampler(ds, samples_per_gpu=spg, num_replicas=world, rank=rank)
        s.set_epoch(epoch)
        idx = list(s)
        assert len(idx) == len(s) and len(idx) % spg == 0
        for i in range(0, len(idx), spg):
            assert len({int(flag[j]) for j in idx[i:i + spg]}) == 1
        per_rank.append(idx)
    assert len({len(p) for p in per_rank}) == 1
    assert set(j for p in per_rank for j in p) == set(range(n))
    np.random.seed(seed % 1000)
    idx = list(GroupSampler(ds, samples_per_gpu=spg))
    assert set(idx) == set(range(n)) and len(idx) % spg == 0
    for i in range(0, len(idx), spg):
        assert len({int(flag[j]) for j in idx[i:i + spg]}) == 1
