"""CPU tests of the harness-side generator helpers (radix_join_b200.synthetic): the multiset checksum that
pins full-size results, the splitmix64 mirrors and the Zipf sampler.  No GPU, no engine."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from helpers import rj  # noqa: F401  (puts the package on the path)
from radix_join_b200 import synthetic as syn


def _cols(rng, n):
    a = torch.from_numpy(rng.integers(-2**62, 2**62, n))
    b = torch.from_numpy(rng.integers(-5, 5, n))
    ok = torch.from_numpy(rng.random(n) > 0.2)
    return [(a, None), (b, ok)]


def test_checksum_ignores_row_order_but_not_content():
    rng = np.random.default_rng(1)
    cols = _cols(rng, 1000)
    base = syn.checksum_add((0, 0, 0), syn.row_hash_torch(cols))
    perm = torch.from_numpy(rng.permutation(1000))
    shuffled = [(v[perm], None if ok is None else ok[perm]) for v, ok in cols]
    assert syn.checksum_add((0, 0, 0), syn.row_hash_torch(shuffled)) == base
    # accumulating in two halves gives the same sums
    halves = (0, 0, 0)
    for lo, hi in ((0, 400), (400, 1000)):
        halves = syn.checksum_add(halves, syn.row_hash_torch([(v[lo:hi], None if ok is None else ok[lo:hi]) for v, ok in cols]))
    assert halves == base
    # one changed cell, one flipped NULL, swapped columns: all detected
    v2 = cols[0][0].clone()
    v2[17] += 1
    assert syn.checksum_add((0, 0, 0), syn.row_hash_torch([(v2, None), cols[1]])) != base
    ok2 = cols[1][1].clone()
    ok2[3] = ~ok2[3]
    assert syn.checksum_add((0, 0, 0), syn.row_hash_torch([cols[0], (cols[1][0], ok2)])) != base
    swapped = [(cols[1][0], cols[1][1]), (cols[0][0], None)]
    assert syn.checksum_add((0, 0, 0), syn.row_hash_torch(swapped)) != base


def test_null_hashes_apart_from_every_value_of_the_cell():
    v = torch.tensor([0, 0, 7, 7], dtype=torch.int64)
    ok = torch.tensor([True, False, True, False])
    h = syn.row_hash_torch([(v, ok)])
    assert h[0] != h[1] and h[2] != h[3]
    assert h[1] == h[3]  # a NULL does not depend on the bytes underneath


def test_splitmix64_mirrors_agree():
    x = np.arange(-50, 50, dtype=np.int64) * 0x1234567
    got = syn.splitmix64_torch(torch.from_numpy(x)).numpy().view(np.uint64)
    want = syn.splitmix64_numpy(x.view(np.uint64))
    assert np.array_equal(got, want)


def test_zipf_sampler_numpy_and_torch_agree_and_are_skewed():
    z = syn.Zipf(1 << 16, 0.75)
    u = np.random.default_rng(2).random(200_000)
    r_np = z.ranks(u, np)
    r_t = z.ranks(torch.from_numpy(u), torch).numpy()
    assert np.array_equal(r_np, r_t)
    assert r_np.min() == 0 and r_np.max() < (1 << 16)
    counts = np.bincount(r_np, minlength=1 << 16)
    # rank 0 is the most popular and far above the uniform share
    assert counts[0] == counts.max() and counts[0] > 50 * (len(u) / (1 << 16))
