"""The exchange layout of the fused multi-GPU join (radix_join_b200.dist_join.exchange_layout): every rank derives,
from the all-gathered histograms alone, where its runs start in every owner's receive arrays.  Simulated here for
G ranks in one process with numpy standing in for the scatter kernel: the runs must tile every owner's arrays
exactly, digit by digit in ascending order, so that the owner's local plan (prefix sums of `local_hist`) finds
each pass-1 region where the senders put it."""
import numpy as np
import pytest
import torch

import helpers as H  # noqa: F401  (sys.path)
from radix_join_b200 import dist_join as dj


def fmix32(h):
    h = h.astype(np.uint64)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return h.astype(np.uint32)


@pytest.mark.parametrize("world,bits", [(2, 15), (4, 12), (8, 15), (8, 9), (4, 6), (2, 1)])
def test_runs_tile_the_owners_arrays(world, bits):
    rng = np.random.default_rng(world * 100 + bits)
    g = dj.log2_exact(world)
    two = bits > dj.MAX_PASS_BITS
    p1 = (bits + 1) // 2 if two else bits
    assert p1 >= g
    keys = [[rng.integers(0, 1 << 20, int(rng.integers(2000, 9000))).astype(np.uint32) for _ in range(2)] for _ in range(world)]
    part = [[fmix32(k) & np.uint32((1 << bits) - 1) for k in sides] for sides in keys]
    Hist = torch.zeros(world, 2, 1 << bits, dtype=torch.int64)
    for s in range(world):
        for side in range(2):
            Hist[s, side] = torch.from_numpy(np.bincount(part[s][side], minlength=1 << bits))
    owned_arrays = [[None, None] for _ in range(world)]
    layouts = [dj.exchange_layout(Hist, me, g, bits, p1) for me in range(world)]
    for o in range(world):
        for side in range(2):
            owned_arrays[o][side] = np.full(int(layouts[o][2][side]), -1, dtype=np.int64)  # holds the final partition of each tuple
    # the scatter, emulated: rank s writes its tuples of digit d at cursor[d]...
    for s in range(world):
        cursor = layouts[s][0].numpy()
        for side in range(2):
            digit = part[s][side] >> np.uint32(bits - p1)
            for d in range(1 << p1):
                sel = part[s][side][digit == d]
                o = d >> (p1 - g)
                at = int(cursor[side, d])
                dst = owned_arrays[o][side]
                assert (dst[at: at + len(sel)] == -1).all(), "runs overlap"
                dst[at: at + len(sel)] = sel
    for o in range(world):
        _cur, local_hist, owned, per_owner, _sent = layouts[o]
        assert torch.equal(per_owner, torch.stack([lay[2] for lay in layouts]).T)  # everybody agrees on what o receives
        lo = o * ((1 << bits) >> g)
        for side in range(2):
            arr = owned_arrays[o][side]
            assert (arr >= 0).all(), "holes in the receive array"
            # grouped by pass-1 digit, ascending: region boundaries = prefix sums of the local histogram
            lh = local_hist[side].numpy()
            assert lh.sum() == len(arr) == int(owned[side])
            digit = arr >> (bits - p1)
            assert (np.diff(digit) >= 0).all()
            per_region = lh.reshape(-1, 1 << (bits - p1)).sum(-1)
            starts = np.concatenate([[0], np.cumsum(per_region)])
            for j in range(len(per_region)):
                seg = arr[starts[j]: starts[j + 1]]
                assert ((seg >> (bits - p1)) == (lo >> (bits - p1)) + j).all()
            # and the final partitions it owns are exactly its range
            assert ((arr >= lo) & (arr < lo + ((1 << bits) >> g))).all()
            assert np.array_equal(np.bincount(arr - lo, minlength=len(lh)), lh)


def test_choose_bits_follows_the_engine():
    assert dj.choose_bits(1 << 26) == 15 and dj.choose_bits((1 << 26) + 1) == 15
    assert dj.choose_bits(2048) == 0 and dj.choose_bits(2049) == 1
    assert dj.choose_bits(1 << 19) == 8 and dj.choose_bits((1 << 19) + 256) == 9


@pytest.mark.parametrize("world,bits", [(2, 15), (4, 12), (8, 15), (8, 9)])
def test_pull_layout_addresses_every_tuple_once(world, bits):
    """pull variant: pass 1 stays local, the owner's second pass reads each region's runs from the senders' arrays
    through a table of biased addresses -- simulated with address = rank * BIG + element index (widths of 1)"""
    rng = np.random.default_rng(world * 10 + bits)
    g = dj.log2_exact(world)
    p1 = (bits + 1) // 2
    BIG = 10**9
    keys = [rng.integers(0, 1 << 20, int(rng.integers(3000, 20000))).astype(np.uint32) for _ in range(world)]
    part = [fmix32(k) & np.uint32((1 << bits) - 1) for k in keys]
    Hist = torch.zeros(world, 2, 1 << bits, dtype=torch.int64)
    for s in range(world):
        Hist[s, 1] = torch.from_numpy(np.bincount(part[s], minlength=1 << bits))
    ptrs = [torch.arange(world, dtype=torch.int64) * BIG for _ in range(5)]
    lays = [dj.pull_layout(Hist, me, g, bits, p1, 1, ptrs, [1] * 5) for me in range(world)]
    # pass 1, emulated: every rank groups its tuples by digit at its own cursor
    arrays = []
    for s in range(world):
        cur = lays[s][0].numpy().copy()
        out = np.full(len(keys[s]), -1, dtype=np.int64)
        digit = part[s] >> np.uint32(bits - p1)
        for d in range(1 << p1):
            sel = part[s][digit == d]
            out[cur[d]: cur[d] + len(sel)] = sel
        assert (out >= 0).all()
        arrays.append(out)
    seen = [np.zeros(len(a), dtype=np.int64) for a in arrays]
    per = (1 << p1) >> g
    for me in range(world):
        _cur, table, start, tile, group = [x.numpy() for x in lays[me]]
        n_sub = per * world
        assert table.shape == (n_sub, 5) and len(start) == len(tile) == n_sub + 1 and len(group) == n_sub
        for x in range(n_sub):
            cnt = start[x + 1] - start[x]
            assert tile[x + 1] - tile[x] == (cnt + dj.SCATTER_TILE - 1) // dj.SCATTER_TILE
            assert group[x] == x // world
            for i in (range(cnt) if cnt < 50 else [0, cnt // 2, cnt - 1]):
                addr = table[x, 0] + start[x] + i      # element index = the virtual position
                q, idx = divmod(int(addr), BIG)
                assert q == x % world
                f = arrays[q][idx]
                assert (f >> (bits - p1)) == me * per + group[x]  # the right pass-1 digit, owned by `me`
            if cnt:
                q, idx0 = divmod(int(table[x, 0] + start[x]), BIG)
                seen[q][idx0: idx0 + cnt] += 1
    for s in range(world):
        assert (seen[s] == 1).all(), "every tuple is read by exactly one owner"


@pytest.mark.gpu
@pytest.mark.parametrize("world,bits", [(2, 15), (8, 15), (4, 10)])
def test_layout_kernel_matches_the_tensor_version(world, bits):
    """rj_dist_layout (one kernel) against pull_layout / exchange_layout (tensor operations) on random histograms"""
    import radix_join_b200 as rj
    ctx = rj.build_context(0)
    try:
        ops = dj.CudaOps(ctx)
        g = dj.log2_exact(world)
        p1 = (bits + 1) // 2
        rng = np.random.default_rng(bits)
        Hn = rng.integers(0, 3000, (world, 2, 1 << bits)).astype(np.int32)
        Hn[:, :, rng.integers(0, 1 << bits, 500)] = 0
        H = torch.from_numpy(Hn).cuda()
        widths = [[4, 8, 0, 1, 1], [4, 8, 4, 1, 1]]
        ptrs = rng.integers(1 << 30, 1 << 40, (2, 5, 8)).astype(np.int64) // 16 * 16
        for me in range(world):
            lay = ops.dist_layout(H.reshape(-1), me, g, bits, p1, torch.from_numpy(ptrs).cuda(), torch.tensor(widths, dtype=torch.int32, device="cuda"))
            torch.cuda.synchronize()
            ndig, nloc = 1 << p1, (1 << bits) >> g
            H64 = H.to(torch.int64)
            _c, local_hist, owned, _po, sent = dj.exchange_layout(H64, me, g, bits, p1)
            assert torch.equal(lay["local_hist"].view(2, nloc).to(torch.int64), local_hist)
            assert lay["scalars"].tolist() == [int(owned[0]), int(sent[0]), int(owned[1]), int(sent[1])]
            for side in range(2):
                tp = [torch.from_numpy(ptrs[side, a, :world].copy()).cuda() for a in range(5)]
                cur, table, start, tile, group = dj.pull_layout(H64, me, g, bits, p1, side, tp, widths[side])
                assert torch.equal(lay["cursor"].view(2, ndig)[side].to(torch.int64), cur)
                assert torch.equal(lay["start"].view(2, ndig + 1)[side].to(torch.int64), start)
                assert torch.equal(lay["tile"].view(2, ndig + 1)[side].to(torch.int64), tile)
                assert torch.equal(lay["group"].view(2, ndig)[side].to(torch.int64), group)
                assert torch.equal(lay["table"].view(2, ndig, 5)[side], table)
        ops.close()
    finally:
        rj.destroy_context(ctx)
