"""Host-side planning of the point-chunk sharding (dgvcc_b200/losses/bl_sharded.py): no GPU needed."""
import numpy as np
import pytest

from dgvcc_b200 import _native, synthetic
from dgvcc_b200.losses.bl_sharded import ShardPlan

CASES = [
    ([200, 0, 37], 2), ([200, 0, 37], 3), ([0, 0, 0], 2), ([5], 4), ([0, 1, 0, 3000, 2, 0], 4),
    (synthetic.config_counts(2), 4), (synthetic.config_counts(3), 8), (synthetic.config_counts(3), 2),
    ([1, 1, 1, 1, 1, 1, 1, 1, 1], 8),
]


@pytest.mark.parametrize("counts,world", CASES)
@pytest.mark.parametrize("use_bg", [True, False])
def test_partition_covers_every_point_once_and_balances(counts, world, use_bg):
    p = ShardPlan(counts, use_bg, world, None, 24, 32, 1024)
    counts = np.asarray(counts)
    # chunks tile every image's points exactly once, in order, and never exceed the chunk size
    for i, n in enumerate(counts):
        cs = range(p.icb[i], p.icb[i + 1])
        assert len(cs) >= 1
        assert sum(p.c_cnt[c] for c in cs) == n and all(p.c_img[c] == i for c in cs)
        pos = 0
        for c in cs:
            assert p.c_start[c] == pos and p.c_cnt[c] <= 1024
            pos += p.c_cnt[c]
    # a rank's chunks are contiguous and its points are the span [bounds[r], bounds[r+1])
    per_rank = np.zeros(world, dtype=np.int64)
    for c in range(p.total_chunks):
        r = p.c_owner[c]
        assert p.chunk_lo[r] <= c < p.chunk_hi[r]
        g0 = p.pt_off[p.c_img[c]] + p.c_start[c]
        if p.c_cnt[c]:
            assert p.bounds[r] <= g0 and g0 + p.c_cnt[c] <= p.bounds[r + 1]
        per_rank[r] += p.c_cnt[c]
    assert per_rank.sum() == counts.sum()
    assert per_rank.max() - per_rank.min() <= 1                       # equal spans, whatever the images look like
    for r in range(world):
        sh = p.shards[r]
        assert (sh.pt_lo, sh.pt_hi) == (p.bounds[r], p.bounds[r + 1])
        if sh.chunk_hi > sh.chunk_lo:
            imgs = p.c_img[sh.chunk_lo:sh.chunk_hi]
            assert sh.img_lo == imgs.min() and sh.img_hi == imgs.max() + 1
    # the meta table of a rank schedules exactly its own chunks
    b = len(counts)
    for r in range(world):
        meta = p.meta_for(r)
        table = meta[4 * b + 3:].reshape(-1, 4)
        n = p.chunk_hi[r] - p.chunk_lo[r]
        assert sorted(table[:n, 3].tolist()) == list(range(p.chunk_lo[r], p.chunk_hi[r]))
    assert sorted(p.meta_all()[4 * b + 3:].reshape(-1, 4)[:, 3].tolist()) == list(range(p.total_chunks))


@pytest.mark.parametrize("counts,world", CASES)
def test_exchange_plan_is_consistent(counts, world):
    owners = (np.arange(len(counts)) * 7 + 3) % world     # owners unrelated to where the points fall
    p = ShardPlan(counts, True, world, owners, 24, 32, 1024)
    L, m4 = p.layout, 4 * 24 * 32
    assert 0 == L.flags < L.err < L.push_ticket < L.amax and L.gpart != L.minpart and L.total > L.gfinal > 0
    recv = np.zeros((world, _native.BL_PHASES), dtype=np.uint32)
    for r in range(world):
        sh, sl = p.shards[r], p.slices[r]
        assert sh.push_first[0] == 0 and all(sh.push_first[k] <= sh.push_first[k + 1] for k in range(_native.BL_PHASES))
        for ph in range(_native.BL_PHASES):
            sent = 0
            for k in range(sh.push_first[ph], sh.push_first[ph + 1]):
                src_off, dst_off, packed = (int(v) for v in sl[k])
                nbytes, dst = packed & 0xffffffff, packed >> 32
                assert 0 < nbytes <= 1 << 30 and nbytes % 4 == 0 and 0 <= dst < world
                assert dst_off + nbytes <= (L.total if ph != _native.BL_PH_OUT else len(p.owned[r]) * m4)
                if ph not in (_native.BL_PH_OUT,) and dst != r:
                    sent |= 1 << dst
                    recv[dst, ph] |= np.uint32(1 << r)
            assert sent == sh.signal_mask[ph]
        # every image this rank owns is gathered exactly once, into consecutive slots of its gradient tensor
        out = sl[sh.push_first[_native.BL_PH_OUT]:sh.push_first[_native.BL_PH_OUT + 1]]
        assert sum(int(v[2]) & 0xffffffff for v in out) == len(p.owned[r]) * m4
    for r in range(world):
        for ph in range(_native.BL_PHASES):
            assert int(recv[r, ph]) == p.shards[r].wait_mask[ph]      # I wait exactly for those who send to me
    # density reaches every rank that sweeps an image; gradient sums reach the lead; the finished gradient its owner
    for i in range(len(counts)):
        for q in p.groups[i]:
            if q != owners[i]:
                assert p.shards[q].wait_mask[_native.BL_PH_DENS] >> int(owners[i]) & 1
        if owners[i] != p.lead[i]:
            assert p.shards[int(owners[i])].wait_mask[_native.BL_PH_GRAD] >> int(p.lead[i]) & 1
