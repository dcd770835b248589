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
    check_partition(counts, world, use_bg)


def check_partition(counts, world, use_bg, chunk=1024):
    p = ShardPlan(counts, use_bg, world, None, 24, 32, chunk)
    counts = np.asarray(counts)
    # chunks tile every image's points exactly once, in order, and never exceed the chunk size
    for i, n in enumerate(counts):
        cs = range(p.icb[i], p.icb[i + 1])
        assert len(cs) >= 1
        assert sum(p.c_cnt[c] for c in cs) == n and all(p.c_img[c] == i for c in cs)
        pos = 0
        for c in cs:
            assert p.c_start[c] == pos and p.c_cnt[c] <= chunk
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
    check_exchange(counts, world)


def check_exchange(counts, world, chunk=1024):
    owners = (np.arange(len(counts)) * 7 + 3) % world     # owners unrelated to where the points fall
    p = ShardPlan(counts, True, world, owners, 24, 32, chunk)
    L, m4 = p.layout, 4 * 24 * 32
    P = _native
    b, c = len(counts), p.total_chunks
    assert 0 == L.flags < L.err < L.push_ticket < L.amax and L.total > L.gfinal > 0 and L.goff > 0 and L.gsorted > L.goff
    recv = np.zeros((world, P.BL_PHASES), dtype=np.uint32)
    for r in range(world):
        sh, sl = p.shards[r], p.slices[r]
        assert sh.push_first[0] == 0 and all(sh.push_first[k] <= sh.push_first[k + 1] for k in range(P.BL_PHASES))
        for ph in range(P.BL_PHASES):
            n = sh.push_first[ph + 1] - sh.push_first[ph]
            assert n == 0 or ph in (P.BL_PH_DENS, P.BL_PH_OUT)        # every other phase travels inside the kernels
            sent = 0
            for k in range(sh.push_first[ph], sh.push_first[ph + 1]):
                src_off, dst_off, packed = (int(v) for v in sl[k])
                nbytes, dst = packed & 0xffffffff, packed >> 32
                assert 0 < nbytes <= 1 << 30 and nbytes % 4 == 0 and 0 <= dst < world
                assert dst_off + nbytes <= (L.total if ph != P.BL_PH_OUT else len(p.owned[r]) * m4)
                if ph == P.BL_PH_DENS and dst != r:
                    sent |= 1 << dst
            if ph == P.BL_PH_DENS:
                assert sent == sh.signal_mask[ph]
            for q in range(world):
                if sh.signal_mask[ph] >> q & 1:
                    assert q != r
                    recv[q, ph] |= np.uint32(1 << r)
        # every image this rank owns is gathered exactly once, into consecutive slots of its gradient tensor
        out = sl[sh.push_first[P.BL_PH_OUT]:sh.push_first[P.BL_PH_OUT + 1]]
        assert sum(int(v[2]) & 0xffffffff for v in out) == len(p.owned[r]) * m4
    for r in range(world):
        for ph in range(P.BL_PHASES):
            assert int(recv[r, ph]) == p.shards[r].wait_mask[ph]      # I wait exactly for those who signal me
    # the destination masks the kernels use agree with the flags
    for r in range(world):
        aux = p.aux[r]
        assert aux.shape == (2 * c + 2 * b,) and aux.dtype == np.uint32
        zmask, gmask, img_mask, owner_mask = aux[:c], aux[c:2 * c], aux[2 * c:2 * c + b], aux[2 * c + b:]
        sh = p.shards[r]
        for ch in range(c):
            g = set(p.groups[p.c_img[ch]])
            if p.c_owner[ch] == r:
                assert zmask[ch] == sum(1 << q for q in g if q != r)
                lead = int(p.lead[p.c_img[ch]])
                assert gmask[ch] == (0 if lead == r else 1 << lead)
            else:
                assert zmask[ch] == 0 and gmask[ch] == 0
        z_dst = int(np.bitwise_or.reduce(zmask)) if c else 0
        assert z_dst == sh.signal_mask[P.BL_PH_Z] and sh.signal_mask[P.BL_PH_MIN] == 0 == sh.wait_mask[P.BL_PH_MIN]
        assert (int(np.bitwise_or.reduce(gmask)) if c else 0) == sh.signal_mask[P.BL_PH_GPART]
        assert int(np.bitwise_or.reduce(img_mask)) == sh.signal_mask[P.BL_PH_CNT]
        assert int(np.bitwise_or.reduce(owner_mask)) == sh.signal_mask[P.BL_PH_GRAD]
        for i in range(b):
            g = p.groups[i]
            assert img_mask[i] == (sum(1 << q for q in g if q != r) if r in g else 0)
            assert owner_mask[i] == ((1 << int(owners[i])) if (p.lead[i] == r and owners[i] != r) else 0)
        leads = sorted(set(int(x) for x in p.lead))
        assert sh.wait_mask[P.BL_PH_LOSS] == sum(1 << q for q in leads if q != r)
    # density reaches every rank that sweeps an image
    for i in range(b):
        for q in p.groups[i]:
            if q != owners[i]:
                assert p.shards[q].wait_mask[P.BL_PH_DENS] >> int(owners[i]) & 1


# ------------------------------------------------------------------------------------------------ row-band sharding
BAND_CASES = [([200, 0, 37], 2, 24, 32), ([200, 0, 37], 3, 24, 32), ([0, 0], 2, 3, 5), ([5], 8, 12, 40),
              (synthetic.config_counts(3), 8, 192, 256), (synthetic.config_counts(3), 4, 192, 256),
              (synthetic.config_counts(2), 5, 128, 96), ([1] * 9, 8, 7, 33)]


@pytest.mark.parametrize("counts,world,hp,wp", BAND_CASES)
def test_band_plan_tiles_the_grid_and_the_exchange_is_consistent(counts, world, hp, wp):
    check_band(counts, world, hp, wp)


def check_band(counts, world, hp, wp):
    from dgvcc_b200.losses.bl_banded import BandPlan, band_chunk_points
    owners = (np.arange(len(counts)) * 7 + 3) % world
    chunk = band_chunk_points(int(np.sum(counts)), world, hp, wp)
    assert 256 <= chunk <= 1024 and chunk % 32 == 0
    p = BandPlan(counts, True, world, owners, hp, wp, chunk)
    L, m4, P, b = p.layout, 4 * hp * wp, _native, len(counts)
    R = L.rows_per_thread
    # the bands are whole rows of pixel tiles, disjoint, in rank order, and cover the grid
    assert p.band_lo[0] == 0 and p.band_hi[-1] == hp
    for r in range(world):
        assert p.band_lo[r] <= p.band_hi[r] and p.band_lo[r] % R == 0 and (p.band_hi[r] % R == 0 or p.band_hi[r] == hp)
        if r:
            assert p.band_lo[r] == p.band_hi[r - 1]
    sizes = (p.band_hi - p.band_lo + R - 1) // R
    assert sizes.max() - sizes.min() <= 1
    assert L.cshare > 0 and L.total >= L.cshare + 4 * world * p.total_rows
    # every chunk is scheduled on every rank
    table = p.meta_for(0)[4 * b + 3:].reshape(-1, 4)
    assert sorted(table[:, 3].tolist()) == list(range(p.total_chunks)) and table[:, 2].max() <= chunk
    got = np.zeros((world, b, hp), dtype=np.int32)      # density rows delivered to each rank
    recv = np.zeros((world, P.BL_PHASES), dtype=np.uint32)
    for r in range(world):
        sh, sl = p.shards[r], p.slices[r]
        assert (sh.band_lo, sh.band_hi) == (p.band_lo[r], p.band_hi[r])
        for ph in range(P.BL_PHASES):
            n = sh.push_first[ph + 1] - sh.push_first[ph]
            assert n == 0 or ph in (P.BL_PH_DENS, P.BL_PH_OUT)
            for k in range(sh.push_first[ph], sh.push_first[ph + 1]):
                src_off, dst_off, packed = (int(v) for v in sl[k])
                nbytes, dst = packed & 0xffffffff, packed >> 32
                assert nbytes > 0 and nbytes % 4 == 0 and 0 <= dst < world
                if ph == P.BL_PH_DENS:
                    assert src_off + nbytes <= len(p.owned[r]) * m4 and L.dens <= dst_off and dst_off + nbytes <= L.dens + b * m4
                    img, rest = divmod(dst_off - L.dens, m4)
                    assert owners[img] == r and src_off % m4 == rest     # the same cells of the same image on both sides
                    cells = np.arange(rest // 4, (rest + nbytes) // 4)
                    np.add.at(got[dst, img], np.unique(cells // wp), 1)
                else:
                    assert dst == r and dst_off + nbytes <= len(p.owned[r]) * m4
            for q in range(world):
                if sh.signal_mask[ph] >> q & 1:
                    assert q != r
                    recv[q, ph] |= np.uint32(1 << r)
        out = sl[sh.push_first[P.BL_PH_OUT]:sh.push_first[P.BL_PH_OUT + 1]]
        assert sum(int(v[2]) & 0xffffffff for v in out) == len(p.owned[r]) * m4
        om = p.aux[r]
        assert om.shape == (b,) and all(om[i] == (0 if owners[i] == r else 1 << int(owners[i])) for i in range(b))
    for r in range(world):
        for ph in range(P.BL_PHASES):
            assert int(recv[r, ph]) == p.shards[r].wait_mask[ph]
        others = sum(1 << q for q in range(world) if q != r)
        assert p.shards[r].signal_mask[P.BL_PH_CNT] == others == p.shards[r].wait_mask[P.BL_PH_CNT]
        for ph in (P.BL_PH_MIN, P.BL_PH_Z, P.BL_PH_LOSS, P.BL_PH_GPART):
            assert p.shards[r].signal_mask[ph] == 0 == p.shards[r].wait_mask[ph]
        # a rank receives exactly the rows of its band, of every image, at least once (slices may split a row)
        for i in range(b):
            rows = np.nonzero(got[r, i])[0]
            assert rows.tolist() == list(range(p.band_lo[r], p.band_hi[r]))


# ------------------------------------------------------------------------------------------------ randomised plans
def _random_counts(rng):
    b = int(rng.integers(1, 13))
    kind = rng.integers(0, 4)
    if kind == 0:       # sparse scenes incl. empty images
        return [int(v) for v in rng.integers(0, 40, size=b)]
    if kind == 1:       # one crowd among small images (the shape that breaks whole-image partitioning)
        c = [int(v) for v in rng.integers(0, 300, size=b)]
        c[int(rng.integers(0, b))] = int(rng.integers(3000, 13000))
        return c
    if kind == 2:       # sizes around the chunk boundaries
        return [int(rng.choice([0, 1, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2048])) for _ in range(b)]
    return [int(v) for v in rng.integers(0, 5000, size=b)]


@pytest.mark.parametrize("seed", range(6))
def test_random_plans_keep_every_invariant(seed):
    """The same invariants as above on random head counts, world sizes 1..8, chunk sizes and grids (25 plans per seed)."""
    rng = np.random.default_rng(1000 + seed)
    for _ in range(25):
        counts, world = _random_counts(rng), int(rng.integers(1, 9))
        chunk = int(rng.choice([128, 256, 352, 512, 1024]))
        check_partition(counts, world, bool(rng.integers(0, 2)), chunk)
        check_exchange(counts, world, chunk)
        hp, wp = int(rng.integers(1, 200)), int(rng.integers(1, 260))
        check_band(counts, world, hp, wp)
