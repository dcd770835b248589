"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the authoring container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [bl|dmap|isw ...]

Needs /root/reference (read-only mount); the GPU box has no such path, which is
why the outputs are committed.  Every array in a fixture is either an input
(prefix ``in_``) or an output of the reference's own code (prefix ``ref_``).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from dgvcc_b200 import synthetic  # noqa: E402


def load_ref(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# --------------------------------------------------------------------------- BL
def bl_case(name, counts, width, height, stride, sigma, use_bg=True, bg_ratio=1.0, seed_cfg=7,
            jitter_outside=False, rows=(0, 1, -1)):
    """Run reference BL on a (possibly rectangular) grid via the square + zero-pad construction."""
    bl = load_ref("losses/bl.py", "ref_bl")
    # outside="clip": the generator as it was when these fixtures were drawn (synthetic.crowd_points)
    pts, tgt, dens, st = synthetic.bl_batch(seed_cfg, counts, width, height, stride, outside="clip")
    if jitter_outside:  # bay_dataset.py:85-98 keeps heads whose box overlaps the crop >= 30 %
        for p in pts:
            if len(p):
                p[::7] += np.float32(9.5)
                p[::11] -= np.float32(6.25)
    c_size = max(width, height)
    g = c_size // stride
    hp, wp = height // stride, width // stride
    B = len(counts)
    sq = np.zeros((B, 1, g, g), dtype=np.float32)
    sq[:, :, :hp, :wp] = dens
    mod = bl.BL(sigma, c_size, stride, bg_ratio, use_bg, "cpu")
    d = torch.from_numpy(sq).requires_grad_(True)
    tp = [torch.from_numpy(p.copy()) for p in pts]
    tt = [torch.from_numpy(t.copy()) for t in tgt]
    ts = torch.from_numpy(st.copy())
    loss = mod(tp, ts, tt, d)
    loss.backward()
    probs = mod.post_prob([torch.from_numpy(p.copy()) for p in pts], ts)
    out = {
        "in_width": width, "in_height": height, "in_stride": stride, "in_sigma": sigma,
        "in_use_bg": int(use_bg), "in_bg_ratio": bg_ratio, "in_st_sizes": st,
        "in_density": dens, "in_counts": np.asarray(counts, dtype=np.int64),
        "ref_loss": loss.detach().numpy(),
        "ref_grad": d.grad.numpy()[:, :, :hp, :wp].copy(),
    }
    for i in range(B):
        out[f"in_points_{i}"] = pts[i]
        out[f"in_targets_{i}"] = tgt[i]
        if probs[i] is None:
            continue
        p3 = probs[i].view(probs[i].shape[0], g, g)[:, :hp, :wp]
        out[f"ref_count_{i}"] = (torch.from_numpy(dens[i]) * p3).reshape(p3.shape[0], -1).sum(1).numpy()
        out[f"ref_colsum_{i}"] = p3.sum(0).numpy()
        sel = sorted({r % p3.shape[0] for r in rows})
        out[f"ref_prob_rows_{i}"] = np.asarray(sel, dtype=np.int64)
        out[f"ref_prob_{i}"] = p3[sel].numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"bl_{name}.npz"), **out)
    print("bl", name, "loss", float(loss))


def make_bl():
    bl_case("c1", [200], 1024, 768, 8, 8.0, seed_cfg=1)                       # BASELINE config 1
    bl_case("mixed", [150, 0, 3, 1, 40], 512, 512, 8, 8.0)                    # empty / tiny images
    bl_case("nobg", [60, 1, 2, 0], 256, 384, 8, 8.0, use_bg=False)            # use_background=False
    bl_case("empty", [0, 0], 256, 256, 8, 8.0)                                # all-empty batch
    bl_case("sigma10", [90, 17], 320, 320, 4, 10.0, bg_ratio=0.15)            # 2*sigma^2 not a power of 2
    bl_case("outside", [120], 384, 256, 8, 8.0, jitter_outside=True)          # heads outside the crop


# ------------------------------------------------------------------------- dmap
def dmap_points(seed, n, w, h, dtype):
    rng = np.random.default_rng(seed)
    return synthetic.crowd_points(rng, n, w, h, dtype=dtype, outside="clip")   # as when the fixtures were drawn


def make_dmap():
    dg = load_ref("utils/dmap_gen.py", "ref_dmap_gen")
    out = {}
    cases = [  # name, (H, W), N, dtype, tweak
        ("a40", (96, 128), 40, np.float64, None),
        ("a4", (64, 80), 4, np.float64, None),        # smallest N that uses the kNN sigma
        ("a3", (64, 80), 3, np.float64, None),        # N <= 3 -> sigma 15
        ("a1", (50, 70), 1, np.float32, None),
        ("a0", (40, 40), 0, np.float64, None),
        ("a25f32", (90, 110), 25, np.float32, None),  # QNRF path yields float32 points
        ("oob", (80, 100), 30, np.float64, "oob"),    # heads right/below the image are skipped but stay neighbours
        ("dup", (72, 72), 12, np.float64, "dup"),     # duplicated heads: sigma can be 0 -> identity filter
    ]
    for idx, (name, shape, n, dtype, tweak) in enumerate(cases):
        pts = dmap_points(4000 + idx, n, shape[1], shape[0], dtype)
        if tweak == "oob":
            pts[::5, 0] += shape[1] * 0.3
            pts[1::7, 1] += shape[0] * 0.25
        if tweak == "dup":
            pts[4:8] = pts[0]
        img = np.zeros(shape + (3,), dtype=np.uint8)
        out[f"{name}_shape"] = np.asarray(shape)
        out[f"{name}_points"] = pts
        out[f"{name}_adaptive"] = dg.gaussian_filter_density(img, pts)
        out[f"{name}_fixed"] = dg.gaussian_filter_density_fixed(img, pts)
        print("dmap", name, out[f"{name}_adaptive"].sum(), out[f"{name}_fixed"].sum())
    # kNN bookkeeping exactly as dmap_gen.py:34-36 builds it
    from scipy.spatial import KDTree
    for name, n, dtype in (("knn2000", 2000, np.float64), ("knn700f32", 700, np.float32)):
        pts = dmap_points(4100 + n, n, 1500, 900, dtype)
        d, loc = KDTree(pts.copy(), leafsize=2048).query(pts, k=4)
        out[f"{name}_points"], out[f"{name}_dist"], out[f"{name}_loc"] = pts, d, loc
    np.savez_compressed(os.path.join(HERE, "dmap_cases.npz"), **out)


# -------------------------------------------------------------------------- isw
def make_isw():
    iw = load_ref("models/ISW/instance_whitening.py", "ref_instance_whitening")
    sys.path.insert(0, ROOT)
    from oracle import isw_oracle
    out = {}
    cases = [  # name, (B, C, H, W), mask keep fraction, margin
        ("c64", (2, 64, 24, 20), 0.5, 0.0),
        ("c128", (3, 128, 10, 12), 1.0, 0.0),
        ("c48", (2, 48, 9, 7), 0.5, 0.0),          # C not a multiple of 64, odd HW
        ("margin", (4, 64, 8, 8), 0.5, 0.02),      # IRW-style positive margin: some samples clamp to 0
    ]
    for idx, (name, shape, frac, margin) in enumerate(cases):
        g = torch.Generator().manual_seed(7000 + idx)
        x = torch.randn(shape, generator=g) * 1.7 + 0.3
        c = shape[1]
        eye = torch.eye(c)
        mask = isw_oracle.upper_mask(c, frac, 7100 + idx)
        num_remove = mask.sum()
        xin = x.clone().requires_grad_(True)
        y, w = iw.InstanceWhitening(c)(xin)
        cov, _ = iw.get_covariance_matrix(w, eye=eye)
        loss = iw.instance_whitening_loss(w, eye, mask, margin, num_remove)
        loss.backward()
        wleaf = w.detach().clone().requires_grad_(True)
        iw.instance_whitening_loss(wleaf, eye, mask, margin, num_remove).backward()
        out.update({f"{name}_x": x.numpy(), f"{name}_mask": mask.numpy(), f"{name}_margin": np.float32(margin),
                    f"{name}_norm": y.detach().numpy(), f"{name}_cov": cov.detach().numpy(),
                    f"{name}_loss": loss.detach().numpy(), f"{name}_grad_x": xin.grad.numpy(),
                    f"{name}_grad_w": wleaf.grad.numpy()})
        print("isw", name, float(loss))
    np.savez_compressed(os.path.join(HERE, "isw_cases.npz"), **out)


# -------------------------------------------------------------------------- bay
def make_bay():
    """BayesianDataset._cal_dists and the crop block of _train_transform, run on the unmodified class."""
    import importlib
    import random
    import types
    from PIL import Image
    pkg = types.ModuleType("datasets")        # the HuggingFace `datasets` package shadows the reference's
    pkg.__path__ = [os.path.join(REF, "datasets")]
    saved = sys.modules.get("datasets")
    sys.modules["datasets"] = pkg
    sys.path.insert(0, REF)
    try:
        bd = importlib.import_module("datasets.bay_dataset")
    finally:
        sys.path.remove(REF)
    out = {}
    rng = np.random.default_rng(8000)
    for name, n, dtype in (("n0", 0, np.float64), ("n1", 1, np.float64), ("n2", 2, np.float64), ("n3", 3, np.float64),
                           ("n4", 4, np.float64), ("n300", 300, np.float64), ("n200f32", 200, np.float32)):
        pts = synthetic.crowd_points(rng, n, 900, 700, dtype=dtype)
        out[f"dist_{name}_pts"] = pts
        out[f"dist_{name}_ref"] = bd.BayesianDataset._cal_dists(None, pts)
    # _train_transform on a fake instance: record the random crop and resize so the crop block can be replayed
    rec = {}
    real_crop = bd.random_crop
    def recording_crop(im_h, im_w, ch, cw):
        rec["ij"] = real_crop(im_h, im_w, ch, cw)
        return rec["ij"]
    bd.random_crop = recording_crop
    ds = bd.BayesianDataset.__new__(bd.BayesianDataset)
    ds.pre_resize, ds.crop_size, ds.transform = 1, (256, 256), (lambda im: torch.zeros(1))
    for k, (w, h, n, dtype, seed) in enumerate([(640, 480, 150, np.float64, 1), (300, 220, 60, np.float64, 2),
                                                (700, 500, 400, np.float32, 3), (500, 400, 0, np.float64, 4)]):
        gt = synthetic.crowd_points(np.random.default_rng(8100 + k), n, w, h, dtype=dtype)
        dists = bd.BayesianDataset._cal_dists(None, gt)
        random.seed(seed)
        state = random.getstate()
        img = Image.fromarray(np.zeros((h, w, 3), dtype=np.uint8))
        _, gt_out, targ, st_size = ds._train_transform(img, gt.copy(), dists)
        # replay the draws to recover the geometry the crop block saw
        random.setstate(state)
        random.random()                                   # grey-scale draw
        factor = 1 * random.random() * 0.8 + 0.6
        nw, nh = int(w * factor), int(h * factor)
        g = gt.copy()
        cw, chh = w, h
        if min(nw, nh) >= 256:
            cw, chh = nw, nh
            g = g * factor
        if min(cw, chh) < 256:
            from utils.misc import get_padding
            (left, top, _, _), chh, cw = get_padding(chh, cw, 256, 256)
            if len(g) > 0:
                g = g + [left, top]
        i, j = rec["ij"]
        out[f"crop_{k}_gt"], out[f"crop_{k}_dists"] = g, dists
        out[f"crop_{k}_ijhw"] = np.asarray([i, j, 256, 256])
        out[f"crop_{k}_ref_gt"], out[f"crop_{k}_ref_targ"] = gt_out.numpy(), targ.numpy()
        print("bay crop", k, "kept", len(targ), "of", n)
    if saved is not None:
        sys.modules["datasets"] = saved
    np.savez_compressed(os.path.join(HERE, "bay_cases.npz"), **out)


# -------------------------------------------------------------------------- den
def make_den():
    """Density side of DenClsDataset.__getitem__ / _train_transform (pad, crop, sum-pool, flip, 16x16 block
    occupancy), run on the unmodified class with its three file loaders stubbed."""
    import importlib
    import random
    import types
    from PIL import Image
    pkg = types.ModuleType("datasets")        # the HuggingFace `datasets` package shadows the reference's
    pkg.__path__ = [os.path.join(REF, "datasets")]
    saved = sys.modules.get("datasets")
    sys.modules["datasets"] = pkg
    sys.path.insert(0, REF)
    try:
        dc = importlib.import_module("datasets.den_cls_dataset")
        from utils.misc import get_padding
    finally:
        sys.path.remove(REF)
    rec = {}
    real_crop = dc.random_crop
    def recording_crop(im_h, im_w, ch, cw):
        rec["ij"] = real_crop(im_h, im_w, ch, cw)
        return rec["ij"]
    dc.random_crop = recording_crop
    out = {}
    cases = [  # (w, h, crop, downsample, heads, seed)
        (640, 480, (256, 256), 1, 300, 1), (640, 480, (256, 256), 8, 300, 2), (300, 200, (256, 256), 4, 80, 3),
        (900, 700, (512, 256), 2, 40, 4), (1100, 800, (320, 320), 4, 2000, 5), (400, 400, (256, 256), 8, 0, 6),
    ]
    for k, (w, h, crop, down, n, seed) in enumerate(cases):
        rng = np.random.default_rng(8300 + k)
        pts = synthetic.crowd_points(rng, n, w, h, dtype=np.float64)
        dmap = np.zeros((h, w), dtype=np.float32)
        for x, y in pts:  # small positive stamps: what a *_dmap.npy holds (non-negative, mostly zero)
            x0, y0 = int(x), int(y)
            dmap[max(0, y0 - 2):y0 + 3, max(0, x0 - 2):x0 + 3] += rng.random(dmap[max(0, y0 - 2):y0 + 3, max(0, x0 - 2):x0 + 3].shape, dtype=np.float32) / 25
        ds = dc.DenClsDataset.__new__(dc.DenClsDataset)
        ds.img_fns, ds.root, ds.method, ds.gt_dir = ["/data/train/img_1.jpg"], "/data", "train", None
        ds.crop_size, ds.downsample, ds.pre_resize = crop, down, 1
        ds.transform = ds.more_transform = (lambda im: torch.zeros(1))
        img = Image.fromarray(np.zeros((h, w, 3), dtype=np.uint8))
        ds._load_img = lambda fn, img=img: (img, ".jpg")
        ds._load_gt = lambda fn, pts=pts: pts.copy()
        ds._load_dmap = lambda fn, dmap=dmap: dmap.copy()
        random.seed(seed)
        state = random.getstate()
        _, _, _, d_out, b_out = ds[0]
        # replay the draws: grey-scale, crop (recorded), flip
        random.setstate(state)
        random.random()
        ph, pw = h, w
        left = top = 0
        if 1.0 * min(w, h) < min(crop):
            (left, top, _, _), ph, pw = get_padding(h, w, crop[0], crop[1])
        i, j = dc.random_crop.__wrapped__(ph, pw, crop[0], crop[1]) if hasattr(dc.random_crop, "__wrapped__") else real_crop(ph, pw, crop[0], crop[1])
        assert (i, j) == rec["ij"]
        flip = int(random.random() > 0.5)
        out[f"den_{k}_dmap"] = dmap
        out[f"den_{k}_geom"] = np.asarray([left, top, i, j, crop[0], crop[1], down, flip])
        out[f"den_{k}_ref_dmap"], out[f"den_{k}_ref_bmap"] = d_out.numpy(), b_out.numpy()
        print("den", k, "geom", out[f"den_{k}_geom"].tolist(), "sum", float(d_out.sum()), "occupied", int(b_out.sum()), "of", b_out.numel())
    if saved is not None:
        sys.modules["datasets"] = saved
    np.savez_compressed(os.path.join(HERE, "den_cases.npz"), **out)


# -------------------------------------------------------------------------- cov
def make_cov():
    """CovMatrix_ISW (models/ISW/cov_settings.py:16-89), unmodified, on CPU: `.cuda()` made a no-op, the absent
    `kmeans1d` stubbed (only the relax_denom == 0 branch uses it; the shipped default is 2.0,
    models/ISW/__init__.py:23)."""
    import importlib
    import types
    pkg = types.ModuleType("refisw")
    pkg.__path__ = [os.path.join(REF, "models", "ISW")]
    sys.modules["refisw"] = pkg
    sys.modules.setdefault("kmeans1d", types.ModuleType("kmeans1d"))
    saved_cuda, saved_dev = torch.Tensor.cuda, torch.cuda.current_device
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.current_device = lambda: 0
    try:
        cs = importlib.import_module("refisw.cov_settings")
        out = {}
        for name, dim, relax, rounds, batches in (("c16", 16, 2.0, 1, 3), ("c64", 64, 2.0, 2, 4), ("c64r3", 64, 3.0, 2, 2),
                                                  ("c256", 256, 2.0, 1, 2)):
            g = torch.Generator().manual_seed(8400 + dim + int(relax))
            cm = cs.CovMatrix_ISW(dim=dim, relax_denom=relax)
            out[f"{name}_cfg"] = np.asarray([dim, relax, rounds, batches])
            for r in range(rounds):
                for b in range(batches):
                    var = (torch.rand(dim, dim, generator=g) ** 3).triu(1)  # what var(f_cor * reverse_eye, dim=0) looks like
                    out[f"{name}_var_{r}_{b}"] = var.numpy()
                    cm.set_variance_of_covariance(var)
                cm.set_mask_matrix()   # the second round ANDs with the first mask (cov_settings.py:70-71)
                eye, mask, margin, num = cm.get_mask_matrix()
                out[f"{name}_mask_{r}"] = mask.numpy().copy()
                out[f"{name}_num_{r}"] = np.asarray([float(num), float(margin), float(cm.margin), float(cm.num_off_diagonal)])
            cm.reset_mask_matrix()
            assert cm.mask_matrix is None
    finally:
        torch.Tensor.cuda, torch.cuda.current_device = saved_cuda, saved_dev
    np.savez_compressed(os.path.join(HERE, "cov_cases.npz"), **out)


# -------------------------------------------------------------------------- aux (lw / ortho losses)
def make_aux():
    """losses/lw.py:5-18 and losses/ortho.py:5-11, the unmodified files loaded by path: values and gradients."""
    lw = load_ref("losses/lw.py", "ref_lw").lw_loss
    ortho = load_ref("losses/ortho.py", "ref_ortho").ortho_loss
    out = {}
    g = torch.Generator().manual_seed(8500)
    for name, (n, c, h, w), masked in (("lw_a", (2, 32, 12, 10), 0), ("lw_b", (3, 64, 16, 16), 1), ("lw_c", (2, 128, 16, 16), 2),
                                        ("lw_d", (1, 48, 9, 7), 0)):
        x = (torch.randn(n, c, h, w, generator=g) * (1 + torch.rand(1, c, 1, 1, generator=g)) + torch.randn(1, c, 1, 1, generator=g))
        x.requires_grad_(True)
        mask = None
        if masked == 1:
            mask = (torch.rand(n, 1, h, w, generator=g) < 0.6).float()
        elif masked == 2:
            mask = torch.rand(n, 1, h, w, generator=g)
        loss = lw(x, mask)
        loss.backward()
        out[f"{name}_x"], out[f"{name}_loss"], out[f"{name}_grad"] = x.detach().numpy(), loss.detach().numpy(), x.grad.numpy()
        if mask is not None:
            out[f"{name}_mask"] = mask.numpy()
    for name, (c, p_) in (("or_a", (32, 200)), ("or_b", (64, 500)), ("or_c", (100, 333)), ("or_d", (128, 512))):
        x = torch.randn(c, p_, generator=g, requires_grad=True)
        y = (0.3 * x.detach() + torch.randn(c, p_, generator=g)).requires_grad_(True)
        loss = ortho(x, y)
        loss.backward()
        out[f"{name}_x"], out[f"{name}_y"], out[f"{name}_loss"] = x.detach().numpy(), y.detach().numpy(), loss.detach().numpy()
        out[f"{name}_gx"], out[f"{name}_gy"] = x.grad.numpy(), y.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "aux_cases.npz"), **out)


SW_CASES = (  # name, module ("plain" / "sync1" / "sync2"), (n, ch, h, w), num_pergroup, sw_type, tie, affine, training
    ("p2", "plain", (3, 32, 6, 5), 16, 2, False, True, True),
    ("p3", "plain", (2, 32, 7, 4), 16, 3, False, True, True),
    ("p5", "plain", (3, 16, 5, 5), 16, 5, False, True, True),
    ("p5t", "plain", (2, 32, 4, 6), 16, 5, True, False, True),
    ("p2e", "plain", (2, 32, 6, 5), 16, 2, False, True, False),
    ("p3e", "plain", (2, 16, 6, 5), 16, 3, True, True, False),
    ("p5e", "plain", (2, 32, 3, 9), 16, 5, False, False, False),
    ("p2g8", "plain", (2, 24, 8, 5), 8, 2, False, True, True),
    ("p3g4", "plain", (3, 8, 6, 6), 4, 3, False, True, True),
    ("p2big", "plain", (2, 64, 24, 20), 16, 2, False, True, True),
    ("s2", "sync1", (3, 32, 6, 5), 16, 2, False, True, True),
    ("s5", "sync1", (2, 32, 5, 4), 16, 5, False, True, True),
    ("s3e", "sync1", (2, 16, 6, 5), 16, 3, False, True, False),
    ("d2", "sync2", (4, 32, 6, 5), 16, 2, False, True, True),
    ("d5", "sync2", (2, 32, 5, 6), 16, 5, True, True, True),
)


def _sw_inputs(seed, shape, cper, sw_type, tie, affine):
    g = torch.Generator().manual_seed(seed)
    n, ch, h, w = shape
    groups = ch // cper
    x = torch.randn(n, ch, h, w, generator=g) * (0.5 + torch.rand(1, ch, 1, 1, generator=g)) + \
        0.8 * torch.randn(1, ch, 1, 1, generator=g)
    x = x + 0.4 * x.roll(1, dims=1)                          # correlated channels: a non-trivial whitening matrix
    d = {"x": x, "gy": torch.randn(n, ch, h, w, generator=g), "mw": torch.randn(sw_type, generator=g)}
    if not tie:
        d["vw"] = torch.randn(sw_type, generator=g)
    if affine:
        d["weight"], d["bias"] = 1 + 0.3 * torch.randn(ch, generator=g), 0.3 * torch.randn(ch, generator=g)
    a = torch.randn(groups, cper, cper, generator=g)
    d["rmean"] = 0.3 * torch.randn(groups, cper, 1, generator=g)
    d["rcov"] = a @ a.transpose(1, 2) / cper + 0.2 * torch.eye(cper)
    return d


def _sw_run(mod_cls, d, x, gy, cper, sw_type, tie, affine, training):
    m = mod_cls(x.shape[1], num_pergroup=cper, sw_type=sw_type, tie_weight=tie, affine=affine)
    with torch.no_grad():
        m.sw_mean_weight.copy_(d["mw"])
        if not tie:
            m.sw_var_weight.copy_(d["vw"])
        if affine:
            m.weight.copy_(d["weight"])
            m.bias.copy_(d["bias"])
        m.running_mean.copy_(d["rmean"])
        m.running_cov.copy_(d["rcov"])
    m.train(training)
    x = x.clone().requires_grad_(True)
    y = m(x)
    y.backward(gy)
    out = {"y": y.detach(), "gx": x.grad, "gmw": m.sw_mean_weight.grad, "rmean": m.running_mean, "rcov": m.running_cov}
    if not tie:
        out["gvw"] = m.sw_var_weight.grad
    if affine:
        out["gweight"], out["gbias"] = m.weight.grad, m.bias.grad
    return {k: v.detach().clone().numpy() for k, v in out.items()}


def _sw_rank(rank, world, port, case, queue):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    name, _, shape, cper, sw_type, tie, affine, training = case
    d = _sw_inputs(8600 + [c[0] for c in SW_CASES].index(name), shape, cper, sw_type, tie, affine)
    per = shape[0] // world
    cls = load_ref("models/ISW/sync_switchwhiten.py", "ref_ssw").SyncSwitchWhiten2d
    out = _sw_run(cls, d, d["x"][rank * per:(rank + 1) * per], d["gy"][rank * per:(rank + 1) * per], cper, sw_type, tie,
                  affine, training)
    queue.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def make_sw():
    """models/ISW/switchwhiten.py:84-183 (SwitchWhiten2d) and models/ISW/sync_switchwhiten.py (SyncSwitchWhiten2d over
    gloo with 1 and 2 ranks; in the 2-rank cases each rank gets half of the batch and the per-rank outputs are
    stored as ``ref_<case>_r<rank>_*``), the unmodified files loaded by path: output, every gradient, running buffers."""
    import warnings
    import torch.multiprocessing as mp
    warnings.filterwarnings("ignore", message=".*baddbmm.*")
    plain = load_ref("models/ISW/switchwhiten.py", "ref_sw").SwitchWhiten2d
    out = {}
    ctx = mp.get_context("spawn")
    for idx, case in enumerate(SW_CASES):
        name, kind, shape, cper, sw_type, tie, affine, training = case
        d = _sw_inputs(8600 + idx, shape, cper, sw_type, tie, affine)
        for k, v in d.items():
            out[f"in_{name}_{k}"] = v.numpy()
        out[f"in_{name}_cfg"] = np.array([cper, sw_type, int(tie), int(affine), int(training),
                                          {"plain": 0, "sync1": 1, "sync2": 2}[kind]])
        if kind == "plain":
            for k, v in _sw_run(plain, d, d["x"], d["gy"], cper, sw_type, tie, affine, training).items():
                out[f"ref_{name}_{k}"] = v
        else:
            world = 1 if kind == "sync1" else 2
            q = ctx.Queue()
            procs = [ctx.Process(target=_sw_rank, args=(r, world, 29610 + idx, case, q)) for r in range(world)]
            for p_ in procs:
                p_.start()
            got = dict(q.get(timeout=120) for _ in range(world))
            for p_ in procs:
                p_.join()
            for r in range(world):
                for k, v in got[r].items():
                    out[f"ref_{name}_r{r}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "sw_cases.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["bl", "dmap", "isw", "bay", "den", "cov", "aux", "sw"]
    torch.manual_seed(0)
    for w in what:
        globals()[f"make_{w}"]()
