"""Parity of the CUDA ISW path (through the C ABI) with the reference fixtures and the CPU oracle.
Tolerance: rtol 1e-5 (BASELINE.json north_star) with an absolute floor of 5e-7*max|ref| for covariance
entries -- off-diagonal entries of a whitened map are sums with heavy cancellation, and the reference's
own fp32 bmm deviates from an fp64 evaluation by 4-6e-7*max (printed by test_config5_shapes_against_oracle)."""
COV_ATOL = 5e-7
import os

import numpy as np
import pytest
import torch

from oracle import isw_oracle
from helpers import GOLDEN, assert_close

pytestmark = pytest.mark.gpu
CASES = ["c64", "c128", "c48", "margin"]
DEV = "cuda:0"


@pytest.fixture(scope="module")
def fixtures():
    return np.load(os.path.join(GOLDEN, "isw_cases.npz"))


def amax(a):
    return float(np.abs(np.asarray(a)).max())


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("name", CASES)
def test_matches_reference_fixture(fixtures, name, tc, monkeypatch):
    from dgvcc_b200.models.ISW import InstanceWhitening, get_covariance_matrix, instance_whitening_loss
    monkeypatch.setenv("DGVCC_ISW_TENSOR_CORES", str(tc))
    x = torch.from_numpy(fixtures[f"{name}_x"]).to(DEV).requires_grad_(True)
    mask = torch.from_numpy(fixtures[f"{name}_mask"]).to(DEV)
    margin = float(fixtures[f"{name}_margin"])
    c = x.shape[1]
    eye = torch.eye(c, device=DEV)
    y, w = InstanceWhitening(c)(x)
    assert y is w
    cov, b = get_covariance_matrix(w, eye=eye)
    assert b == x.shape[0]
    loss = instance_whitening_loss(w, eye, mask, margin, mask.sum())
    loss.backward()
    ref_cov = fixtures[f"{name}_cov"]
    assert_close(w.detach().cpu(), fixtures[f"{name}_norm"], 1e-5, 1e-6, "instance norm")
    assert_close(cov.detach().cpu(), ref_cov, 1e-5, COV_ATOL * amax(ref_cov), "covariance")
    assert_close(loss.detach().cpu(), fixtures[f"{name}_loss"], 1e-5, 0, "loss")
    gx = fixtures[f"{name}_grad_x"]
    assert_close(x.grad.cpu(), gx, 1e-5, 2e-6 * amax(gx), "grad through norm")
    # gradient w.r.t. the whitened map itself (the tensor the reference feeds to the loss)
    wl = torch.from_numpy(fixtures[f"{name}_norm"]).to(DEV).requires_grad_(True)
    instance_whitening_loss(wl, eye, mask, margin, mask.sum()).backward()
    gw = fixtures[f"{name}_grad_w"]
    assert_close(wl.grad.cpu(), gw, 1e-5, 1e-6 * amax(gw), "grad f_map")


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("shape", [(8, 64, 160, 160), (8, 256, 80, 80), (8, 512, 40, 40)])
def test_config5_shapes_against_oracle(shape, tc, monkeypatch):
    """BASELINE config 5: VGG/ResNet-shaped maps at crop 320."""
    from dgvcc_b200.models.ISW import InstanceWhitening, get_covariance_matrix, instance_whitening_loss
    monkeypatch.setenv("DGVCC_ISW_TENSOR_CORES", str(tc))
    g = torch.Generator().manual_seed(5000 + shape[1])
    x = torch.randn(shape, generator=g)
    c = shape[1]
    mask = isw_oracle.upper_mask(c, 0.5, 5100 + c)
    eye = torch.eye(c)
    xr = x.clone().requires_grad_(True)
    wr = isw_oracle.instance_standardize(xr)
    cov_r, _ = isw_oracle.covariance(wr, eye)
    loss_r = isw_oracle.whitening_loss(wr, eye, mask, 0, mask.sum())
    loss_r.backward()
    # fp64 yardstick for the Gram: the reference's own fp32 error
    cov64, _ = isw_oracle.covariance(wr.detach().double(), eye.double())

    xd = x.to(DEV).requires_grad_(True)
    y, w = InstanceWhitening(c)(xd)
    cov, _ = get_covariance_matrix(w, eye=eye.to(DEV))
    loss = instance_whitening_loss(w, eye.to(DEV), mask.to(DEV), 0, mask.sum().to(DEV))
    loss.backward()
    assert_close(y.detach().cpu(), wr.detach(), 1e-5, 1e-6, "instance norm")
    assert_close(cov.detach().cpu(), cov_r.detach(), 1e-5, COV_ATOL * amax(cov_r.detach()), "covariance")
    assert_close(loss.detach().cpu(), loss_r.detach(), 1e-5, 0, "loss")
    # d|x|/dx = sign(x): a masked covariance entry smaller than the fp32 evaluation error has no well-defined
    # sign (the reference's own fp32 and fp64 evaluations disagree there), and one flipped sign changes two
    # whole channel rows of the gradient.  Those rows are only checked to be of the right magnitude.
    ambiguous = (cov_r.detach().abs() < 2 * COV_ATOL * amax(cov_r.detach())) & (mask > 0)
    rows = torch.zeros(shape[0], c, dtype=torch.bool)
    for b_i, i, j in ambiguous.nonzero().tolist():
        rows[b_i, i] = rows[b_i, j] = True
    got_g, ref_g = xd.grad.cpu(), xr.grad
    keep = ~rows[:, :, None, None].expand_as(ref_g)
    assert_close(torch.where(keep, got_g, ref_g), ref_g, 1e-5, 2e-6 * amax(ref_g), "grad")
    assert float((got_g - ref_g)[~keep].abs().max() if (~keep).any() else 0.0) <= 4 * amax(ref_g)
    print(f"{shape} tc={tc}: {int(rows.sum())} channel rows with a sign-ambiguous covariance entry")
    err_ref = float((cov_r.detach().double() - cov64).abs().max())
    err_ours = float((cov.detach().cpu().double() - cov64).abs().max())
    print(f"{shape} tc={tc}: max |cov - fp64|: reference fp32 {err_ref:.3e}, ours {err_ours:.3e}")


def test_covariance_backward_generic_upstream():
    from dgvcc_b200.models.ISW import get_covariance_matrix
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 64, 12, 12), generator=g)
    up = torch.randn((2, 64, 64), generator=g)
    xr = x.clone().requires_grad_(True)
    cr, _ = isw_oracle.covariance(xr, torch.eye(64))
    (cr * up).sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    cd, _ = get_covariance_matrix(xd)  # eye=None -> built on the map's device (instance_whitening.py:34-35)
    (cd * up.to(DEV)).sum().backward()
    assert_close(cd.detach().cpu(), cr.detach(), 1e-5, COV_ATOL * amax(cr.detach()), "covariance")
    assert_close(xd.grad.cpu(), xr.grad, 1e-5, 1e-6 * amax(xr.grad), "grad")


def test_cpu_tensor_is_rejected():
    from dgvcc_b200.models.ISW import get_covariance_matrix
    with pytest.raises(RuntimeError):
        get_covariance_matrix(torch.zeros(1, 8, 4, 4), eye=torch.eye(8))


@pytest.mark.parametrize("shape", [(2, 64, 1600), (3, 64, 25600), (1, 48, 640), (2, 128, 640), (3, 256, 6400),
                                   (2, 512, 1600), (1, 192, 260)])
def test_tensor_core_gram_partials_directly(shape):
    """dgvcc_isw_gram_tc_partials must take these shapes (no silent fallback) and match an fp64 Gram.
    Long K ranges per CTA exercise the in-TMEM segments + register drains; C <= 64 the two-samples-per-tile mode."""
    from dgvcc_b200 import _native
    b, c, hw = shape
    g = torch.Generator().manual_seed(11 + c)
    x = torch.randn((b, c, hw), generator=g)
    xd = x.to(DEV).contiguous()
    kps = ((hw + 1) // 2 + 31) // 32 * 32      # two splits: up to 12 800 k (25 segments) per CTA
    splits = (hw + kps - 1) // kps
    tile = 64 if c <= 64 else 128
    t1 = (c + tile - 1) // tile
    n_tiles = t1 * (t1 + 1) // 2
    part = torch.zeros((b, splits, n_tiles, tile, tile), device=DEV)
    rc = _native.lib().dgvcc_isw_gram_tc_partials(_native.ptr(xd), b, c, hw, splits, kps, _native.ptr(part),
                                                  _native.stream_ptr(torch.device(DEV)))
    assert rc == 0, f"tensor-core Gram refused shape {shape}: rc={rc}"
    torch.cuda.synchronize()
    tiles = part.sum(1).cpu().double()
    ref = torch.bmm(x.double(), x.double().transpose(1, 2))
    t = 0
    for ti in range(t1):
        for tj in range(ti, t1):
            r0, r1 = ti * tile, min(c, ti * tile + tile)
            c0, c1 = tj * tile, min(c, tj * tile + tile)
            got = tiles[:, t, :r1 - r0, :c1 - c0]
            want = ref[:, r0:r1, c0:c1]
            err = (got - want).abs().max().item()
            assert err <= 5e-6 * ref.abs().max().item(), f"tile ({ti},{tj}) max err {err} vs {ref.abs().max().item()}"
            t += 1


@pytest.mark.parametrize("shape", [(2, 64, 1600), (2, 128, 644), (2, 256, 6400), (1, 512, 1600), (1, 192, 260)])
def test_tensor_core_backward_gemm_directly(shape):
    """dgvcc_isw_sx_tc (dX = S X, MN-major B operand) must take these shapes and match an fp64 product."""
    from dgvcc_b200 import _native
    b, c, hw = shape
    g = torch.Generator().manual_seed(23 + c)
    x = torch.randn((b, c, hw), generator=g)
    s = torch.randn((b, c, c), generator=g) * 0.01
    s = s + s.transpose(1, 2)
    xd, sd = x.to(DEV).contiguous(), s.to(DEV).contiguous()
    dx = torch.full((b, c, hw), float("nan"), device=DEV)
    rc = _native.lib().dgvcc_isw_sx_tc(_native.ptr(sd), _native.ptr(xd), b, c, hw, _native.ptr(dx),
                                       _native.stream_ptr(torch.device(DEV)))
    assert rc == 0, f"tensor-core backward refused shape {shape}: rc={rc}"
    torch.cuda.synchronize()
    ref = torch.bmm(s.double(), x.double())
    err = (dx.cpu().double() - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item(), f"max err {err} vs max |ref| {ref.abs().max().item()}"


@pytest.mark.parametrize("c,hw", [(64, 40 * 40), (256, 20 * 20)])
def test_variance_of_covariance_covstat(c, hw):
    """cal_covstat on an (image, augmented image) pair, models/ISW/__init__.py:93-104."""
    from dgvcc_b200.models.ISW import variance_of_covariance
    g = torch.Generator().manual_seed(c)
    h = int(hw ** 0.5)
    x = torch.randn((2, c, h, h), generator=g)
    eye, rev = torch.eye(c), torch.ones(c, c).triu(diagonal=1)
    ref = isw_oracle.covstat_variance(isw_oracle.instance_standardize(x), eye, rev)
    got = variance_of_covariance(isw_oracle.instance_standardize(x).to(DEV), eye.to(DEV), rev.to(DEV))
    assert_close(got.cpu(), ref, 1e-4, 1e-6 * amax(ref), "variance of covariance")  # a difference of two close covariances


# ------------------------------------------------------------------ CovMatrix_ISW
COV_CASES = ["c16", "c64", "c64r3", "c256"]


@pytest.mark.parametrize("name", COV_CASES)
def test_cov_matrix_isw_matches_reference_fixture(name):
    """The drop-in class replays the statistics the unmodified reference class saw: masks bit-exact."""
    from dgvcc_b200.models.ISW import CovMatrix_ISW
    fx = np.load(os.path.join(GOLDEN, "cov_cases.npz"))
    dim, relax, rounds, batches = fx[f"{name}_cfg"]
    cm = CovMatrix_ISW(dim=int(dim), relax_denom=float(relax))
    eye, rev = cm.get_eye_matrix()
    assert eye.is_cuda and torch.equal(eye.cpu(), torch.eye(int(dim))) and torch.equal(rev.cpu(), torch.ones(int(dim), int(dim)).triu(1))
    for r in range(int(rounds)):
        for b in range(int(batches)):
            cm.set_variance_of_covariance(torch.from_numpy(fx[f"{name}_var_{r}_{b}"]).cuda())
        cm.set_mask_matrix()
        eye, mask, margin, num = cm.get_mask_matrix()
        assert np.array_equal(mask.cpu().numpy(), fx[f"{name}_mask_{r}"])
        ref_num, ref_margin_ret, ref_margin, ref_off = fx[f"{name}_num_{r}"]
        assert float(num) == ref_num and margin == ref_margin_ret == 0
        assert float(cm.margin) == ref_margin and float(cm.num_off_diagonal) == ref_off
    cm.reset_mask_matrix()
    assert cm.mask_matrix is None


def test_topk_mask_ties_and_edges():
    from dgvcc_b200.models.ISW.cov_settings import topk_mask
    g = torch.Generator().manual_seed(3)
    v = torch.rand(3, 5000, generator=g).cuda()
    vals, mask = topk_mask(v, 3, 1234)
    vc = v.cpu()
    ref_vals = ((vc[0] + vc[1]) + vc[2]) / 3   # on the CPU like the reference: IEEE division (torch's CUDA
    assert torch.equal(vals.cpu(), ref_vals)   # division by a scalar multiplies by the reciprocal instead)
    ref = torch.zeros(5000)
    ref[torch.topk(ref_vals, 1234).indices] = 1
    assert torch.equal(mask.cpu(), ref)
    # ties at the threshold (a half-empty variance matrix): exactly k ones, taken in index order
    t = torch.tensor([[0.0, 2.0, 0.0, 1.0, 0.0, 1.0, 1.0, 0.0, -1.0]]).cuda()
    for k, want in ((0, [0] * 9), (1, [0, 1, 0, 0, 0, 0, 0, 0, 0]), (3, [0, 1, 0, 1, 0, 1, 0, 0, 0]),
                    (5, [1, 1, 0, 1, 0, 1, 1, 0, 0]), (9, [1] * 9), (20, [1] * 9)):
        _, m = topk_mask(t, 1, k)
        assert m.cpu().tolist() == [float(x) for x in want], k
    _, m = topk_mask(t, 1, 5, prev_mask=torch.tensor([1.0, 0, 1, 1, 1, 0, 1, 1, 1]).cuda())
    assert m.cpu().tolist() == [1.0, 0, 0, 1, 0, 0, 1, 0, 0]


def test_cal_covstat_to_mask_end_to_end():
    """cal_covstat statistic (variance_of_covariance) -> CovMatrix_ISW -> mask -> loss, all on the device,
    against the oracle restatements on the same inputs."""
    from dgvcc_b200.models.ISW import CovMatrix_ISW, InstanceWhitening, instance_whitening_loss, variance_of_covariance
    from oracle import isw_oracle
    from oracle.cov_settings_oracle import CovMatrixISW
    c = 64
    cm, ocm = CovMatrix_ISW(dim=c, relax_denom=2.0), CovMatrixISW(c, 2.0)
    eye, rev = cm.get_eye_matrix()
    for s in range(3):
        x = torch.randn(2, c, 24, 20, generator=torch.Generator().manual_seed(50 + s))
        w = isw_oracle.instance_standardize(x)
        stat = variance_of_covariance(w.cuda(), eye, rev)
        ostat = isw_oracle.covstat_variance(w, torch.eye(c), torch.ones(c, c).triu(1))
        torch.testing.assert_close(stat.cpu(), ostat, rtol=1e-4, atol=1e-7)
        cm.set_variance_of_covariance(stat)
        ocm.set_variance_of_covariance(stat.cpu())   # same statistic: the selection itself must then agree exactly
    _, mask, margin, num = cm.get_mask_matrix()
    ocm.set_mask_matrix()
    assert torch.equal(mask.cpu(), ocm.mask_matrix) and float(num) == float(ocm.num_sensitive)
    xin = torch.randn(4, c, 24, 20, generator=torch.Generator().manual_seed(99)).cuda().requires_grad_(True)
    _, wt = InstanceWhitening(c)(xin)
    loss = instance_whitening_loss(wt, eye, mask, margin, num)
    loss.backward()
    xo = xin.detach().cpu().requires_grad_(True)
    lo = isw_oracle.whitening_loss(isw_oracle.instance_standardize(xo), torch.eye(c), ocm.mask_matrix, 0, ocm.num_sensitive)
    lo.backward()
    torch.testing.assert_close(loss.cpu(), lo, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(xin.grad.cpu(), xo.grad, rtol=1e-4, atol=1e-6 * float(xo.grad.abs().max()))


def test_non_binary_mask_takes_the_general_backward():
    """A mask with fractional weights is legal input; the factored (sign-matrix) tensor-core backward is only
    promised for 0/1 masks, so this goes through the general 3xTF32 GEMM.  Both against the oracle."""
    from dgvcc_b200.models.ISW import InstanceWhitening, instance_whitening_loss
    c, h, w = 128, 32, 40
    g = torch.Generator().manual_seed(77)
    x = torch.randn(4, c, h, w, generator=g)
    eye = torch.eye(c)
    for binary in (True, False):
        mask = isw_oracle.upper_mask(c, 0.5, 5)
        if not binary:
            mask = mask * (0.25 + torch.rand(c, c, generator=g))
        num = mask.sum()
        xin = x.clone().cuda().requires_grad_(True)
        _, wt = InstanceWhitening(c)(xin)
        loss = instance_whitening_loss(wt, eye.cuda(), mask.cuda(), 0, num.cuda())
        loss.backward()
        xo = x.clone().requires_grad_(True)
        lo = isw_oracle.whitening_loss(isw_oracle.instance_standardize(xo), eye, mask, 0, num)
        lo.backward()
        torch.testing.assert_close(loss.cpu(), lo, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(xin.grad.cpu(), xo.grad, rtol=1e-4, atol=2e-6 * float(xo.grad.abs().max()))


@pytest.mark.parametrize("shape", [(4, 64, 40, 40), (2, 256, 20, 24), (8, 64, 160, 160)])
def test_module_step_replays_as_a_cuda_graph(shape):
    """InstanceWhitening + instance_whitening_loss, forward and backward, captured ONCE with torch.cuda.graph and replayed
    on new inputs: every launch goes to the capturing stream, nothing synchronises or allocates outside the graph's pool
    after the warm-up -- so the ISW loss can sit inside a graph-captured training step (no launch gaps: what the module's
    seven short kernels need).  Replayed results are bit-identical to eager ones (the kernels are deterministic)."""
    from dgvcc_b200.models.ISW import InstanceWhitening, instance_whitening_loss
    dev = torch.device("cuda:0")
    b, c, h, w = shape
    gen = torch.Generator().manual_seed(77)
    xs = [torch.randn(b, c, h, w, generator=gen).to(dev) for _ in range(3)]
    mask = (torch.rand(c, c, generator=gen) < 0.5).float().triu(1).to(dev)
    eye, nrm = torch.eye(c, device=dev), mask.sum()
    iw = InstanceWhitening(c)

    def step(x):
        _, wt = iw(x)
        loss = instance_whitening_loss(wt, eye, mask, 0, nrm)
        loss.backward()
        return loss

    eager = []
    for x in xs:
        xe = x.clone().requires_grad_(True)
        eager.append((step(xe).detach().clone(), xe.grad.clone()))
    static = xs[0].clone().requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):            # warm-up on a side stream, as torch's graph recipe asks
        for _ in range(2):
            static.grad = None
            step(static)
    torch.cuda.current_stream().wait_stream(side)
    static.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_loss = step(static)
    for x, (l_ref, g_ref) in zip(xs, eager):
        with torch.no_grad():
            static.copy_(x)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(static_loss.detach(), l_ref) and torch.equal(static.grad, g_ref)
