"""Randomised soak of the ISW path (InstanceWhitening + covariance + whitening loss, forward + backward) against the
CPU oracle: random batch, channel count (tensor-core and CUDA-core shapes), spatial size, mask density, binary and
weighted masks.  Gates as in tests/test_isw_gpu.py; channel rows touched by a sign-ambiguous covariance entry are only
checked for magnitude (d|x|/dx has no defined sign there).

    python scripts/soak_isw.py [seconds]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import isw_oracle
from dgvcc_b200.models.ISW import InstanceWhitening, get_covariance_matrix, instance_whitening_loss

oracle.warm_up()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
dev = torch.device("cuda:0")
COV_ATOL = 5e-7
rng = np.random.default_rng(555)
t0, cases, worst_cov, worst_g, amb_rows = time.time(), 0, 0.0, 0.0, 0
while time.time() - t0 < budget:
    b = int(rng.integers(1, 7))
    c = int(rng.choice([8, 24, 32, 48, 64, 96, 128, 160, 256, 320]))
    h, w = int(rng.integers(3, 40)), int(rng.integers(3, 40))
    if rng.random() < 0.5:
        w = (w + 3) // 4 * 4  # HW % 4 == 0: the TMA path
    os.environ["DGVCC_ISW_TENSOR_CORES"] = str(int(rng.integers(0, 2)))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn(b, c, h, w, generator=g) * (0.5 + 2 * torch.rand(1, c, 1, 1, generator=g)) + torch.randn(1, c, 1, 1, generator=g)
    mask = isw_oracle.upper_mask(c, float(rng.uniform(0.1, 1.0)), int(rng.integers(1 << 30)))
    if rng.random() < 0.25:
        mask = mask * (0.25 + torch.rand(c, c, generator=g))
    if mask.sum() == 0:
        continue
    eye = torch.eye(c)
    xr = x.clone().requires_grad_(True)
    wr = isw_oracle.instance_standardize(xr)
    cov_r, _ = isw_oracle.covariance(wr, eye)
    loss_r = isw_oracle.whitening_loss(wr, eye, mask, 0, mask.sum())
    loss_r.backward()
    xd = x.to(dev).requires_grad_(True)
    _, wt = InstanceWhitening(c)(xd)
    cov, _ = get_covariance_matrix(wt, eye=eye.to(dev))
    loss = instance_whitening_loss(wt, eye.to(dev), mask.to(dev), 0, mask.sum().to(dev))
    loss.backward()
    cr = cov_r.detach().double(); m = float(cr.abs().max())
    ecov = float(((cov.detach().cpu().double() - cr).abs() / (1e-5 * cr.abs() + COV_ATOL * m)).max())
    # loss: a sum of |cov * mask| entries -> absolute floor follows the summed magnitude
    el = abs(float(loss.detach()) - float(loss_r)) / max(abs(float(loss_r)), 1e-30)
    ambiguous = (cov_r.detach().abs() < 2 * COV_ATOL * m) & (mask > 0)
    rows = torch.zeros(b, c, dtype=torch.bool)
    for b_i, i, j in ambiguous.nonzero().tolist():
        rows[b_i, i] = rows[b_i, j] = True
    got_g, ref_g = xd.grad.cpu().double(), xr.grad.double()
    keep = ~rows[:, :, None, None].expand_as(ref_g)
    gm = float(ref_g.abs().max())
    eg = float((((got_g - ref_g).abs() / (1e-5 * ref_g.abs() + 2e-6 * gm + 1e-300))[keep]).max()) if keep.any() else 0.0
    if ecov > 1.0 or el > 2e-5 or eg > 1.0:
        raise SystemExit(f"MISMATCH: B={b} C={c} HW={h}x{w} tc={os.environ['DGVCC_ISW_TENSOR_CORES']}: cov err/tol {ecov:.3g}, "
                         f"loss rel err {el:.3g}, grad err/tol {eg:.3g} ({int(rows.sum())} ambiguous rows)")
    worst_cov, worst_g = max(worst_cov, ecov), max(worst_g, eg)
    amb_rows += int(rows.sum())
    cases += 1
print(f"soak ok: {cases} random cases, worst covariance err/tol {worst_cov:.2f}, worst gradient err/tol {worst_g:.2f}, "
      f"{amb_rows} sign-ambiguous channel rows excluded, {time.time() - t0:.0f} s")
