import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The sharded Bayesian loss is also tested with up to 8 "ranks" inside ONE process (LocalComm: one stream per rank plus
# the library's side stream).  A rank's wait kernel spins until a peer's kernel has run; with the default 8 hardware
# queues two of those streams can share a queue, and the spinning kernel then blocks the very kernel it waits for.
# One queue per stream (must be set before the CUDA context exists; real runs have one rank per process).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_report_header(config):
    """Which library the suite runs through: the product build, or the checked one (DGVCC_BOUNDS_CHECK=1)."""
    try:
        from dgvcc_b200 import _native
        lib = _native.lib()
        return f"dgvcc library: {lib._name} (device-side index checks: {'ON' if lib.dgvcc_bounds_checked() else 'off'})"
    except Exception as exc:   # the tests themselves will say why
        return f"dgvcc library: not loadable ({type(exc).__name__}: {exc})"


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def warm_cpu_oracle():
    """The first multi-threaded torch-CPU evaluation of a process is not reproducible on the GPU box's host
    (oracle.warm_up documents the measurement); take it before any test takes a reference value."""
    import oracle
    oracle.warm_up()


# ---------------------------------------------------------------------------------------------------------------
# Parity margins.  Every tolerance check (helpers.assert_close, torch.testing.assert_close, numpy's assert_allclose)
# records worst |err| / gate; the session summary prints one line per GPU parity test, so the driver's log shows
# how close to its gate each test sits ("nobody can see how close to the edge they sit").
def _install_margin_hooks():
    import numpy as np
    import torch
    import helpers

    torch_close, np_close = torch.testing.assert_close, np.testing.assert_allclose

    def _record(actual, expected, rtol, atol, what):
        try:
            a = torch.as_tensor(np.asarray(actual.detach().cpu()) if isinstance(actual, torch.Tensor) else np.asarray(actual))
            e = torch.as_tensor(np.asarray(expected.detach().cpu()) if isinstance(expected, torch.Tensor) else np.asarray(expected))
            if a.shape != e.shape or not (a.is_floating_point() or e.is_floating_point()):
                return
            a, e = a.double(), e.double()
            fin = torch.isfinite(a) & torch.isfinite(e)
            helpers.record_margin(what, (a - e).abs()[fin], (atol + rtol * e.abs())[fin], rtol, atol)
        except Exception:  # bookkeeping must never fail a test
            pass

    def torch_wrapper(actual, expected, *args, rtol=None, atol=None, **kw):
        if rtol is not None and atol is not None:
            _record(actual, expected, rtol, atol, "torch.testing.assert_close")
        return torch_close(actual, expected, *args, rtol=rtol, atol=atol, **kw)

    def np_wrapper(actual, desired, rtol=1e-7, atol=0, *args, **kw):
        _record(actual, desired, rtol, atol, "np.testing.assert_allclose")
        return np_close(actual, desired, rtol, atol, *args, **kw)

    torch.testing.assert_close = torch_wrapper
    np.testing.assert_allclose = np_wrapper


_install_margin_hooks()


def pytest_terminal_summary(terminalreporter):
    import helpers
    worst, info, acc = {}, {}, {}
    for test, what, ratio, rtol, atol in helpers.MARGINS:
        if "(accuracy)" in what:        # never asserted: distance to the fp64 evaluation, of the kernels and of the reference
            acc.setdefault(test, {})[what] = max(ratio, acc.get(test, {}).get(what, 0.0))
            continue
        if "(informational)" in what:   # never asserted: e.g. the same data against SURVEY 8d's tighter floor
            if test not in info or ratio > info[test][1]:
                info[test] = (what, ratio, rtol, atol)
            continue
        if test not in worst or ratio > worst[test][1]:
            worst[test] = (what, ratio, rtol, atol)
    gpu_only = {t: v for t, v in worst.items() if "_gpu.py" in t}
    if not gpu_only:
        return
    tr = terminalreporter
    tr.write_line("")
    tr.write_line(f"parity margins: worst |err| / gate per test (1.0 = at the gate), {len(gpu_only)} tests")
    for test, (what, ratio, rtol, atol) in sorted(gpu_only.items(), key=lambda kv: -kv[1][1]):
        tr.write_line(f"  {ratio:8.3g}  {test}  [{what}; rtol {rtol:g}, atol {atol:.3g}]")
    over = {t: v for t, v in info.items() if v[1] > 1.0}
    if info:
        tr.write_line(f"informational margins (not asserted): {len(info)} tests, {len(over)} of them beyond 1.0:")
        for test, (what, ratio, rtol, atol) in sorted(info.items(), key=lambda kv: -kv[1][1])[:12]:
            tr.write_line(f"  {ratio:8.3g}  {test}  [{what}]")
    try:
        _write_accuracy(tr, acc)
    except Exception:   # bookkeeping must never change the session's outcome
        pass


def _write_accuracy(tr, acc):
    acc = {t: v for t, v in acc.items() if "_gpu.py" in t}
    if acc:
        # "also print error vs the fp64 oracle" -- how far the kernels and the reference's own fp32 evaluation
        # sit from the fp64 evaluation of the same formulas, in units of the parity gate (rtol 1e-5 of the fp64 value)
        tr.write_line(f"accuracy against the fp64 evaluation (not asserted; 1.0 = 1e-5 relative), {len(acc)} tests:")
        for test, kinds in sorted(acc.items(), key=lambda kv: -max(kv[1].values()))[:16]:
            tr.write_line(f"  {test}  " + "; ".join(f"{k.replace(' (accuracy)', '')} {v:.3g}" for k, v in sorted(kinds.items())))
