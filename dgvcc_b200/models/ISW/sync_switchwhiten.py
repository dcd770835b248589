"""Drop-in for the reference's ``models/ISW/sync_switchwhiten.py`` (and ``models/SW/ops/sync_switchwhiten.py``).

``SyncSwitchWhiten2d`` is ``SwitchWhiten2d`` whose batch mean / covariance -- and, in the backward, their adjoints --
are averaged over the default process group exactly where SyncMeanCov does it (sync_switchwhiten.py:21,25,44,45):
four all-reduces per step on [g, c] / [g, c, c] fp64 device tensors (NCCL on the GPUs), everything else local.
Like the reference it needs an initialised process group, also in eval mode and with one rank.
"""
from .switchwhiten import SwitchWhiten2d, _Exchange


class SyncSwitchWhiten2d(SwitchWhiten2d):
    _sw_types = (2, 3, 4, 5)   # sync_switchwhiten.py:84-86

    def _exchange(self):
        return _Exchange()
