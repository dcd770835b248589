"""Drop-in for the reference's ``models/ISW/sync_switchwhiten.py`` (and ``models/SW/ops/sync_switchwhiten.py``).

``SyncSwitchWhiten2d`` is ``SwitchWhiten2d`` whose batch mean / covariance -- and, in the backward, their adjoints --
are averaged over the default process group exactly where SyncMeanCov does it (sync_switchwhiten.py:21,25,44,45):
four all-reduces per step on [g, c] / [g, c, c] fp64 device tensors (NCCL on the GPUs), everything else local.

Like the reference, a process group is needed only in TRAINING mode: SyncMeanCov makes no ``dist`` call when
``training`` is false, neither in forward (sync_switchwhiten.py:27-28) nor in backward (:48-55 only rescales), so
single-process validation / inference works without ``init_process_group`` (models/ISW/Resnet.py builds this
layer for ``iw=5`` and the reference code base never initialises a group).
"""
from .switchwhiten import SwitchWhiten2d, _Exchange


class SyncSwitchWhiten2d(SwitchWhiten2d):
    _sw_types = (2, 3, 4, 5)   # sync_switchwhiten.py:84-86
    _sync = True               # eval-mode backward keeps SyncMeanCov's 1 / (n * hw) scaling (sync_switchwhiten.py:48-55)

    def _exchange(self):
        return _Exchange() if self.training else None
