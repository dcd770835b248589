from .bl import BL, Bay_Loss, PackedBatch, Post_Prob, pack_batch  # noqa: F401
from .lw import lw_loss  # noqa: F401
from .ortho import ortho_loss  # noqa: F401
