from .bl import BL, Bay_Loss, PackedBatch, Post_Prob, pack_batch  # noqa: F401
