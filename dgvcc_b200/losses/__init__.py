from .bl import BL, Bay_Loss, Post_Prob  # noqa: F401
