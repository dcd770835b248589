from .instance_whitening import (InstanceWhitening, get_covariance_matrix, instance_whitening_loss,  # noqa: F401
                                 variance_of_covariance)
from .cov_settings import CovMatrix_IRW, CovMatrix_ISW, make_cov_index_matrix  # noqa: F401
from .switchwhiten import SwitchWhiten2d  # noqa: F401
from .sync_switchwhiten import SyncSwitchWhiten2d  # noqa: F401
