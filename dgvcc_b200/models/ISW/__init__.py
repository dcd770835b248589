from .instance_whitening import (InstanceWhitening, get_covariance_matrix, instance_whitening_loss,  # noqa: F401
                                 variance_of_covariance)
