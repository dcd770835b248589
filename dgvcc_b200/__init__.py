"""dgvcc_b200 -- B200-native (sm_100a) kernels for DGVCC's density-supervision hot path.

Sub-packages mirror the reference's module paths so they drop in:
    dgvcc_b200.losses.bl                     <- losses/bl.py
    dgvcc_b200.utils.dmap_gen                <- utils/dmap_gen.py
    dgvcc_b200.models.ISW.instance_whitening <- models/ISW/instance_whitening.py
"""
__version__ = "0.1.0"
