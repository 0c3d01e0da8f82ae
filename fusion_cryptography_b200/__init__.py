"""fusion_cryptography_b200 -- B200-native (sm_100a CUDA) implementation of the batched Fusion
lattice-signature hot path, behind the unchanged ``fusion.fusion`` API of
geometry-labs/fusion-cryptography.

    from fusion_cryptography_b200 import native          # batched device API (torch tensors as buffers)
    from fusion_cryptography_b200.fusion import fusion   # drop-in for the reference's fusion.fusion

There is no CPU fallback: importing ``native`` needs libfusion_b200.so (built by
``python -m fusion_cryptography_b200.build``) and using it needs a CUDA device.
"""
__version__ = "0.1.0"
