"""Stands in for `deephar.models.blocks` (injected into sys.modules by run_reference_backbone.py): the reference's four
parameter-free head builders (blocks.py:217-343) as deephar_b200.keras_compat's recording objects.  Everything else
the reference's model files need comes from the reference itself."""
from deephar_b200.keras_compat import (build_context_aggregation, build_joints_probability,  # noqa: F401
                                       build_softargmax_1d, build_softargmax_2d)
