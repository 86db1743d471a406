"""Mirror of the reference package ``casapose.pose_estimation`` for the voting hot path."""
from .ransac_voting import ransac_voting_layer_all_masks  # noqa: F401
from .voting_layers_2d import CoordLSVotingWeighted  # noqa: F401
