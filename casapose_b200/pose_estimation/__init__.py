"""Mirror of the reference package ``casapose.pose_estimation`` for the voting hot path."""
from .pose_evaluation import (estimate_and_evaluate_poses, evaluate_pose_estimates, pose_estimation,  # noqa: F401
                              poses_pnp)
from .ransac_voting import estimate_poses, evaluate_poses, pnp, ransac_voting_layer_all_masks  # noqa: F401
from .voting_layers_2d import CoordLSVotingWeighted  # noqa: F401
from .device_eval import DeviceEvaluator  # noqa: F401
