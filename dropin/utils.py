import _path  # noqa: F401
from normalizing_flows_dpfs_b200.utils import *  # noqa: F401,F403 -- drop-in shim for the reference's top-level 'utils' module
from normalizing_flows_dpfs_b200.utils import device, et_distance, compute_normal_density, normalize_log_probs, particle_initialization, freeze_model, unfreeze_model, checkpoint_state, load_model  # noqa: F401
