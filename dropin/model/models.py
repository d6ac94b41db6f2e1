import _path  # noqa: F401
from normalizing_flows_dpfs_b200.model.models import *  # noqa: F401,F403 -- drop-in shim for the reference's 'model.models' module
