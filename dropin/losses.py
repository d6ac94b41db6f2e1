import _path  # noqa: F401
from normalizing_flows_dpfs_b200.losses import *  # noqa: F401,F403 -- drop-in shim for the reference's top-level 'losses' module
