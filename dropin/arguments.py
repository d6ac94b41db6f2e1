import _path  # noqa: F401
from normalizing_flows_dpfs_b200.arguments import *  # noqa: F401,F403 -- drop-in shim for the reference's top-level 'arguments' module
