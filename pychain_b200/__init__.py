"""pychain_b200 -- B200-native LF-MMI chain-loss engine behind pychain's API."""
from .graph import *  # noqa: F401,F403
