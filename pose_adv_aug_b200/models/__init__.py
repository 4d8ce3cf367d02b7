from .asn_stacked_hg import *   # noqa: F401,F403
