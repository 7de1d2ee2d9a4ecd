"""Drop-in alias: ``import aecf`` resolves to the B200-native implementation in ``aecf_b200``."""
from aecf_b200 import *  # noqa: F401,F403
from aecf_b200 import __all__, __version__  # noqa: F401
from aecf_b200 import linear, project_tokens  # noqa: F401  (extensions beyond the reference's four names; not in __all__)
