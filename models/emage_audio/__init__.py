from pantomatrix_b200.emage_audio import *  # noqa: F401,F403
from pantomatrix_b200.emage_audio import __all__  # noqa: F401
