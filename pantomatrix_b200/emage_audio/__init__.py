"""Same export list as the reference package (/root/reference/models/emage_audio/__init__.py:1-12)."""
from .configuration import EmageAudioConfig, EmageVAEConvConfig, EmageVQVAEConvConfig
from .modeling import EmageAudioModel, EmageVAEConv, EmageVQModel, EmageVQVAEConv

__all__ = [
    "EmageAudioConfig",
    "EmageAudioModel",
    "EmageVQVAEConvConfig",
    "EmageVQVAEConv",
    "EmageVQModel",
    "EmageVAEConvConfig",
    "EmageVAEConv",
]
