"""Drop-in replacements for the reference's `models.camn_audio` and `models.disco_audio` packages
(BASELINE configs[2], [3]; same export lists as the reference __init__.py files)."""
from .modeling import (CamnAudioConfig, CamnAudioModel, CamnAudioPreTrainedModel, DiscoAudioConfig, DiscoAudioModel,
                       DiscoAudioPreTrainedModel)

__all__ = ["CamnAudioConfig", "CamnAudioModel", "CamnAudioPreTrainedModel", "DiscoAudioConfig", "DiscoAudioModel",
           "DiscoAudioPreTrainedModel"]
