from pantomatrix_b200.lstm_audio import CamnAudioConfig, CamnAudioModel, CamnAudioPreTrainedModel  # noqa: F401

__all__ = ["CamnAudioConfig", "CamnAudioModel", "CamnAudioPreTrainedModel"]
