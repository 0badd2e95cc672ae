from pantomatrix_b200.lstm_audio import DiscoAudioConfig, DiscoAudioModel, DiscoAudioPreTrainedModel  # noqa: F401

__all__ = ["DiscoAudioConfig", "DiscoAudioModel", "DiscoAudioPreTrainedModel"]
