"""Config classes of the drop-in EMAGE modules.

Same names, `model_type`s and constructor contract as the reference
(/root/reference/models/emage_audio/configuration_emage_audio.py:4-32): an optional OmegaConf node is
splatted into keyword arguments, everything else is a plain `transformers.PretrainedConfig`, so
`config.json` files written by either implementation load in the other.  omegaconf is optional here
(it is only needed when a config node is actually passed).
"""
from transformers import PretrainedConfig


def _splat(config_obj, kwargs):
    if config_obj is not None:
        try:
            from omegaconf import OmegaConf
            kwargs.update(OmegaConf.to_container(config_obj, resolve=True))
        except ImportError:                      # plain dict / namespace when omegaconf is absent
            kwargs.update(dict(config_obj) if not hasattr(config_obj, "__dict__") else vars(config_obj))
    return kwargs


class EmageAudioConfig(PretrainedConfig):
    model_type = "emage_audio"

    def __init__(self, config_obj=None, **kwargs):
        super().__init__(**_splat(config_obj, kwargs))


class EmageVQVAEConvConfig(PretrainedConfig):
    model_type = "emage_vqvaeconv"

    def __init__(self, config_obj=None, **kwargs):
        super().__init__(**_splat(config_obj, kwargs))


class EmageVAEConvConfig(PretrainedConfig):
    model_type = "emage_vaeconv"

    def __init__(self, config_obj=None, **kwargs):
        super().__init__(**_splat(config_obj, kwargs))
