"""`sae_auto_interp.sae`: the TopK SAE module backed by the B200 engine (same public names as the reference)."""
from . import config as _config, sae as _sae

Sae, EncoderOutput, ForwardOutput = _sae.Sae, _sae.EncoderOutput, _sae.ForwardOutput
SaeConfig, TrainConfig = _config.SaeConfig, _config.TrainConfig
__all__ = ["Sae", "SaeConfig", "TrainConfig"]
