"""usdm_b200 — B200-native (sm_100a) token-Voicebox mel generator: drop-in for the sampling path of
ishine/usdm `src/decoder` (voicebox.model.voicebox.Voicebox.generate / sample / solve_* / estimator) and of the BigVGAN vocoder forward that follows it
(voicebox.vocoder.models.BigVGAN.forward)."""
from .voicebox import Voicebox  # noqa: F401
from .vocoder import BigVGAN  # noqa: F401
from . import engine  # noqa: F401

__all__ = ["Voicebox", "BigVGAN", "engine"]
