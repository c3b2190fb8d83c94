from .adversarial import DiscriminatorAdversarialLoss, FeatureMatchLoss, GeneratorAdversarialLoss  # noqa: F401
from .spectral import MelSpectrogramLoss, MultiResolutionSTFTLoss  # noqa: F401
