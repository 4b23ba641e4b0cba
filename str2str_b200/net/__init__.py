from .denoising_ipa import DenoisingNet, EmbeddingModule  # noqa: F401
from .ipa import InvariantPointAttention, TranslationIPA  # noqa: F401
