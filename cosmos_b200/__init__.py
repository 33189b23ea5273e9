"""cosmos_b200 - B200-native COSMOS loss head (InfoNCE + cross-modality self-distillation + EMA
teacher update) behind the reference's open_clip loss API.  See DESIGN.md / INTEGRATION.md."""
from .loss import ClipLoss, COSMOSLoss, CoCaLoss, DistillClipLoss, SigLipLoss, gather_features  # noqa: F401
from .ema import EmaPlan, clamp_logit_scales_, ema_update_  # noqa: F401
from .infonce import Comm, pairs_infonce  # noqa: F401
from .retrieval import retrieval_ranks  # noqa: F401

__version__ = "0.1.0"
