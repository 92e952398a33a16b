"""B200-native (sm_100a) Conformer-encoder hot path of icadriani/nn_conformer_for_speech_recognition.

Public surface = the reference's own seams:
    ConvSubSampling(hp, in_size, out_nodes)           <- lib.convsubsampling.ConvSubSampling
    Conformer(d, heads, ffn, layers, kernel, dropout) <- torchaudio.models.Conformer (asrnn.py:29)
    specaugment.SpecAugment / time_warping / ...      <- ASRNN methods (asrnn.py:91-192)
    patch_asrnn(model)                                 swaps all three into a reference ASRNN instance
Everything computes in libb200conformer.so (hand-written CUDA, C ABI in include/b200conformer.h); importing the
package does not need a GPU, calling it does -- there is no CPU fallback.
"""
from . import _lib, ops, specaugment  # noqa: F401
from ._lib import B2CError  # noqa: F401
from .conformer import Conformer, ConformerLayer  # noqa: F401
from .convsubsampling import ConvSubSampling  # noqa: F401
from .ctc import CTCLoss, ctc_loss_from_logits  # noqa: F401
from .encoder import ConformerEncoder, Linear  # noqa: F401
from .frontend import LogMelFrontend, PaddedBatcher  # noqa: F401
from .lstm import LSTM, lstm_from_reference  # noqa: F401
from .prefetch import HostPrefetcher, ResultFetcher  # noqa: F401
from .relpos import RelativePositionalEmbeddings  # noqa: F401
from .patch import conformer_from_torchaudio, convsub_from_reference, patch_asrnn  # noqa: F401

__version__ = "0.1.0"
