"""conformer_b200 - B200 (sm_100a) native Conformer encoder forward path.

Drop-in mirrors of Alan-404/Conformer's encoder-path nn.Modules over a C-ABI library of
hand-written CUDA kernels.  Importing the package does not need a GPU; calling
any forward does.
"""
from .modules import (GLU, Conformer, ConformerBlock, ConvolutionModule, ConvolutionSubsampling, Decoder, Encoder,
                      FeedForwardModule, MultiHeadSelfAttentionModule, RelativeMultiHeadAttention,
                      RelativePositionalEncoding, Swish, generate_padding_mask, greedy_ctc_decode)
from .frontend import LogMelFrontEnd
from .checkpoint import CheckpointManager, load_checkpoint, load_model, save_checkpoint
from .synth import CONFORMER_L, CONFORMER_M, CONFORMER_S, EncoderConfig


def set_pdl(enable: bool) -> None:
    """Programmatic dependent launch of the library's kernels (process-wide, default on; `cfm_set_pdl`): best
    single-stream latency on, slightly better multi-stream throughput off.  Results do not depend on it."""
    from ._lib import lib
    lib().cfm_set_pdl(1 if enable else 0)


__all__ = ["GLU", "Conformer", "ConformerBlock", "ConvolutionModule", "ConvolutionSubsampling", "Decoder", "Encoder",
           "FeedForwardModule", "MultiHeadSelfAttentionModule", "RelativeMultiHeadAttention",
           "RelativePositionalEncoding", "Swish", "generate_padding_mask", "greedy_ctc_decode", "EncoderConfig", "CONFORMER_S",
           "CONFORMER_M", "CONFORMER_L", "LogMelFrontEnd", "CheckpointManager", "load_checkpoint", "load_model", "save_checkpoint", "set_pdl"]
