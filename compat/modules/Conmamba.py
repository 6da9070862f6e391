"""Shim for ``modules.Conmamba`` (reference modules/Conmamba.py; imported at modules/Transformer.py:248)."""
from mamba_asr_b200.conmamba import (ConmambaEncoder, ConmambaEncoderLayer, ConvolutionModule,  # noqa: F401
                                      MambaDecoder, MambaDecoderLayer)
