"""Return contracts of encode / forward / loss.

Same three TypedDicts, same keys, as the reference's
``midi_autoencoder/types_helpers.py:15-37`` so that callers such as
``train.py:636,644,672-674`` and ``evaluation.py:63-65`` work unchanged.
``ModelOutput`` may additionally carry ``logits`` (pre-sigmoid decoder output)
which lets ``loss`` run the fused BCE-with-logits kernel.
"""

from typing import TypedDict

from torch import Tensor


class EncoderOutput(TypedDict):
    mu: Tensor
    log_var: Tensor
    pre_latents: Tensor


class ModelOutput(TypedDict, total=False):
    output: Tensor
    input: Tensor
    encoded: EncoderOutput
    latents: Tensor
    logits: Tensor


class LossOutput(TypedDict):
    loss: Tensor
    reconstruction_loss: Tensor
    kld_loss: Tensor
