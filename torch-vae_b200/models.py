"""``VanillaVAE``: drop-in for the reference's ``midi_autoencoder/models.py:7-272``.

Same constructor, methods, attributes, sub-module names and ``state_dict`` keys
(SURVEY.md §8b), so ``train.py:110,205-228,244,325-326,635-650`` and
``evaluation.py:42-65`` run against it unchanged -- but ``forward`` / ``loss`` /
``backward`` execute the hand-written sm_100a kernels behind
``include/vae_b200.h`` instead of ATen.  The ``nn.Conv2d`` / ``nn.BatchNorm2d`` /
``nn.Linear`` objects below are parameter CONTAINERS only (their ``forward`` is
never called); they are created in the reference's order so that the same
``torch.manual_seed`` yields bit-identical initial weights.

Differences a caller can observe (all additive):
  * the two literals that pin the reference to 32x32 inputs (models.py:33,166)
    are computed from ``input_dim`` (an int for square inputs, or ``(H, W)``);
    H and W must be multiples of 16.  At 32x32 the shapes equal the reference's.
  * ``compute_dtype`` (float32 default, or bfloat16) selects the activation dtype;
    parameters and their ``.grad`` stay fp32 in the reference layout.
  * ``forward(x, eps=None)`` / ``reparameterize(mu, log_var, eps=None)`` accept an
    injected eps (parity tests); without it eps comes from a device Philox stream.
  * ``ModelOutput`` carries an extra ``logits`` entry used by the fused loss.
There is no CPU path: tensors must live on a CUDA device.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch
from torch import Tensor, nn

from . import engine, ops
from .types_helpers import EncoderOutput, LossOutput, ModelOutput


def _xavier_like_reference(seq: nn.Sequential) -> None:
    """models.py:227-236: xavier-uniform + zero bias for Conv2d / Linear found in
    ``seq``; BatchNorm weight 1 / bias 0.  ConvTranspose2d is NOT a Conv2d
    subclass, so (as in the reference) it keeps PyTorch's default init."""
    for m in seq.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class VanillaVAE(nn.Module):
    name = "VanillaVAE"

    def __init__(
        self,
        in_channels: int,
        embed_dim: int,
        input_dim: Union[int, Sequence[int]],
        hidden_dims: Optional[List[int]] = None,
        kld_weight: float = 1.0,
        verbose: bool = False,
        *,
        compute_dtype: torch.dtype = torch.float32,
    ):
        super().__init__()
        self.latent_dim = embed_dim
        self.input_dim = input_dim
        self.in_channels = in_channels
        self.verbose = verbose
        self.kld_weight = kld_weight
        self.compute_dtype = compute_dtype
        ops.code(compute_dtype)  # validates

        if hidden_dims is None:
            hidden_dims = [32, 64, 128, 256]
        self.hidden_dims = hidden_dims

        height, width = (input_dim, input_dim) if isinstance(input_dim, int) else tuple(input_dim)
        fh, fw = height, width
        for _ in hidden_dims:
            fh, fw = (fh - 1) // 2 + 1, (fw - 1) // 2 + 1
        self.last_conv_size = fh * fw  # reference literal: 4 (models.py:33), i.e. 2x2 at 32x32
        self.flattened_size = self.last_conv_size * hidden_dims[-1]
        self._feat = (hidden_dims[-1], fh, fw)

        # encoder: [Conv2d(k3,s2,p1) -> BatchNorm2d -> LeakyReLU] per width (models.py:41-52)
        blocks, cin = [], in_channels
        for c in hidden_dims:
            blocks.append(nn.Sequential(nn.Conv2d(cin, c, kernel_size=3, stride=2, padding=1), nn.BatchNorm2d(c), nn.LeakyReLU()))
            cin = c
        self.encoder = nn.Sequential(*blocks)
        _xavier_like_reference(self.encoder)

        # latent heads and decoder input (models.py:55-59): torch default init
        self.fc_mu = nn.Linear(self.flattened_size, self.latent_dim)
        self.fc_var = nn.Linear(self.flattened_size, self.latent_dim)
        self.decoder_input = nn.Linear(self.latent_dim, self.flattened_size)

        hidden_dims.reverse()  # in place, as the reference does (models.py:60)

        blocks = []
        for a, b in zip(hidden_dims[:-1], hidden_dims[1:]):
            blocks.append(
                nn.Sequential(
                    nn.ConvTranspose2d(a, b, kernel_size=3, stride=2, padding=1, output_padding=1),
                    nn.BatchNorm2d(b),
                    nn.LeakyReLU(),
                )
            )
        self.decoder = nn.Sequential(*blocks)
        _xavier_like_reference(self.decoder)

        c = hidden_dims[-1]
        self.final_layer = nn.Sequential(
            nn.ConvTranspose2d(c, c, kernel_size=3, stride=2, padding=1, output_padding=1),
            nn.BatchNorm2d(c),
            nn.LeakyReLU(),
            nn.Conv2d(c, 1, kernel_size=3, stride=1, padding=1),
            nn.Sigmoid(),
        )
        _xavier_like_reference(self.final_layer)

        # Philox stream for reparameterize: (seed, device counter)
        self._rng_seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self._rng_offset: Optional[Tensor] = None
        self._grad_sink = None  # set by parallel.GradientBuckets

    # ------------------------------------------------------------------ helpers
    def manual_seed(self, seed: int, rank: int = 0) -> None:
        """Seed the device RNG used by ``reparameterize`` (per-rank offset follows
        the reference's ``seed + global_rank`` convention, train.py:356-363)."""
        self._rng_seed = (int(seed) + int(rank)) & 0xFFFFFFFFFFFFFFFF
        if self._rng_offset is not None:
            self._rng_offset.zero_()

    def _enc_specs(self):
        return [engine.ConvBlockSpec(False, blk[0].in_channels, blk[0].out_channels, blk[1]) for blk in self.encoder]

    def _dec_specs(self):
        # ConvTranspose2d(cin, cout) == data-gradient of Conv2d(Ci=cout -> Co=cin)
        blks = list(self.decoder) + [self.final_layer]
        return [engine.ConvBlockSpec(True, blk[0].out_channels, blk[0].in_channels, blk[1]) for blk in blks]

    @staticmethod
    def _block_params(blocks):
        out = []
        for blk in blocks:
            out += [blk[0].weight, blk[0].bias, blk[1].weight, blk[1].bias]
        return out

    def _check_input(self, x: Tensor) -> None:
        if not x.is_cuda:
            raise RuntimeError("torch_vae_b200.VanillaVAE runs on CUDA only (no CPU fallback); move inputs with .to('cuda')")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise RuntimeError(f"expected input [B, {self.in_channels}, H, W], got {tuple(x.shape)}")
        n = len(self.encoder)
        if x.shape[2] % (1 << n) or x.shape[3] % (1 << n):
            raise RuntimeError(f"H and W must be multiples of {1 << n}, got {tuple(x.shape[2:])}")
        _, fh, fw = self._feat
        if (x.shape[2] >> n, x.shape[3] >> n) != (fh, fw):
            raise RuntimeError(
                f"mat1 and mat2 shapes cannot be multiplied: input {tuple(x.shape[2:])} does not match "
                f"input_dim={self.input_dim} this model was built for"
            )

    # ---------------------------------------------------------------- public API
    def _prepack(self, x: Tensor):
        """Weight images of every tensor-core conv layer for this step, packed on a side stream
        (ops.PackBatch).  Returns (encoder ConvWeights, decoder ConvWeights) or (None, None)."""
        if self.compute_dtype != torch.bfloat16 or not x.is_cuda or not ops.umma_available():
            return None, None
        B, H, W = x.shape[0], x.shape[2], x.shape[3]
        bwd = torch.is_grad_enabled()
        batch = ops.PackBatch()
        enc, dec = [], []
        h, w = H, W
        for i, blk in enumerate(self.encoder):
            cw = ops.ConvWeights(blk[0].weight, self.compute_dtype, 2)
            batch.add(cw, B, h, w, 0)
            if bwd and (i > 0 or x.requires_grad):
                batch.add(cw, B, h, w, 1)
            enc.append(cw)
            h, w = h // 2, w // 2
        for blk in list(self.decoder) + [self.final_layer]:
            cw = ops.ConvWeights(blk[0].weight, self.compute_dtype, 2)
            h, w = h * 2, w * 2            # the ConvTranspose output is the high-resolution side
            batch.add(cw, B, h, w, 1)
            if bwd:
                batch.add(cw, B, h, w, 0)
            dec.append(cw)
        batch.launch()
        return enc, dec

    def encode(self, x: Tensor, _cws=None) -> EncoderOutput:
        self._check_input(x)
        params = self._block_params(self.encoder)
        infer = engine.inference_mode(self.training, x, *self.fc_mu.parameters(), *self.fc_var.parameters(), *params)
        mu, log_var, pre = engine.EncodeFn.apply(
            self._enc_specs(), self.compute_dtype, self.training, infer, self._grad_sink, _cws, x,
            self.fc_mu.weight, self.fc_mu.bias, self.fc_var.weight, self.fc_var.bias,
            *params,
        )
        return EncoderOutput(mu=mu, log_var=log_var, pre_latents=pre)

    def _decode_both(self, z: Tensor, _cws=None):
        if not z.is_cuda:
            raise RuntimeError("torch_vae_b200.VanillaVAE runs on CUDA only (no CPU fallback)")
        if z.dim() != 2 or z.shape[1] != self.latent_dim:
            raise RuntimeError(f"expected latents [B, {self.latent_dim}], got {tuple(z.shape)}")
        blocks = list(self.decoder) + [self.final_layer]
        last = self.final_layer[3]
        params = self._block_params(blocks)
        infer = engine.inference_mode(self.training, z, *self.decoder_input.parameters(), *last.parameters(), *params)
        return engine.DecodeFn.apply(
            self._dec_specs(), self._feat, self.compute_dtype, self.training, infer, self._grad_sink, _cws, z,
            self.decoder_input.weight, self.decoder_input.bias, last.weight, last.bias,
            *params,
        )

    def decode(self, z: Tensor) -> Tensor:
        return self._decode_both(z)[0]

    def reparameterize(self, mu: Tensor, log_var: Tensor, eps: Optional[Tensor] = None) -> Tensor:
        if eps is None and self._rng_offset is None:
            self._rng_offset = torch.zeros(1, dtype=torch.int64, device=mu.device)
        return engine.ReparamFn.apply(mu, log_var, eps, self._rng_seed, None if eps is not None else self._rng_offset)

    def forward(self, x: Tensor, eps: Optional[Tensor] = None) -> ModelOutput:
        self._check_input(x)
        enc_cws, dec_cws = self._prepack(x)
        encoding = self.encode(x, enc_cws)
        z = self.reparameterize(encoding["mu"], encoding["log_var"], eps)
        probs, logits = self._decode_both(z, dec_cws)
        return ModelOutput(output=probs, input=x, encoded=encoding, latents=z, logits=logits)

    def loss(self, output: ModelOutput) -> LossOutput:
        enc = output["encoded"]
        logits = output.get("logits")
        if logits is not None:
            pred, is_prob = logits, False
        else:  # hand-built ModelOutput: BCE on probabilities, exactly models.py:208
            pred, is_prob = output["output"], True
        loss, parts = engine.ElboFn.apply(pred, output["input"], enc["mu"], enc["log_var"], float(self.kld_weight), is_prob)
        return LossOutput(loss=loss, reconstruction_loss=parts[1], kld_loss=parts[2])

    def sample(self, num_samples: int, current_device, device_rng: bool = False, **kwargs) -> Tensor:
        """models.py:250-263.  By default z is drawn on the host with ``torch.randn``
        exactly like the reference (same seed -> same z); ``device_rng=True`` draws
        it on the GPU from the model's Philox stream instead (no H2D copy)."""
        if device_rng:
            dev = torch.device("cuda", current_device) if isinstance(current_device, int) else torch.device(current_device)
            if self._rng_offset is None:
                self._rng_offset = torch.zeros(1, dtype=torch.int64, device=dev)
            n = num_samples * self.latent_dim
            # host-side counter read is avoided: offset lives on the device
            zero = torch.zeros(n, dtype=torch.float32, device=dev)
            z, _ = ops.reparam_fwd(zero, zero, None, seed=self._rng_seed ^ 0x5A5A5A5A, offset=0, dev_offset=self._rng_offset)
            self._rng_offset.add_((n + 3) // 4)
            z = z.view(num_samples, self.latent_dim)
        else:
            z = torch.randn(num_samples, self.latent_dim).to(current_device)
        return self.decode(z)

    def sample_to_host(self, num_samples: int, out: Optional[Tensor] = None, chunk: int = 512, z: Optional[Tensor] = None) -> Tensor:
        """``sample(n, device).cpu()`` (models.py:250-263 and the copy its callers make) as one pipelined call: z is drawn on
        the host with ``torch.randn`` like the reference, the batch is decoded in chunks, and each chunk's device->host
        copy runs on a side stream under the next chunk's decode.  ``out``: pinned fp32 [n, C, H, W] buffer to fill
        (allocated when absent).  Returns when the images are in host memory."""
        dev = self.decoder_input.weight.device
        if dev.type != "cuda":
            raise RuntimeError("torch_vae_b200.VanillaVAE runs on CUDA only (no CPU fallback)")
        if z is None:
            z = torch.randn(num_samples, self.latent_dim)
        if z.shape[0] != num_samples or chunk <= 0:
            raise RuntimeError("sample_to_host: z must hold num_samples rows and chunk must be positive")
        main = torch.cuda.current_stream(dev)
        copy = ops.copy_stream(dev)
        zd = z.to(dev, non_blocking=True)
        with torch.no_grad():
            for i in range(0, num_samples, chunk):
                img = self.decode(zd[i : i + chunk])
                if out is None:
                    out = torch.empty((num_samples,) + tuple(img.shape[1:]), dtype=torch.float32).pin_memory()
                copy.wait_stream(main)
                with torch.cuda.stream(copy):
                    out[i : i + chunk].copy_(img, non_blocking=True)
                img.record_stream(copy)
        copy.synchronize()
        return out

    def generate(self, x: Tensor, **kwargs) -> Tensor:
        return self.forward(x)["output"]
