"""Tiny stand-ins for the components that stay PyTorch / third-party in the reference pipeline (VAE, CLIP text
encoder, tokenizer).  diffusers / pretrained weights are not available offline; the pipeline only needs their
call surface."""
import zlib

import torch
import torch.nn as nn


class _Dist:
    def __init__(self, mean):
        self.mean = mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + 0.1 * noise


class _Enc:
    def __init__(self, mean):
        self.latent_dist = _Dist(mean)


class _Dec:
    def __init__(self, sample):
        self.sample = sample


class FakeVAE(nn.Module):
    """8x down / up sampling 'VAE' (3 <-> 4 channels) with the AutoencoderKL call surface the pipeline uses."""

    class config:
        block_out_channels = (128, 256, 512, 512)

    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.linspace(-1, 1, 12).reshape(4, 3), requires_grad=False)

    def encode(self, x):
        pooled = torch.nn.functional.avg_pool2d(x, 8)
        return _Enc(torch.einsum("oc,bchw->bohw", self.w.to(x.dtype), pooled))

    def decode(self, z):
        up = torch.nn.functional.interpolate(z, scale_factor=8, mode="nearest")
        return _Dec(torch.einsum("oc,bohw->bchw", self.w.to(z.dtype), up))


class FakeTokenizer:
    def __call__(self, texts, padding=None, max_length=None, truncation=None, return_tensors=None):
        texts = [texts] if isinstance(texts, str) else texts
        ids = torch.zeros((len(texts), max_length), dtype=torch.long)
        for i, t in enumerate(texts):
            for j, wd in enumerate(t.split()[:max_length]):
                ids[i, j] = zlib.crc32(wd.encode()) % 1000 + 1

        class _Out:
            input_ids = ids
        return _Out()


class FakeTextEncoder(nn.Module):
    def __init__(self, max_len=7, dim=96):
        super().__init__()
        self.max_position_embeddings = max_len
        self.emb = nn.Embedding(1001, dim)
        self.pos = nn.Parameter(torch.randn(max_len, dim) * 0.1)

        class _Cfg:
            pass
        self.config = _Cfg()

    def forward(self, ids):
        class _Out:
            pass
        o = _Out()
        o.last_hidden_state = self.emb(ids) + self.pos
        return o
