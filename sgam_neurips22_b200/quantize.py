"""Mirror of VectorQuantizer2 (sgam/generative_sensing_module/modules/vqvae/quantize.py:213-381) on the
B200 nearest-neighbour kernel.  Inference surface only: forward, get_multiple_codewords(topk=1), get_codebook_entry."""
import torch

from . import ops


class VectorQuantizer2(torch.nn.Module):
    def __init__(self, n_e, e_dim, beta=0.25, remap=None, unknown_index="random", sane_index_shape=False,
                 legacy=True, kmean_init_codebook_path=None):
        super().__init__()
        if remap is not None:
            raise NotImplementedError("index remapping is not used by any reference config")
        self.n_e, self.e_dim, self.beta, self.legacy = n_e, e_dim, beta, legacy
        self.sane_index_shape = sane_index_shape
        self.embedding = torch.nn.Embedding(n_e, e_dim)
        self.kmean_init_codebook_path = kmean_init_codebook_path
        if kmean_init_codebook_path is None:
            self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)        # quantize.py:231-233
        else:
            import numpy as np
            self.embedding.weight.data.copy_(torch.from_numpy(np.load(kmean_init_codebook_path)))     # quantize.py:234-235
        self.embedding.weight.requires_grad_(False)

    def _nearest(self, z):
        B, D, h, w = z.shape
        tokens = z.permute(0, 2, 3, 1).contiguous().view(-1, D)
        idx, z_q = ops.vq_nearest(tokens, self.embedding.weight.contiguous())
        return idx, z_q.view(B, h, w, D)

    @torch.no_grad()
    def forward(self, z, temp=None, rescale_logits=False, return_logits=False, encoding_indices=None, valid_mask=None):
        """quantize.py:275-319: returns (z_q [B,D,h,w], loss, (None, None, idx [B,h,w]))."""
        assert temp is None or temp == 1.0, "Only for interface compatible with Gumbel"
        assert rescale_logits is False and return_logits is False, "Only for interface compatible with Gumbel"
        if encoding_indices is not None:
            idx = encoding_indices.reshape(-1).to(torch.int64)
            z_q = self.embedding.weight[idx].view(z.shape[0], z.shape[2], z.shape[3], z.shape[1])
        else:
            idx, z_q = self._nearest(z)
        z_q = z_q.permute(0, 3, 1, 2)
        diff = z_q - z
        loss = (1.0 + self.beta) * torch.mean(diff * diff)                    # legacy=True (quantize.py:300-301)
        return z_q, loss, (None, None, idx.view(z.shape[0], z.shape[2], z.shape[3]))

    @torch.no_grad()
    def get_multiple_codewords(self, z, topk=10, sample_number=1, extrapolation_mask=None, return_exp_probility=None, temp=1):
        """quantize.py:344-381 for topk=1 (the pipeline's setting, inference_pipeline.py:24,877): the multinomial
        over one candidate draws it, and pinning to the nearest code (:364-367) is the identity, so the result is
        the arg-min for every token.  Returns (z_qs [B,S,D,h,w], None, (None, None, idx [B,S,h,w]))."""
        B, D, h, w = z.shape
        if topk != 1 or sample_number != 1:
            tokens = z.permute(0, 2, 3, 1).contiguous().view(-1, D)
            m = None
            if extrapolation_mask is not None:
                m = (extrapolation_mask != 0).to(torch.uint8).contiguous()
            self._calls = getattr(self, "_calls", 0) + 1
            out = ops.vq_topk_sample(tokens, self.embedding.weight.contiguous(), int(topk), int(sample_number), (h, w), mask=m,
                                     seed=(torch.initial_seed() * 0x9E3779B1 + self._calls) & 0xFFFFFFFFFFFFFFFF)
            z_qs = out["z_q"].view(B, h, w, sample_number, D).permute(0, 3, 4, 1, 2).contiguous()
            return z_qs, None, (None, None, out["idx"].view(B, h, w, sample_number).permute(0, 3, 1, 2))
        idx, z_q = self._nearest(z)
        return z_q.permute(0, 3, 1, 2).unsqueeze(1), None, (None, None, idx.view(B, 1, h, w))

    def get_codebook_entry(self, indices, shape):
        """quantize.py:327-342: indices flat, shape (b, h, w, c) -> [b, c, h, w]."""
        z_q = self.embedding.weight[indices.reshape(-1).to(torch.int64)]
        if shape is not None:
            z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
        return z_q
