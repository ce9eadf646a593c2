"""Drop-in for sound_synthesis/modeling/embeddings/dalle_mask_image_embedding.py (same ctor, same state_dict keys)."""
import torch
import torch.nn as nn

from ... import ops


class DalleMaskImageEmbedding(nn.Module):
    def __init__(self, num_embed=8192, spatial_size=[32, 32], embed_dim=3968, trainable=True, pos_emb_type="embedding"):
        super().__init__()
        if isinstance(spatial_size, int):
            spatial_size = [spatial_size, spatial_size]
        self.spatial_size = spatial_size
        self.num_embed = num_embed + 1  # + [MASK]   (reference :21)
        self.embed_dim = embed_dim
        self.trainable = trainable
        self.pos_emb_type = pos_emb_type
        if pos_emb_type != "embedding":
            raise NotImplementedError("only pos_emb_type='embedding' (the Diffsound configs) is implemented")
        self.emb = nn.Embedding(self.num_embed, embed_dim)
        self.height_emb = nn.Embedding(self.spatial_size[0], embed_dim)
        self.width_emb = nn.Embedding(self.spatial_size[1], embed_dim)
        if not trainable:
            for p in self.parameters():
                p.requires_grad = False

    def get_loss(self):
        return None

    @torch.no_grad()
    def forward(self, index, **kwargs):
        assert index.dim() == 2  # B x L
        err = torch.zeros(1, dtype=torch.int32, device=index.device)
        out = ops.embed_tokens(index.contiguous(), self.emb.weight, self.height_emb.weight, self.width_emb.weight, err_flag=err)
        if int(err.item()) != 0:  # same error the reference raises (:42-44)
            raise RuntimeError("IndexError: index out of range in self, max index {}, num embed {}".format(index.max(), self.num_embed))
        return out
