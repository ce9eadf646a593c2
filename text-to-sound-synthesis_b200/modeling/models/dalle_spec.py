"""Drop-in for sound_synthesis/modeling/models/dalle_spec.py::DALLE: same constructor keys, the same
`generate_content` / `decode_to_img` / `get_ema_model` surface and the same state_dict prefixes (`transformer.*`,
`content_codec.*`, `first_stage_permuter.*`), so `ckpt["model"]` loads with strict=False and `ckpt["ema"]` overlays
`get_ema_model()` exactly as generate_samples_batch.py:78-85 does.

Differences that are deliberate (and documented in INTEGRATION.md):
  * truncation (`sample_type` 'top0.85r' / 'top100p') is passed to the fused sampler kernel instead of monkey-patching
    `predict_start` with sort/cumsum torch ops (reference :146-177, :207-210);
  * the text tokenizer / CLIP tower is a 'next' row: without a `condition_codec`, captions arrive
    as pre-computed embeddings `batch['condition_embed']` (B, 77, 512), the reference's own bypass
    (diffusion_transformer.py:623-627).
"""
import torch
from torch import nn

from ...utils.misc import instantiate_from_config


def disabled_train(self, mode=True):
    return self


class DALLE(nn.Module):
    def __init__(self, *, content_info={"key": "image"}, condition_info={"key": "text"}, content_codec_config, condition_codec_config,
                 first_stage_permuter_config, diffusion_config):
        super().__init__()
        self.content_info = content_info
        self.condition_info = condition_info
        model = instantiate_from_config(content_codec_config).eval()
        model.train = disabled_train.__get__(model)
        self.content_codec = model
        self.condition_codec = instantiate_from_config(condition_codec_config)
        self.transformer = instantiate_from_config(diffusion_config)
        self.first_stage_permuter = instantiate_from_config(config=first_stage_permuter_config)
        self.truncation_forward = False

    @property
    def device(self):
        return self.transformer.device

    def get_ema_model(self):
        return self.transformer

    @torch.no_grad()
    def decode_to_img(self, index, zshape, stage="first"):
        """Token ids in the transformer's (column-major) order -> mel (B,1,80,848)  (reference :80-91).  The un-permute and the
        codebook gather are one kernel inside VQModel.decode_tokens."""
        if stage != "first":
            raise NotImplementedError
        return self.content_codec.decode_tokens(index, (zshape[2], zshape[3]))

    @torch.no_grad()
    def prepare_condition(self, batch, condition=None):
        if self.condition_codec is None:
            emb = batch["condition_embed"] if condition is None else condition
            return {"condition_token": None, "condition_mask": None, "condition_embed_token": emb.to(self.device)}
        cond = batch[self.condition_info["key"]] if condition is None else condition
        cond = self.condition_codec.get_tokens(cond)
        return {"condition_" + k: (v.to(self.device) if torch.is_tensor(v) else v) for k, v in cond.items()}

    @torch.no_grad()
    def generate_content(self, *, batch, condition=None, filter_ratio=0.5, temperature=1.0, content_ratio=0.0, replicate=1,
                         return_att_weight=False, sample_type="top0.85r"):
        self.eval()
        condition = self.prepare_condition(batch=batch, condition=condition)
        if replicate != 1:
            for k in condition.keys():
                if condition[k] is not None:
                    condition[k] = torch.cat([condition[k] for _ in range(replicate)], dim=0)
        parts = sample_type.split(",")
        # 'top0.85r,q0.3': with probability 0.3 a step's p_sample is applied a second time at the same t (p_sample_with_truncation, reference :135-143)
        self.transformer.resample_rate = float(parts[1][1:]) if (len(parts) > 1 and parts[1][:1] == "q") else 0.0
        self.transformer.truncation = parts[0] if parts[0][:3] == "top" else None
        emb = condition.get("condition_embed_token", None)
        bsz = (condition["condition_token"] if condition["condition_token"] is not None else emb).shape[0]
        kw = dict(condition_token=condition["condition_token"], condition_mask=condition.get("condition_mask", None), condition_embed=emb,
                  content_token=None, filter_ratio=filter_ratio, temperature=temperature, return_att_weight=return_att_weight, return_logits=False,
                  print_log=False, sample_type=sample_type, batch_size=bsz)
        if len(parts) == 2 and parts[1][:4] == "fast":
            trans_out = self.transformer.sample_fast(skip_step=int(parts[1][4:]), **kw)
        else:
            trans_out = self.transformer.sample(**kw)
        zshape = (trans_out["content_token"].shape[0], self.content_codec.quantize.e_dim, *self._grid())
        content = self.decode_to_img(trans_out["content_token"], zshape)
        self.train()
        return {"content": content, "content_token": trans_out["content_token"]}

    def _grid(self):
        p = self.first_stage_permuter
        return (p.H, p.W)

    @torch.no_grad()
    def reconstruct(self, input):
        """mel (B,1,80,848) -> encoder -> nearest codes -> decoder (reference :250-262; the reference body calls a VQModel.get_tokens that
        does not exist in its codec -- `sefl.encode`, spec_codec/vqgan.py:84-85 -- so this implements the evident intent with DALLE.get_tokens)."""
        if torch.is_tensor(input):
            input = input.to(self.device)
        quant_z, indices = self.get_tokens(input)
        return self.decode_to_img(indices, quant_z.shape)

    @torch.no_grad()
    def sample(self, batch, clip=None, temperature=1.0, return_rec=True, filter_ratio=[0, 0.5, 1.0], content_ratio=[1], return_att_weight=False,
               return_logits=False, sample_type="normal", **kwargs):
        """Training-time preview sampler (reference :264-338), the call `Solver.sample` makes every `sample_iterations`
        (engine/solver_spec.py:209-213: `model.sample(batch=batch, step=self.last_iter)`).  For each filter_ratio fr the ground-truth tokens are
        noised to t = 100*fr - 1 (fr = 0: all-[MASK] start) and denoised back; every grid is decoded to a mel.  Returns the reference's dict:
        'condition', 'input_image', 'reconstruction_image', 'cond1_cont{cr}_fr{fr}_image' (+ 'logits')."""
        if return_att_weight:
            raise NotImplementedError("attention maps are never materialised by the fused attention kernels (the reference's own path reads "
                                      "self.content.token_shape, which does not exist: dalle_spec.py:330)")
        if sample_type == "debug":
            raise NotImplementedError("sample_debug does not exist in the reference's DiffusionTransformer either")
        self.eval()
        condition = self.prepare_condition(batch)
        content = self.prepare_content(batch)
        content_samples = {"input_image": batch.get(self.content_info["key"])}
        B = content["content_token"].shape[0]
        zshape = content["content_quant"].shape if "content_quant" in content else (B, self.content_codec.quantize.e_dim, *self._grid())
        if return_rec:
            content_samples["reconstruction_image"] = self.decode_to_img(content["content_token"], zshape)
        prev_trunc, prev_rate = self.transformer.truncation, self.transformer.resample_rate
        self.transformer.truncation, self.transformer.resample_rate = None, 0.0  # the reference's sample() path does not wrap predict_start
        for fr in filter_ratio:
            for cr in content_ratio:
                num_content_tokens = int(content["content_token"].shape[1] * cr)
                if num_content_tokens < 0:
                    continue
                content_token = content["content_token"][:, :num_content_tokens]
                trans_out = self.transformer.sample(condition_token=condition["condition_token"], condition_mask=condition.get("condition_mask", None),
                                                    condition_embed=condition.get("condition_embed_token", None), content_token=content_token,
                                                    filter_ratio=fr, temperature=temperature, return_att_weight=False, return_logits=return_logits,
                                                    content_logits=content.get("content_logits", None), sample_type=sample_type, batch_size=B, **kwargs)
                content_samples["cond1_cont{}_fr{}_image".format(cr, fr)] = self.decode_to_img(trans_out["content_token"], zshape)
                if return_logits:
                    content_samples["logits"] = trans_out["logits"]
        self.transformer.truncation, self.transformer.resample_rate = prev_trunc, prev_rate
        self.train()
        output = {"condition": batch.get(self.condition_info["key"])}
        output.update(content_samples)
        return output

    def parameters(self, recurse=True, name=None):
        """Reference override (dalle_spec.py:51-62): `name` selects sub-modules ('transformer') and forwards to their own parameters(name=...)."""
        if name is None or name == "none":
            return super().parameters(recurse=recurse)
        params = []
        for n in name.split("+"):
            sub = getattr(self, n)
            try:
                params += sub.parameters(recurse=recurse, name=n)
            except TypeError:  # plain nn.Module.parameters has no `name`
                params += list(sub.parameters(recurse=recurse))
        return params

    @torch.no_grad()
    def get_tokens(self, spec):
        """mel (B, 1, 80, 848) -> (quant_z (B, E, 5, 53), token ids (B, 265) in the transformer's column-major order)  (reference :71-78)."""
        quant_z, _, info = self.content_codec.encode(spec)
        indices = self.first_stage_permuter(info[2].view(quant_z.shape[0], -1))
        self.zshape = quant_z.shape
        return quant_z, indices

    @torch.no_grad()
    def prepare_content(self, batch, with_mask=False):
        """Reference :107-126: mels under batch[content_info['key']] are tokenised by the frozen SpecVQGAN encoder (get_tokens); a batch may
        instead carry pre-tokenised grids under 'content_token' (B, 265) int64 in the transformer's (column-major) order."""
        if "content_token" in batch:
            return {"content_token": batch["content_token"].to(self.device)}
        quant_z, indices = self.get_tokens(batch[self.content_info["key"]].to(self.device))
        return {"content_token": indices, "content_quant": quant_z}

    @torch.no_grad()
    def prepare_input(self, batch):
        inp = self.prepare_condition(batch)
        inp.update(self.prepare_content(batch))
        return inp

    def forward(self, batch, name="none", **kwargs):
        """Training / validation step entry (reference :340-351): `Solver.step` calls model(batch=..., return_loss=True, step=...)."""
        return self.transformer(self.prepare_input(batch), **kwargs)
