"""Caption file -> .npy mels + .wav clips with the B200 kernels: the flow of Diffsound/evaluation/generate_samples_batch.py
(`Diffusion.__init__` :45-86, `read_tsv` :121-139, `generate_sample` :141-187) on the drop-in classes.

    python tools/generate_samples.py --config /path/to/caps_text.yaml --ckpt model.pth --vocoder-ckpt best_netG.pt \\
        --captions val.csv --out samples/ [--truncation 0.85] [--fast 0] [--replicate 2] [--clip-ckpt ViT-B-32.pt] [--bpe vocab.txt.gz]

The YAML is the reference's own file: its `target:` strings are rewritten to this package (utils.misc.retarget_config).  Captions come from
a CSV with `file_name,caption` columns (the reference's tsv) -- one output family `{file}_mel_sample_{n}` per file name, `replicate`
samples per caption, written in the layout Codebook/evaluate.py reads (pipeline.save_clip).  --dry-run builds everything on the CPU and stops
before the first kernel (used by the CPU test suite)."""
import argparse
import csv
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from diffsound_b200 import pipeline  # noqa: E402
from diffsound_b200.utils.misc import instantiate_from_config, retarget_config  # noqa: E402
from diffsound_b200.vocoder.modules import Generator  # noqa: E402


def read_captions(path):
    """file_name -> [captions] in file order (generate_samples_batch.py:121-139)."""
    caps = {}
    with open(path, newline="") as f:
        for row in csv.DictReader(f):
            caps.setdefault(row["file_name"], []).append(row["caption"])
    return caps


def build(args):
    with open(args.config) as f:
        cfg = retarget_config(yaml.full_load(f)["model"])
    p = cfg["params"]
    p["content_codec_config"]["params"]["ckpt_path"] = args.codec_ckpt          # the YAML's absolute path is the authors' machine
    p["content_codec_config"]["params"]["lossconfig"] = None
    if p.get("condition_codec_config") is not None and args.bpe:
        p["condition_codec_config"]["params"]["tokenizer_config"]["params"]["bpe_path"] = args.bpe
    emb = p["diffusion_config"]["params"].get("condition_emb_config")
    if emb is not None and args.clip_ckpt:
        emb["params"]["clip_ckpt_path"] = args.clip_ckpt
    model = instantiate_from_config(cfg)
    if args.ckpt:
        ckpt = torch.load(args.ckpt, map_location="cpu")
        missing, unexpected = model.load_state_dict(ckpt["model"], strict=False)            # :74
        print(f"model: {len(missing)} missing / {len(unexpected)} unexpected keys")
        if not args.no_ema and "ema" in ckpt:
            model.get_ema_model().load_state_dict(ckpt["ema"], strict=False)                # :80-83
            print("using the EMA weights")
    vocoder = None
    if args.vocoder_ckpt:
        vocoder = Generator(80, 32, 3)                                                       # vocoder/logs/vggsound/args.yml
        vocoder.load_state_dict(torch.load(args.vocoder_ckpt, map_location="cpu"))
    return model, vocoder


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", required=True)
    ap.add_argument("--ckpt", default=None, help="Diffsound checkpoint ({'model':..., 'ema':...})")
    ap.add_argument("--codec-ckpt", default=None, help="SpecVQGAN Lightning checkpoint (only needed if --ckpt does not hold content_codec.*)")
    ap.add_argument("--vocoder-ckpt", default=None, help="MelGAN generator state_dict (Diffsound/vocoder/logs/vggsound/best_netG.pt)")
    ap.add_argument("--clip-ckpt", default=None, help="OpenAI CLIP ViT-B/32 weights (state_dict or TorchScript archive)")
    ap.add_argument("--bpe", default=None, help="bpe_simple_vocab_16e6.txt.gz (default: $DIFFSOUND_BPE_VOCAB or the reference checkout)")
    ap.add_argument("--captions", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--truncation", type=float, default=0.85)
    ap.add_argument("--fast", type=int, default=0, help="skip-step sampler: 'r,fast{N-1}' as in the reference")
    ap.add_argument("--replicate", type=int, default=2)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--no-ema", action="store_true")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args(argv)
    sample_type = f"top{a.truncation}r" + (f",fast{a.fast - 1}" if a.fast else "")
    model, vocoder = build(a)
    caps = read_captions(a.captions)
    print(f"{len(caps)} files, {sum(len(v) for v in caps.values())} captions, sample_type {sample_type}")
    if a.dry_run:
        return model, vocoder, caps, sample_type
    model = model.cuda().eval()
    vocoder = vocoder.cuda().eval() if vocoder is not None else None
    torch.manual_seed(a.seed)
    for name, texts in caps.items():
        out = model.generate_content(batch={"text": texts, "image": None}, filter_ratio=0, replicate=a.replicate, content_ratio=1, sample_type=sample_type)
        mel = out["content"]
        wav = vocoder((mel[:, 0] + 1) / 2) if vocoder is not None else None                 # batched; the reference vocodes one clip at a time
        base = name.split(".")[0]
        for n in range(mel.shape[0]):
            pipeline.save_clip(a.out, base, n, mel[n], None if wav is None else wav[n])
    return 0


if __name__ == "__main__":
    main()
