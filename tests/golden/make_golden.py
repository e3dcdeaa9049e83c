"""
Generates tests/golden/*.npz + manifest.json by IMPORTING the reference
(/root/reference/vqvae/layers.py, evonorm.py) and running its own modules on CPU fp32.
Run only where /root/reference exists (the build container):

    python tests/golden/make_golden.py

The fixtures pin oracle/ (tests/test_oracle_golden.py) and, through it, the CUDA path.
Weights/inputs are NOT stored: they are regenerated from frozen numpy streams
(tests/golden/common.py); fixtures hold the reference's outputs and the reference's
state_dict key/shape listing (checkpoint-compatibility pin).
"""
import ast
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from common import golden_state_dict, spec_of, portable_randn, portable_volume  # noqa: E402
import vqvae.layers as ref  # noqa: E402  (the reference)
import vqvae.evonorm as ref_evo  # noqa: E402

torch.set_num_threads(1)
manifest = {"torch": torch.__version__, "cpu_capability": torch.backends.cpu.get_cpu_capability(), "cases": {}}


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()})


# ---------------------------------------------------------------- quantizer ------------
QUANT_CASES = {
    # name: (B, D, K, (H, W, Z))
    "q_d2_k128": (1, 2, 128, (8, 8, 8)),
    "q_d8_k256": (2, 8, 256, (4, 4, 4)),
    "q_d32_k512": (1, 32, 512, (4, 4, 2)),
    "q_d1_k8": (1, 1, 8, (4, 4, 4)),
    "q_d3_k7": (1, 3, 7, (5, 3, 2)),
    "q_d5_k33": (1, 5, 33, (3, 4, 5)),
    "q_d6_k40": (2, 6, 40, (3, 3, 3)),
    "q_d64_k96": (1, 64, 96, (2, 2, 2)),
}


def quantizer_case(name, B, D, K, sp, seed):
    x = portable_randn((B, D) + sp, seed)
    q = ref.Quantizer(num_embeddings=K, embedding_dim=D, commitment_cost=0.1)
    embed = portable_randn((K, D), seed + 1)
    out = {}
    # eval forward
    q.eval()
    q.embed.copy_(embed); q.embed_avg.copy_(embed); q.cluster_size.zero_(); q.first_pass.fill_(0)
    loss, quant, idx = q(x)
    flat = x.permute(0, 2, 3, 4, 1).reshape(-1, D)
    out.update(eval_loss=loss, eval_quantized=quant, eval_idx=idx,
               eval_cdist=torch.cdist(flat, embed, compute_mode="donot_use_mm_for_euclid_dist"))
    # first training forward (data-dependent init + EMA), then a second one
    q.train()
    q.embed.copy_(embed); q.embed_avg.copy_(embed); q.cluster_size.zero_(); q.first_pass.fill_(1)
    loss1, quant1, idx1 = q(x)
    out.update(t1_loss=loss1, t1_quantized=quant1, t1_idx=idx1, t1_embed=q.embed.clone(),
               t1_embed_avg=q.embed_avg.clone(), t1_cluster_size=q.cluster_size.clone(),
               t1_first_pass=q.first_pass.clone())
    x2 = portable_randn((B, D) + sp, seed + 2)
    loss2, quant2, idx2 = q(x2)
    out.update(t2_loss=loss2, t2_idx=idx2, t2_embed=q.embed.clone(), t2_embed_avg=q.embed_avg.clone(),
               t2_cluster_size=q.cluster_size.clone())
    # straight-through backward of the eval forward (layers.py:716-720)
    xg = x.clone().requires_grad_(True)
    q.eval(); q.embed.copy_(embed)
    l, qq, _ = q(xg)
    gq = portable_randn(qq.shape, seed + 3)
    (l * 1.7 + (qq * gq).sum()).backward()
    out.update(bwd_grad_q=gq, bwd_grad_x=xg.grad)
    save(name, **out)
    manifest["cases"][name] = {"kind": "quantizer", "B": B, "D": D, "K": K, "spatial": list(sp), "seed": seed}


# tie / sqrt-merge cases: duplicated codewords and inputs sitting exactly on codewords
def quantizer_ties():
    D, K = 2, 16
    embed = portable_randn((K, D), 901)
    embed[5] = embed[2]; embed[9] = embed[2]; embed[11] = embed[3]
    x = torch.cat([embed, embed[[2, 3, 5, 9]] + 1e-7, embed * 1.0000001]).T.reshape(1, D, 6, 6, 1).contiguous()
    q = ref.Quantizer(K, D, 0.1).eval()
    q.embed.copy_(embed); q.first_pass.fill_(0)
    _, _, idx = q(x)
    # near-equal distances whose sqrt rounds to the same float: lower index must win
    D2, K2 = 4, 64
    base = portable_randn((1, D2), 902)
    e2 = base + portable_randn((K2, D2), 903) * 1e-3 + 3.0
    x2 = base.repeat(32, 1).T.reshape(1, D2, 4, 4, 2).contiguous()
    q2 = ref.Quantizer(K2, D2, 0.1).eval()
    q2.embed.copy_(e2); q2.first_pass.fill_(0)
    _, _, idx2 = q2(x2)
    save("q_ties", embed=embed, x=x, idx=idx, embed2=e2, x2=x2, idx2=idx2)
    manifest["cases"]["q_ties"] = {"kind": "quantizer_ties"}


# ---------------------------------------------------------------- blocks ---------------
BLOCK_CASES = {
    # name: (class, cin, cout, mode, input shape)
    "preact_same_c6": ("PreActFixupResBlock", 6, 6, "same", (2, 6, 6, 4, 8)),
    "preact_same_c2": ("PreActFixupResBlock", 2, 2, "same", (1, 2, 4, 6, 8)),
    "preact_same_skip_18_2": ("PreActFixupResBlock", 18, 2, "same", (1, 18, 4, 6, 4)),
    "preact_out_4_1": ("PreActFixupResBlock", 4, 1, "out", (1, 4, 4, 4, 4)),
    "preact_down_4_8": ("PreActFixupResBlock", 4, 8, "down", (2, 4, 8, 4, 12)),
    "preact_up_8_4": ("PreActFixupResBlock", 8, 4, "up", (2, 8, 3, 4, 5)),
    "preact_up_18_8": ("PreActFixupResBlock", 18, 8, "up", (1, 18, 2, 3, 2)),
    "preact_same_thin": ("PreActFixupResBlock", 4, 4, "same", (1, 4, 2, 1, 8)),
    "fixup_same_4": ("FixupResBlock", 4, 4, "same", (2, 4, 4, 6, 4)),
    "fixup_out_4_2": ("FixupResBlock", 4, 2, "out", (1, 4, 4, 4, 4)),
    "fixup_down_4_8": ("FixupResBlock", 4, 8, "down", (1, 4, 8, 4, 4)),
    "fixup_up_8_4": ("FixupResBlock", 8, 4, "up", (1, 8, 3, 4, 2)),
    "evonorm_same_16": ("EvonormResBlock", 16, 16, "same", (1, 16, 4, 4, 4)),
    "evonorm_down_8_16": ("EvonormResBlock", 8, 16, "down", (1, 8, 4, 8, 4)),
    "evonorm_up_16_8": ("EvonormResBlock", 16, 8, "up", (1, 16, 2, 3, 4)),
}


def block_case(name, cls, cin, cout, mode, shape, seed):
    torch.manual_seed(0)
    m = getattr(ref, cls)(cin, cout, mode).eval()
    spec = spec_of(m.state_dict())
    m.load_state_dict(golden_state_dict(spec, seed))
    x = portable_randn(shape, seed + 7)
    with torch.no_grad():
        y = m(x)
    save(name, y=y)
    manifest["cases"][name] = {"kind": "block", "cls": cls, "cin": cin, "cout": cout, "mode": mode,
                               "shape": list(shape), "seed": seed, "spec": spec}


# ---------------------------------------------------------------- models ---------------
MODEL_CASES = {
    "tiny2_preact": (dict(n_bottleneck_blocks=2, num_embeddings=[16, 32], n_pre_quantization_blocks=2,
                          n_post_quantization_blocks=2, n_post_upscale_blocks=1, n_post_downscale_blocks=1,
                          block_type="pre-activation", base_network_channels=4), (1, 1, 32, 48, 16)),
    "tiny3_preact": (dict(n_bottleneck_blocks=3, num_embeddings=[8, 16, 32], n_pre_quantization_blocks=1,
                          n_post_quantization_blocks=1, n_post_upscale_blocks=1, n_post_downscale_blocks=0,
                          block_type="pre-activation", base_network_channels=2), (1, 1, 64, 64, 128)),
    "tiny2_regular": (dict(n_bottleneck_blocks=2, num_embeddings=[16, 32], n_pre_quantization_blocks=1,
                           n_post_quantization_blocks=1, n_post_upscale_blocks=0, n_post_downscale_blocks=1,
                           block_type="regular", base_network_channels=4), (1, 1, 16, 32, 16)),
    "tiny2_evonorm": (dict(n_bottleneck_blocks=2, num_embeddings=[16, 32], n_pre_quantization_blocks=1,
                           n_post_quantization_blocks=1, n_post_upscale_blocks=1, n_post_downscale_blocks=0,
                           block_type="evonorm", base_network_channels=4), (1, 1, 16, 16, 32)),
}
_RESBLOCKS = {"regular": "FixupResBlock", "pre-activation": "PreActFixupResBlock", "evonorm": "EvonormResBlock"}


class RefModel(torch.nn.Module):
    """The ten lines of glue of vqvae/model.py:45-65,79-89 (model.py itself needs
    pytorch_lightning, which is not installed): same ctor kwargs, same forward."""

    def __init__(self, cfg):
        super().__init__()
        rb = getattr(ref, _RESBLOCKS[cfg["block_type"]])
        self.encoder = ref.Encoder2(
            in_channels=1, base_network_channels=cfg["base_network_channels"], n_enc=cfg["n_bottleneck_blocks"],
            n_down_per_enc=2, n_pre_q_blocks=cfg["n_pre_quantization_blocks"],
            n_post_downscale_blocks=cfg["n_post_downscale_blocks"], n_post_upscale_blocks=cfg["n_post_upscale_blocks"],
            num_embeddings=cfg["num_embeddings"], resblock=rb)
        self.decoder = ref.Decoder(
            out_channels=1, base_network_channels=cfg["base_network_channels"], n_enc=cfg["n_bottleneck_blocks"],
            n_up_per_enc=2, n_post_q_blocks=cfg["n_post_quantization_blocks"],
            n_post_upscale_blocks=cfg["n_post_upscale_blocks"], resblock=rb)

    def forward(self, data):
        commitment_loss, quantizations, encoding_idx = zip(*self.encoder(data))
        return self.decoder(quantizations), (commitment_loss, quantizations, encoding_idx)


def model_case(name, cfg, shape, seed):
    torch.manual_seed(0)
    m = RefModel(cfg)
    spec = spec_of(m.state_dict())
    sd = golden_state_dict(spec, seed)
    x = portable_volume(shape, seed + 11)
    out = {}
    latents = {}

    def hook_for(i):
        def hook(mod, inp):
            latents[i] = inp[0].detach().clone()
        return hook
    for i, qz in enumerate(m.encoder.quantize):
        qz.register_forward_pre_hook(hook_for(i))

    m.load_state_dict(sd); m.eval()
    with torch.no_grad():
        dec, (losses, quants, idxs) = m(x)
    full = dec.numel() <= 40000
    out["eval_decoded"] = dec if full else dec[..., ::4, ::4, ::4]
    out["eval_decoded_sum"] = dec.double().sum(); out["eval_decoded_abssum"] = dec.double().abs().sum()
    for i in range(len(losses)):
        out[f"eval_loss_{i}"] = losses[i]; out[f"eval_idx_{i}"] = idxs[i]
        out[f"eval_quantized_{i}"] = quants[i]; out[f"eval_latent_{i}"] = latents[i]
    # training-mode forward from fresh buffers (first_pass = 1): data-dependent init + EMA
    sd_t = {k: v.clone() for k, v in sd.items()}
    for k in sd_t:
        if k.endswith("first_pass"):
            sd_t[k] = torch.ones_like(sd_t[k])
        if k.endswith("cluster_size"):
            sd_t[k] = torch.zeros_like(sd_t[k])
    m.load_state_dict(sd_t); m.train()
    with torch.no_grad():
        dec_t, (losses_t, _, idxs_t) = m(x)
    out["train_decoded_sum"] = dec_t.double().sum()
    for i in range(len(losses_t)):
        out[f"train_loss_{i}"] = losses_t[i]; out[f"train_idx_{i}"] = idxs_t[i]
        for b in ("embed", "embed_avg", "cluster_size", "first_pass"):
            out[f"train_{b}_{i}"] = getattr(m.encoder.quantize[i], b).clone()
    save(name, **out)
    manifest["cases"][name] = {"kind": "model", "cfg": cfg, "shape": list(shape), "seed": seed, "spec": spec,
                               "decoded_full": bool(full)}


# ---------------------------------------------------------------- misc -----------------
def cylinder_and_evonorm():
    # run the reference's mask builder without importing its module (needs nrrd/monai):
    src = open("/root/reference/utils/load_nrrd_dataset.py").read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ExtractCenterCylinder")
    ns = {"torch": torch, "np": np, "Union": __import__("typing").Union, "Tuple": __import__("typing").Tuple}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), "ref_cylinder", "exec"), ns)
    masks = {f"mask_{h}_{w}": ns["ExtractCenterCylinder"].create_cylinder_xy_mask((h, w)) for h, w in ((8, 8), (16, 12), (7, 9))}
    # EvoNorm forward/backward (evonorm.py:59-76 and the custom backward :36-47)
    torch.manual_seed(0)
    en = ref_evo.EvoNorm3DS0(16)
    spec = spec_of(en.state_dict())
    en.load_state_dict(golden_state_dict(spec, 77))
    x = portable_randn((1, 16, 4, 6, 8), 78).requires_grad_(True)
    y = en(x)
    g = portable_randn(y.shape, 79)
    y.backward(g)
    save("misc", evonorm_y=y, evonorm_gx=x.grad, evonorm_gv=en.v.grad, evonorm_ggamma=en.gamma.grad,
         evonorm_gbeta=en.beta.grad, **masks)
    manifest["cases"]["misc"] = {"kind": "misc", "evonorm_spec": spec}


if __name__ == "__main__":
    for i, (name, (B, D, K, sp)) in enumerate(QUANT_CASES.items()):
        quantizer_case(name, B, D, K, sp, 100 + 10 * i)
    quantizer_ties()
    for i, (name, (cls, cin, cout, mode, shape)) in enumerate(BLOCK_CASES.items()):
        block_case(name, cls, cin, cout, mode, shape, 300 + i)
    for i, (name, (cfg, shape)) in enumerate(MODEL_CASES.items()):
        model_case(name, cfg, shape, 500 + i)
    cylinder_and_evonorm()
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("golden fixtures written:", sorted(manifest["cases"]))
