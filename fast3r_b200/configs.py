"""Model hyper-parameter dicts (the three plain dicts Fast3R.__init__ consumes, SURVEY.md §5 'Config')."""


def vit_large_args(attn_implementation: str = "flash_attention"):
    """ViT-L/512: configs/model/fast3r.yaml:50-88 + configs/experiment/super_long_training/
    super_long_training.yaml:52-66, with the overrides every inference caller applies
    (fast3r/utils/checkpoint_utils.py:37-38): PatchEmbedDust3R, landscape_only=False."""
    enc = dict(encoder_type="croco", img_size=512, patch_size=16, patch_embed_cls="PatchEmbedDust3R",
               embed_dim=1024, num_heads=16, depth=24, mlp_ratio=4, pos_embed="RoPE100",
               attn_implementation=attn_implementation)
    dec = dict(decoder_type="fast3r", random_image_idx_embedding=True, enc_embed_dim=1024, embed_dim=1024,
               num_heads=16, depth=24, mlp_ratio=4.0, qkv_bias=True, drop=0.0, attn_drop=0.0,
               attn_implementation=attn_implementation)
    head = dict(head_type="dpt", output_mode="pts3d", landscape_only=False,
                depth_mode=["exp", float("-inf"), float("inf")], conf_mode=["exp", 1, float("inf")],
                patch_size=16, with_local_head=True)
    return enc, dec, head


def tiny_args(attn_implementation: str = "flash_attention", dec_depth: int = 12):
    """D=128, 2 heads (head_dim 64), encoder depth 2, decoder depth 12 — the fast parity-test model."""
    enc, dec, head = vit_large_args(attn_implementation)
    enc.update(embed_dim=128, num_heads=2, depth=2)
    dec.update(enc_embed_dim=128, embed_dim=128, num_heads=2, depth=dec_depth)
    return enc, dec, head
