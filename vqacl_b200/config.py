"""Model configuration of the VL-T5 hot path.

Mirrors what `TrainerBase.create_config` builds (VL-T5/src/trainer_base.py:57-89): `T5Config.from_pretrained('t5-base')`
plus the VL extras set from `param.py` flags. There is no hub access on the target boxes, so the t5-base
hyper-parameters are tabulated here.
"""

_T5_BASE = dict(
    vocab_size=32128, d_model=768, d_kv=64, d_ff=3072, num_layers=12, num_decoder_layers=12, num_heads=12,
    relative_attention_num_buckets=32, relative_attention_max_distance=128, dropout_rate=0.1,
    layer_norm_epsilon=1e-6, initializer_factor=1.0, feed_forward_proj="relu", pad_token_id=0, eos_token_id=1,
    decoder_start_token_id=0, tie_word_embeddings=True, is_encoder_decoder=True, use_cache=True,
)

_VL_EXTRAS = dict(
    feat_dim=2048, pos_dim=4, n_images=2,                          # trainer_base.py:71-73
    use_vis_order_embedding=True, use_vis_layer_norm=True, individual_vis_layer_norm=True,   # param.py:92-94
    share_vis_lang_layer_norm=False, classifier=False, losses="vqa",
    # SI prototype bank (modeling_t5_our.py:381-382; Question_type.py:16-24)
    proto_split_L=20, n_ques_classes=10, n_cate_classes=80,
)


class VLT5Config:
    """Attribute bag with the fields the reference reads from its T5Config."""

    def __init__(self, **kw):
        for k, v in {**_T5_BASE, **_VL_EXTRAS}.items():
            setattr(self, k, v)
        for k, v in kw.items():
            setattr(self, k, v)
        # the reference sets all four from --dropout (trainer_base.py:77-80)
        if "dropout_rate" in kw:
            self.dropout = self.attention_dropout = self.activation_dropout = self.dropout_rate

    @classmethod
    def from_pretrained(cls, name="t5-base", **kw):
        if name not in ("t5-base", "t5_base"):
            raise ValueError(f"vqacl_b200 kernels are specialised for the t5-base geometry (got backbone {name!r}); "
                             "the reference's VQACL scripts all use --backbone t5-base (VL-T5/scripts/VQACL_train.sh:22)")
        return cls(**kw)

    @classmethod
    def from_args(cls, args):
        """`TrainerBase.create_config` for a parsed `param.py` namespace (trainer_base.py:57-89)."""
        cfg = cls.from_pretrained(getattr(args, "backbone", "t5-base"))
        for src, dst in (("feat_dim", "feat_dim"), ("pos_dim", "pos_dim"),
                         ("use_vis_order_embedding", "use_vis_order_embedding"),
                         ("use_vis_layer_norm", "use_vis_layer_norm"),
                         ("individual_vis_layer_norm", "individual_vis_layer_norm"),
                         ("share_vis_lang_layer_norm", "share_vis_lang_layer_norm"),
                         ("classifier", "classifier"), ("losses", "losses")):
            if hasattr(args, src):
                setattr(cfg, dst, getattr(args, src))
        if hasattr(args, "dropout"):
            cfg.dropout_rate = cfg.dropout = cfg.attention_dropout = cfg.activation_dropout = args.dropout
        return cfg

    def check_supported(self):
        """The CUDA path implements the configuration every VQACL script uses; anything else fails loudly."""
        bad = []
        if self.feed_forward_proj != "relu":
            bad.append("feed_forward_proj must be 'relu' (t5-base; SURVEY.md H4)")
        if not (self.use_vis_order_embedding and self.use_vis_layer_norm and self.individual_vis_layer_norm):
            bad.append("use_vis_order_embedding/use_vis_layer_norm/individual_vis_layer_norm must all be True (param.py:92-94)")
        if self.share_vis_lang_layer_norm:
            bad.append("share_vis_lang_layer_norm is not supported")
        if self.classifier:
            bad.append("classifier=True references a non-existent answer_head in the reference (vqa_model.py:81-108)")
        if self.pos_dim != 4:
            bad.append("pos_dim must be 4")
        if bad:
            raise ValueError("unsupported VL-T5 configuration for the B200 path: " + "; ".join(bad))

    def to_dict(self):
        return dict(self.__dict__)
