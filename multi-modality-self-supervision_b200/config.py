"""Offline stand-in for transformers.BertConfig (the reference calls BertConfig.from_pretrained over the network,
models/train_origin.py:36-47).  Any object with the same attribute names (e.g. a real transformers.BertConfig) is
accepted wherever this class is."""
import json
import os

_PRESETS = {
    "bert-base-uncased": dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                              intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2),
    "google/bert_uncased_L-4_H-512_A-8": dict(vocab_size=30522, hidden_size=512, num_hidden_layers=4, num_attention_heads=8,
                                              intermediate_size=2048, max_position_embeddings=512, type_vocab_size=2),
}


class BertConfig:
    model_type = "bert"

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
                 type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0, **kw):
        self.vocab_size, self.hidden_size, self.num_hidden_layers = vocab_size, hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size, self.hidden_act = num_attention_heads, intermediate_size, hidden_act
        self.hidden_dropout_prob, self.attention_probs_dropout_prob = hidden_dropout_prob, attention_probs_dropout_prob
        self.max_position_embeddings, self.type_vocab_size = max_position_embeddings, type_vocab_size
        self.initializer_range, self.layer_norm_eps, self.pad_token_id = initializer_range, layer_norm_eps, pad_token_id
        for k, v in kw.items():
            setattr(self, k, v)

    @classmethod
    def from_pretrained(cls, name_or_path, **kw):
        path = os.path.join(str(name_or_path), "config.json")
        if os.path.isfile(path):
            with open(path) as f:
                d = json.load(f)
            d.update(kw)
            return cls(**{k: v for k, v in d.items() if k not in ("model_type", "architectures")})
        if name_or_path in _PRESETS:
            return cls(**dict(_PRESETS[name_or_path], **kw))
        raise ValueError("unknown config %r (offline presets: %s)" % (name_or_path, sorted(_PRESETS)))

    def to_dict(self):
        d = {k: v for k, v in self.__dict__.items() if isinstance(v, (int, float, str, bool, type(None), list, dict))}
        d["model_type"] = self.model_type
        return d

    def save_pretrained(self, path):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)


AutoConfig = BertConfig
