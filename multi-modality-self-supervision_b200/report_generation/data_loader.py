"""Drop-in mirror of .../sc/data_loader.py for the report-generation fine-tune path: `truncate_tokens_pair` (:24-59) and
`Preprocess4Seq2seq` (:297-452).  Host integer work; consumes Python's module-level `random` generator in exactly the
reference's order (truncation draws, shuffle of the candidate positions, one draw for the forced final-[SEP] mask), so a
seeded run selects the same tokens (bit-exact parity target, tests/golden/finetune_preprocess.npz).

Besides the reference's 12-tuple the pipeline can emit the compact mask description the CUDA path consumes:
`compact_mask=True` replaces the [max_len, max_len] int64 mask (2 MB per sample at L = 512) by (mode, t_len).
"""
import random
from random import random as rand
from random import shuffle

import torch

from .._lib import MODE_BAR_FT, MODE_BIDIR, MODE_S2S_FT


def truncate_tokens_pair(tokens_a, tokens_b, max_len, max_len_a=0, max_len_b=0, trunc_seg=None, always_truncate_tail=False):
    n_a, n_b = [0, 0], [0, 0]
    while len(tokens_a) + len(tokens_b) > max_len:
        if max_len_a > 0 and len(tokens_a) > max_len_a:
            side, cnt = tokens_a, n_a
        elif max_len_b > 0 and len(tokens_b) > max_len_b:
            side, cnt = tokens_b, n_b
        elif trunc_seg:
            side, cnt = (tokens_a, n_a) if trunc_seg == "a" else (tokens_b, n_b)
        else:
            side, cnt = (tokens_a, n_a) if len(tokens_a) > len(tokens_b) else (tokens_b, n_b)
        if (not always_truncate_tail) and rand() < 0.5:
            del side[0]
            cnt[0] += 1
        else:
            side.pop()
            cnt[1] += 1
    return n_a, n_b


class Preprocess4Seq2seq:
    """Pre-processing of one (image, report) instance for the fine-tune step (tasks == 'report_generation')."""

    def __init__(self, args, max_pred, mask_prob, vocab_words, indexer, max_len, bar, block_mask=False, new_segment_ids=False,
                 truncate_config={}, mode=None, len_vis_input=None, local_rank=-1, load_vqa_set=False, compact_mask=False,
                 image_loader=None):
        assert mode in ("s2s", "bi", "bar")
        self.tasks = getattr(args, "tasks", "report_generation")
        if self.tasks != "report_generation":
            raise NotImplementedError("only tasks='report_generation' is on the B200 path (VQA is out of scope)")
        self.max_pred, self.mask_prob, self.vocab_words, self.indexer = max_pred, mask_prob, vocab_words, indexer
        self.max_len, self.bar, self.new_segment_ids, self.mode = max_len, bar, new_segment_ids, mode
        self.always_truncate_tail = truncate_config.get("always_truncate_tail", False)
        self.max_len_b = truncate_config.get("max_len_b", None)
        self.trunc_seg = truncate_config.get("trunc_seg", None)
        self.task_idx = 3 if mode == "s2s" else 0
        self.len_vis_input = len_vis_input
        self.compact_mask = compact_mask
        self.image_loader = image_loader
        self._tril = None if compact_mask else torch.tril(torch.ones((max_len, max_len), dtype=torch.long))

    def _load_image(self, img_path):
        if self.image_loader is not None:
            return self.image_loader(img_path)
        import torchvision.transforms as transforms
        from PIL import Image

        img = transforms.Grayscale(num_output_channels=3)(Image.open(img_path))
        if self.len_vis_input < 100:
            img = transforms.Resize(224)(img)
        img = transforms.ToTensor()(img)
        return transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])(img)

    def __call__(self, instance):
        img_path, tokens_b, target, ans_type, organ = instance
        tokens_b = list(tokens_b)
        tokens_a = ["[UNK]"] * self.len_vis_input
        truncate_tokens_pair(tokens_a, tokens_b, self.len_vis_input + self.max_len_b, max_len_b=self.max_len_b,
                             trunc_seg=self.trunc_seg, always_truncate_tail=self.always_truncate_tail)
        tokens = ["[CLS]"] + tokens_a + ["[SEP]"] + tokens_b + ["[SEP]"]
        A = len(tokens_a) + 2
        if self.new_segment_ids and self.mode == "s2s":
            segment_ids = [4] * A + [5] * (len(tokens_b) + 1)
        else:       # 'bi' with new_segment_ids and the default both use 0 / 1
            segment_ids = [0] * A + [1] * (len(tokens_b) + 1)
        n_pred = min(self.max_pred, max(1, int(round(len(tokens_b) * self.mask_prob))))
        cand_pos = [i for i, tk in enumerate(tokens) if i >= A and tk != "[CLS]"]
        shuffle(cand_pos)
        if random.random() > 0.5:           # the final [SEP] is force-masked half of the time
            masked_pos = cand_pos[:n_pred - 1]
            masked_pos.append(len(tokens) - 1)
        else:
            masked_pos = cand_pos[:n_pred]
        masked_tokens = [tokens[p] for p in masked_pos]
        for p in masked_pos:
            tokens[p] = "[MASK]"
        masked_weights = [1] * len(masked_tokens)
        input_ids = self.indexer(tokens)
        masked_ids = self.indexer(masked_tokens)
        n_pad = self.max_len - len(input_ids)
        input_ids.extend([0] * n_pad)
        segment_ids.extend([0] * n_pad)
        t_len = len(tokens_b) + 1
        mode_id = MODE_BAR_FT if self.bar else (MODE_S2S_FT if self.mode == "s2s" else MODE_BIDIR)
        if self.compact_mask:
            input_mask = torch.tensor([mode_id, t_len], dtype=torch.long)
        else:
            end = A + t_len
            input_mask = torch.zeros(self.max_len, self.max_len, dtype=torch.long)
            if self.bar:
                input_mask[:, :A] = 1
                input_mask[:A, :] = 1
                input_mask[A:end, A:end] = self._tril[:t_len, :t_len]
            elif self.mode == "s2s":
                input_mask[:, :A] = 1
                input_mask[A:end, A:end] = self._tril[:t_len, :t_len]
            else:
                input_mask[:, :len(tokens)] = 1
        if self.max_pred > n_pred:
            pad = self.max_pred - n_pred
            masked_ids.extend([0] * pad)
            masked_pos.extend([0] * pad)
            masked_weights.extend([0] * pad)
        img = self._load_image(img_path)
        vis_pe = torch.arange(2048, dtype=torch.float).unsqueeze(0).expand(len(tokens_a), 2048)
        zero = torch.tensor(0)
        return (input_ids, segment_ids, input_mask, masked_ids, masked_pos, masked_weights, self.task_idx, img, vis_pe, zero, zero, zero)
