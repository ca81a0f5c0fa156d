"""`CXRBERT_Trainer` — drop-in mirror of /root/reference/models/train_origin.py:19-266.

Same constructor (`args, train_dataloader, test_dataloader=None`), `.train(epoch)`, `.save(epoch, file_path)` and the
same step semantics (joint MLM + ITM cross-entropy, `zero_grad -> backward -> AdamW(lr).step`, ITM / MLM accuracy
accounting, `save_pretrained` per epoch), but every step is one fused call into libmedvill_sm100
(`CXRBERT.pretrain_step`): no [B, L, V] logits, no `.item()` per loss, no additive mask tensor.

Multi-GPU: the reference wraps the model in single-process `nn.DataParallel` (train_origin.py:53-55).  Here it is one
process per GPU (torchrun); when `torch.distributed` is initialised the trainer hands rank 0's NCCL unique id to the
engine, which all-reduces gradient buckets on its own stream while backward is still running.
"""
import os

import numpy as np
import torch

from ..config import AutoConfig, BertConfig
from ..data.prefetch import DevicePrefetcher
from .cxrbert_origin import CXRBERT

try:  # logging only; the reference requires wandb, here it is optional
    import wandb
except Exception:  # pragma: no cover
    wandb = None

try:
    import tqdm
except Exception:  # pragma: no cover
    tqdm = None


def _log(payload, step):
    if wandb is not None and getattr(wandb, "run", None) is not None:
        wandb.log(payload, step=step)


class CXRBERT_Trainer():
    def __init__(self, args, train_dataloader, test_dataloader=None):
        self.args = args
        if not (torch.cuda.is_available() and args.with_cuda):
            raise RuntimeError("CXRBERT_Trainer needs a CUDA (sm_100a) device: this implementation has no CPU path")
        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.rank = torch.distributed.get_rank() if self.dist else 0
        self.world = torch.distributed.get_world_size() if self.dist else 1
        local = int(os.environ.get("LOCAL_RANK", 0)) if self.dist else torch.cuda.current_device()
        self.device = torch.device("cuda", local)
        torch.cuda.set_device(self.device)
        print('Current cuda device ', torch.cuda.current_device())

        if args.weight_load:
            config = AutoConfig.from_pretrained(args.pre_trained_model_path)
            state = torch.load(os.path.join(args.pre_trained_model_path, 'pytorch_model.bin'), map_location="cpu")
            self.model = CXRBERT.from_pretrained(args.pre_trained_model_path, state_dict=state, config=config, args=args).to(self.device)
            print('training restart with mid epoch')
        else:
            name = {"bert-small-scratch": "google/bert_uncased_L-4_H-512_A-8", "bert-base-scratch": "bert-base-uncased"}.get(
                args.bert_model, args.bert_model)
            config = getattr(args, "bert_config", None) or BertConfig.from_pretrained(name)
            self.model = CXRBERT(config, args).to(self.device)

        eng = self.model.engine()
        if args.weight_load and self.model.load_optimizer(args.pre_trained_model_path):
            print('optimizer state restored (step %d)' % eng.step_count)
        if self.world > 1:
            self.model.init_distributed()
            print("Using %d GPUS for BERT" % self.world)

        self.train_data = train_dataloader
        self.test_data = test_dataloader
        self.lr = args.lr               # AdamW(self.model.parameters(), lr=args.lr): every other Adam flag is unused
        self.log_freq = args.log_freq
        self.step_cnt = 0
        print("Total Parameters:", sum(p.nelement() for p in self.model.parameters()))

    def _unpack(self, data):
        cls_tok, input_ids, txt_labels, attn_masks, img, segment, is_aligned, sep_tok, itm_prob = data
        mode = t_len = None
        if attn_masks.dim() == 2 and attn_masks.shape[-1] == 2 and attn_masks.shape[-1] != input_ids.shape[-1]:
            mode, t_len = attn_masks[:, 0].to(torch.uint8), attn_masks[:, 1].to(torch.int32)   # compact (mode, t_len) form
        return cls_tok, input_ids, txt_labels, attn_masks, img, segment, is_aligned, sep_tok, mode, t_len

    def _iterate(self, loader, epoch, train):
        # host -> device staging of step i+1 overlaps step i (two persistent device buffer sets); txt_labels stay on the
        # host: selecting the labelled rows is host integer work (replaces the blocking .to(device) calls of :95-104)
        it = enumerate(DevicePrefetcher(loader, self.device, host_indices=(2, 8)))
        if tqdm is not None and self.rank == 0:
            it = tqdm.tqdm(it, desc=f'EP_:{epoch}', total=len(loader), bar_format='{l_bar}{r_bar}')
        losses, mlm_losses, itm_losses = [], [], []
        itm_ok = itm_n = mlm_ok = mlm_n = 0
        pending = None        # training statistics are read back asynchronously and consumed one step late

        def account(out):
            nonlocal itm_ok, itm_n, mlm_ok, mlm_n
            losses.append(out["loss"]); mlm_losses.append(out["mlm_loss"]); itm_losses.append(out["itm_loss"])
            itm_ok += out["itm_correct"]; itm_n += out["batch"]
            mlm_ok += out["mlm_correct"]; mlm_n += out["n_labelled"]
        for i, data in it:
            cls_tok, input_ids, txt_labels, attn_masks, img, segment, is_aligned, sep_tok, mode, t_len = self._unpack(data)
            if train:
                nxt = self.model.pretrain_step(cls_tok, input_ids, txt_labels, attn_masks, img, segment, is_aligned, sep_tok,
                                               lr=self.lr, mode=mode, t_len=t_len, lazy=True)
                self.step_cnt += 1
                if pending is not None:
                    account(pending())
                pending = nxt
            else:
                account(self.model.eval_step(cls_tok, input_ids, txt_labels, attn_masks, img, segment, is_aligned, sep_tok,
                                             mode=mode, t_len=t_len))
        if pending is not None:
            account(pending())
        return dict(loss=float(np.mean(losses)) if losses else float("nan"), mlm_loss=float(np.mean(mlm_losses)) if losses else float("nan"),
                    itm_loss=float(np.mean(itm_losses)) if losses else float("nan"), itm_acc=100.0 * itm_ok / max(1, itm_n),
                    mlm_acc=100.0 * mlm_ok / max(1, mlm_n))

    def train(self, epoch):
        self.model.train()
        r = self._iterate(self.train_data, epoch, train=True)
        print("avg loss per epoch", r["loss"])
        print("avg itm acc per epoch", round(r["itm_acc"], 3))
        _log({"avg_loss": r["loss"], "avg_mlm_loss": r["mlm_loss"], "avg_itm_loss": r["itm_loss"], "itm_acc": r["itm_acc"],
              "mlm_acc": r["mlm_acc"]}, epoch)
        self.last_train = r
        if self.test_data is not None:
            self.model.eval()
            e = self._iterate(self.test_data, epoch, train=False)
            print("avg loss in testset", e["loss"])
            print("avg itm acc in testset", round(e["itm_acc"], 3))
            _log({"eval_avg_loss": e["loss"], "eval_mlm_loss": e["mlm_loss"], "eval_itm_loss": e["itm_loss"],
                  "eval_itm_acc": e["itm_acc"], "eval_mlm_acc": e["mlm_acc"]}, epoch)
            self.last_eval = e
        return r

    def save(self, epoch, file_path):
        if self.rank != 0:
            return
        save_path_per_ep = os.path.join(file_path, str(epoch))
        os.makedirs(save_path_per_ep, exist_ok=True)
        os.chmod(save_path_per_ep, 0o777)
        self.model.save_pretrained(save_path_per_ep)
        self.model.save_optimizer(save_path_per_ep)      # Adam moments + step: a restart continues instead of resetting Adam
        print(f'EP: {epoch} Model saved on {save_path_per_ep}')
        os.chmod(save_path_per_ep + '/pytorch_model.bin', 0o777)
