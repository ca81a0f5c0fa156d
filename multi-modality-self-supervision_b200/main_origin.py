"""MedViLL pre-training entry point — drop-in for /root/reference/main_origin.py (same flags, same defaults).

    python -m medvill_b200.main_origin --train_dataset Train_253.jsonl --test_dataset Valid_253.jsonl
    torchrun --nproc-per-node 8 -m medvill_b200.main_origin ...          (one process per GPU, NCCL)

Differences from the reference script, all deliberate: CUDA_VISIBLE_DEVICES is respected instead of being forced to
"5,7" (main_origin.py:6-7); the output directory is created when training starts, not at import; wandb is optional;
`--precision {bf16,fp32}` and `--compact_masks` are new (the second ships (mode, t_len) instead of a [L, L] mask).
"""
import argparse
import os
from datetime import datetime

import torch
from torch.utils.data import DataLoader

from .data.dataset_origin import CXRDataset
from .data.helper import get_transforms
from .models.train_origin import CXRBERT_Trainer
from .utils.utils import set_seed


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--train_dataset", type=str, default='/home/mimic-cxr/dataset/new_dset/Train_253.jsonl')
    p.add_argument("--test_dataset", type=str, default='/home/mimic-cxr/dataset/new_dset/Valid_253.jsonl')
    p.add_argument("--output_path", type=str, default=None, help="ex)path/to/save/model")
    p.add_argument("--log_freq", type=int, default=10)
    p.add_argument("--with_cuda", type=bool, default=True)
    p.add_argument("--cuda_devices", type=int, nargs='+', default=None)
    p.add_argument("--mlm_task", type=str, default=True)
    p.add_argument("--itm_task", type=str, default=True)
    p.add_argument('--attn_1d', type=bool, default=False)
    p.add_argument('--BAR_attn', default=True, type=bool)
    p.add_argument('--Mixed', default=False, type=bool)
    p.add_argument('--s2s_prob', default=1.0, type=float)
    p.add_argument('--bi_prob', default=0.0, type=float)
    p.add_argument('--disturbing_mask', default=False, type=bool)
    p.add_argument("--epochs", type=int, default=50)
    p.add_argument("--batch_size", type=int, default=36)
    p.add_argument("--num_workers", type=int, default=20)
    p.add_argument("--hidden_size", type=int, default=768, choices=[768, 512, 128])
    p.add_argument("--embedding_size", type=int, default=768, choices=[768, 512, 128])
    p.add_argument("--weight_load", type=bool, default=False)
    p.add_argument("--pre_trained_model_path", type=str, default='/home/cxr-bert/clinicalbert_vlp_re35_5')
    p.add_argument("--bert_model", type=str, default="bert-base-scratch")
    p.add_argument("--vocab_size", type=int, default=30522, choices=[30522, 30000, 28996])
    p.add_argument("--img_postion", default=True)
    p.add_argument("--seq_len", type=int, default=253, choices=[128, 253])
    p.add_argument("--max_seq_len", type=int, default=512)
    p.add_argument("--img_hidden_sz", type=int, default=2048)
    p.add_argument("--img_encoder", type=str, default='random-pixel', choices=['random-pixel', 'full-fiber', 'ViT'])
    p.add_argument("--img_channel", type=int, default=3, choices=[1, 3])
    p.add_argument("--num_image_embeds", type=int, default=180, choices=[36, 49, 180, 256])
    p.add_argument("--img_size", type=int, default=512, choices=[224, 512])
    p.add_argument("--img_embed_pool_type", type=str, default="max", choices=["max", "avg"])
    p.add_argument("--lr", type=float, default=1e-5)
    p.add_argument("--gradient_accumulation_steps", type=int, default=4)   # parsed, unused (as in the reference)
    p.add_argument("--warmup", type=float, default=0.1)
    p.add_argument("--seed", type=int, default=123)
    p.add_argument("--warmup_steps", type=int, default=0)
    p.add_argument("--dropout_prob", type=float, default=0.1)
    p.add_argument("--beta1", type=float, default=0.9)
    p.add_argument("--beta2", type=float, default=0.999)
    p.add_argument("--eps", type=float, default=1e-6)
    p.add_argument("--weight_decay", type=float, default=0.01)
    # new, B200-specific
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--compact_masks", action="store_true", help="ship (mode, t_len) instead of the [L, L] mask tensor")
    p.add_argument("--max_micro_batch", type=int, default=64)
    p.add_argument("--allow_random_trunk", action="store_true",
                   help="construct the frozen ResNet-50 with random weights when the ImageNet checkpoint is not in the torch-hub cache")
    p.add_argument("--resnet_weights", type=str, default=None, help="path of resnet50-0676ba61.pth (default: torch-hub cache)")
    return p


def train(args):
    if "RANK" in os.environ and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        torch.distributed.init_process_group("nccl")
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    try:
        import wandb

        if rank == 0:
            wandb.init(config=args, project='CXR-BERT')
    except Exception:
        pass
    set_seed(args.seed + rank)
    from transformers import AutoTokenizer, BertTokenizer

    name = {"bert-small-scratch": "google/bert_uncased_L-4_H-512_A-8", "bert-base-scratch": "bert-base-uncased"}.get(
        args.bert_model, args.bert_model)
    tok_cls = BertTokenizer if ("bert-base" in name or "google" in name) else AutoTokenizer
    tokenizer = tok_cls.from_pretrained(name, do_lower_case=True).tokenize
    transforms = get_transforms(args)
    print("Load Train dataset", args.train_dataset)
    train_dataset = CXRDataset(args.train_dataset, tokenizer, transforms, args)
    print("Load Test dataset", args.test_dataset)
    test_dataset = CXRDataset(args.test_dataset, tokenizer, transforms, args) if args.test_dataset is not None else None
    sampler = torch.utils.data.distributed.DistributedSampler(train_dataset) if torch.distributed.is_initialized() else None
    train_loader = DataLoader(train_dataset, batch_size=args.batch_size, num_workers=args.num_workers, shuffle=sampler is None,
                              sampler=sampler, pin_memory=True)
    test_loader = DataLoader(test_dataset, batch_size=args.batch_size, num_workers=args.num_workers, shuffle=False,
                             pin_memory=True) if test_dataset is not None else None
    trainer = CXRBERT_Trainer(args, train_dataloader=train_loader, test_dataloader=test_loader)
    if args.output_path is None:
        args.output_path = 'output/' + str(datetime.now())
    if rank == 0:
        os.makedirs(args.output_path, exist_ok=True)
    print("Training Start!")
    for epoch in range(args.epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        trainer.train(epoch)
        trainer.save(epoch, args.output_path)


if __name__ == '__main__':
    train(build_parser().parse_args())
