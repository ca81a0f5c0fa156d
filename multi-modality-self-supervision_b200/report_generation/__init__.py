"""Report-generation fine-tune step on the B200 engine: drop-in mirror of
/root/reference/Downstream_task/report_generation_and_vqa/sc/{data_loader.py, pytorch_pretrained_bert/model.py,
pytorch_pretrained_bert/optimization.py} for the training step of finetune.py:421-470."""
from .data_loader import Preprocess4Seq2seq, truncate_tokens_pair  # noqa: F401
from .model import BertForPreTrainingLossMask, pretrain_to_finetune_key  # noqa: F401
from .optimization import SCHEDULES, BertAdam, warmup_constant, warmup_cosine, warmup_linear  # noqa: F401
