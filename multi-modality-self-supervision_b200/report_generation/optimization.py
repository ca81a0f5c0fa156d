"""Drop-in mirror of .../sc/pytorch_pretrained_bert/optimization.py (BertAdam :58-182, schedules :32-55).

`BertAdam.step()` is one launch pair of libmedvill_sm100 (mv_bert_adam_step: per-tensor gradient norms, then the fused
clip + moments + decoupled weight decay + update + zero_grad + bf16 shadow refresh over the whole arena) instead of a
Python loop over ~200 tensors x 10 kernels.  Semantics kept: clip_grad_norm_(p, max_grad_norm) PER PARAMETER, no bias
correction, update += weight_decay * p for names without 'bias' / 'LayerNorm', lr * schedule(step / t_total, warmup)
with the 0-based per-parameter step counter, parameters without a gradient (pooler) untouched.
"""
import math

from .._lib import MedvillError


def warmup_cosine(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return 1.0


def warmup_linear(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return max((x - 1.0) / (warmup - 1.0), 0)


SCHEDULES = {"warmup_cosine": warmup_cosine, "warmup_constant": warmup_constant, "warmup_linear": warmup_linear}


class BertAdam:
    """BertAdam(optimizer_grouped_parameters, lr=, warmup=, schedule=, t_total=) as built at finetune.py:383-395.

    The parameter groups are accepted for call compatibility; which tensors decay is decided by the engine from the
    parameter names exactly as finetune.py's `no_decay` list does, and every group must use the library's two decay
    values (the first group's `weight_decay`, and 0).  The engine is located through the parameters themselves (they
    are views of its arena) or passed as `engine=` / `model=`."""

    def __init__(self, params, lr=None, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6,
                 weight_decay=0.01, max_grad_norm=1.0, engine=None, model=None):
        if lr is None or lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if schedule not in SCHEDULES:
            raise ValueError("Invalid schedule parameter: {}".format(schedule))
        if not 0.0 <= warmup < 1.0 and not warmup == -1:
            raise ValueError("Invalid warmup: {} - should be in [0.0, 1.0[ or -1".format(warmup))
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        decays = sorted({float(g.get("weight_decay", weight_decay)) for g in groups})
        nonzero = [d for d in decays if d > 0.0]
        if len(nonzero) > 1:
            raise MedvillError("BertAdam on the B200 engine supports one non-zero weight_decay value, got %s" % decays)
        self.weight_decay = nonzero[0] if nonzero else 0.0
        self.param_groups = [dict(g, lr=g.get("lr", lr), schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                                  max_grad_norm=max_grad_norm) for g in groups]
        self.lr, self.warmup, self.t_total, self.schedule = lr, warmup, t_total, schedule
        self.b1, self.b2, self.e, self.max_grad_norm = b1, b2, e, max_grad_norm
        self._engine, self._model = engine, model
        self.state = {"step": 0}

    def _find_engine(self):
        if self._engine is not None:
            return self._engine
        if self._model is not None:
            return self._model.engine()
        from ..engine import find_engine

        for g in self.param_groups:
            for p in g["params"]:
                eng = find_engine(p)
                if eng is not None:
                    return eng
        raise MedvillError("BertAdam: no parameter belongs to a live B200 engine (move the model to a CUDA device and run a "
                           "forward first, or pass model=)")

    def scheduled_lr(self):
        if self.t_total != -1:
            return self.lr * SCHEDULES[self.schedule](self.state["step"] / self.t_total, self.warmup)
        return self.lr

    def get_lr(self):
        return [self.scheduled_lr()] if self.state["step"] > 0 else [0]

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self._find_engine().bert_adam_step(self.scheduled_lr(), betas=(self.b1, self.b2), eps=self.e, weight_decay=self.weight_decay,
                                           max_grad_norm=self.max_grad_norm)
        self.state["step"] += 1
        return loss

    def zero_grad(self):
        """fused into step() (mv_bert_adam_step writes zeros behind the gradients it consumed)"""

    def state_dict(self):
        """What finetune.py:484-486 writes to optim.{epoch}.bin: the schedule position plus, when the engine exists, the Adam
        moments keyed by parameter name (engine.optimizer_state_dict)."""
        sd = {"state": dict(self.state), "lr": self.lr, "warmup": self.warmup, "t_total": self.t_total, "schedule": self.schedule}
        try:
            sd["moments"] = self._find_engine().optimizer_state_dict()
        except MedvillError:
            pass                                   # no engine yet: nothing has been stepped, the moments are zero
        return sd

    def load_state_dict(self, sd):
        """finetune.py:396-402 (recover_step): restores the schedule position and the moments."""
        self.state = dict(sd.get("state", {"step": 0}))
        if "moments" in sd:
            self._find_engine().load_optimizer_state_dict(sd["moments"])
