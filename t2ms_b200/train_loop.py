"""Host-side semantics of the reference training driver around the fused step:

* ``collate_by_length``  — datafactory/dataloader.py:115-133 (``custom_collate_fn``): a mixed batch of
  ``((text, x, embedding), dataset_idx)`` items becomes up to three length-grouped sub-batches, in dataset order;
* ``one_cycle_lr``       — the ``OneCycleLR(max_lr=1e-4, total_steps=len(dataloader)*epochs)`` schedule of
  train.py:38 (cosine, pct_start 0.3, div_factor 25, final_div_factor 1e4), as a pure function of the step;
* ``optimizer_state_dict`` / ``load_optimizer_state_dict`` — the fused AdamW state in the exact
  ``torch.optim.AdamW.state_dict()`` format train.py:93-94 saves and train.py:42-47 resumes from;
* ``fit``                — the mix-train loop of train.py:52-95: one optimizer step per sub-batch, the scheduler
  stepped once per dataloader batch (:90), ``loss_list`` appended per step (:86), a checkpoint
  ``{model, optimizer, epoch, loss_list}`` every 1000 epochs and at the end (:92-95).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Iterable, List, Optional, Sequence

import torch

from .training import DitTrainer, trainable_names


def collate_by_length(batch: Sequence):
    """datafactory/dataloader.py:115-133."""
    import numpy as np
    groups: Dict[int, list] = {0: [], 1: [], 2: []}
    for data, dataset_idx in batch:
        groups[dataset_idx].append(data)
    out = []
    for idx in (0, 1, 2):
        items = groups[idx]
        if not items:
            continue
        texts, xs, embs = zip(*items)
        to_t = lambda v: torch.from_numpy(v) if isinstance(v, np.ndarray) else v
        out.append(([to_t(t) for t in texts], torch.stack([to_t(x) for x in xs]), torch.stack([to_t(e) for e in embs])))
    return out


def one_cycle_lr(step: int, total_steps: int, max_lr: float = 1e-4, pct_start: float = 0.3, div_factor: float = 25.0,
                 final_div_factor: float = 1e4) -> float:
    """Learning rate AFTER ``step`` calls of ``OneCycleLR.step()`` (step 0 = the rate of the first optimizer step)."""
    initial_lr, min_lr = max_lr / div_factor, max_lr / div_factor / final_div_factor
    up_end = float(pct_start * total_steps) - 1
    down_end = total_steps - 1

    def cos(start, end, pct):
        return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1)

    if step <= up_end:
        return cos(initial_lr, max_lr, step / up_end)
    return cos(max_lr, min_lr, (step - up_end) / (down_end - up_end))


def optimizer_state_dict(trainer: DitTrainer) -> dict:
    """``torch.optim.AdamW(model.parameters(), lr, weight_decay=0).state_dict()`` of the fused optimizer: parameter
    indices follow ``model.parameters()``; parameters that never received a gradient (``pos_embed``, ``unpatch.*``,
    ``encoder.*``) have no state entry, as in the reference."""
    names = [n for n, _ in trainer.model.named_parameters()]
    train = set(trainable_names())
    state = {}
    if trainer.step_count > 0:
        for i, n in enumerate(names):
            if n in train:
                off, shape = trainer.params.offsets[n]
                sl = slice(off, off + shape.numel())
                state[i] = {"step": torch.tensor(float(trainer.step_count)),
                            "exp_avg": trainer.exp_avg[sl].view(shape).clone(),
                            "exp_avg_sq": trainer.exp_avg_sq[sl].view(shape).clone()}
    group = {"lr": trainer.lr, "betas": tuple(trainer.betas), "eps": trainer.eps, "weight_decay": trainer.weight_decay,
             "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
             "fused": None, "decoupled_weight_decay": True, "params": list(range(len(names)))}
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(trainer: DitTrainer, sd: dict) -> None:
    names = [n for n, _ in trainer.model.named_parameters()]
    train = set(trainable_names())
    steps = set()
    with torch.no_grad():
        for i, st in sd["state"].items():
            n = names[int(i)]
            if n not in train:
                continue
            off, shape = trainer.params.offsets[n]
            sl = slice(off, off + shape.numel())
            trainer.exp_avg[sl].copy_(st["exp_avg"].reshape(-1).to(trainer.exp_avg.device))
            trainer.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1).to(trainer.exp_avg.device))
            steps.add(int(float(st["step"])))
    if len(steps) > 1:
        raise ValueError(f"per-parameter AdamW step counts differ ({sorted(steps)}); the fused optimizer keeps one")
    trainer.step_count = steps.pop() if steps else 0
    g = sd["param_groups"][0]
    trainer.lr, trainer.betas, trainer.eps, trainer.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]


def save_checkpoint(path: str, trainer: DitTrainer, epoch: int, loss_list: List[float]) -> None:
    """train.py:92-95."""
    torch.save(dict(model=trainer.model.state_dict(), optimizer=optimizer_state_dict(trainer), epoch=epoch, loss_list=loss_list), path)


def load_checkpoint(path: str, trainer: DitTrainer, map_location=None, strict: bool = True):
    """train.py:42-47 -> (start_epoch, loss_list).  Like the reference's ``model.load_state_dict(ckpt['model'])`` (strict), a
    checkpoint with missing or unexpected keys raises; ``strict=False`` returns after reporting them with a warning."""
    ck = torch.load(path, map_location=map_location or trainer.device)
    with torch.no_grad():
        own = dict(trainer.model.state_dict(keep_vars=True))
        missing = sorted(k for k in own if k not in ck["model"])
        unexpected = sorted(k for k in ck["model"] if k not in own)
        if missing or unexpected:
            msg = f"checkpoint {path}: missing keys {missing}, unexpected keys {unexpected}"
            if strict:
                raise RuntimeError(msg)
            import warnings
            warnings.warn(msg)
        for k, v in ck["model"].items():
            if k in own:
                own[k].data.copy_(v)                      # in place: the parameters are views of the flat buffer
        trainer.model._packed = None
    load_optimizer_state_dict(trainer, ck["optimizer"])
    return ck["epoch"] + 1, list(ck["loss_list"])


def fit(trainer: DitTrainer, dataloader: Iterable, epochs: int, backbone: str = "flowmatching", total_step: int = 100,
        encoder=None, ddpm=None, start_epoch: int = 0, loss_list: Optional[List[float]] = None, save_path: Optional[str] = None,
        max_lr: float = 1e-4, log_every: int = 100, log=print) -> List[float]:
    """The mix-train loop of train.py:52-95 on the fused step.  ``dataloader`` yields lists of up to three
    ``(texts, x_1, embedding)`` sub-batches (``collate_by_length``)."""
    loss_list = [] if loss_list is None else loss_list
    total = len(dataloader) * epochs
    sched_step = start_epoch * len(dataloader)
    for epoch in range(start_epoch, epochs):
        for batch, subs in enumerate(dataloader):
            lr = one_cycle_lr(min(sched_step, total - 1), total, max_lr)
            for _, x_1, emb in subs:
                if x_1 is None:
                    continue
                loss = trainer.train_batch(x_1.float(), emb.float().to(trainer.device), backbone=backbone, total_step=total_step,
                                           encoder=encoder, ddpm=ddpm, lr=lr)
                loss_list.append(loss.item())                                   # train.py:86 (host sync per step, as the reference)
                if batch % log_every == 0:
                    log(f"[Epoch {epoch}] [batch {batch}] loss: {loss_list[-1]}")
            sched_step += 1                                                     # scheduler.step() once per dataloader batch (:90)
        if save_path is not None and (epoch % 1000 == 0 or epoch == epochs - 1):
            save_checkpoint(os.path.join(save_path, f"model_{epoch}.pth"), trainer, epoch, loss_list)
    return loss_list
