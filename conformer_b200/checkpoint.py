"""Checkpoint interop with the reference.

The reference stores ``{'model', 'optimizer', 'scheduler', 'n_steps', 'n_epochs'}`` with ``torch.save``
(manager.py:33-41) and, because training wraps the model in DistributedDataParallel (train.py:185-186), the
``model`` entry's keys usually carry a ``module.`` prefix; ``load_model`` (checkpoint.py:27-35) adds or strips
that prefix to match the model it is loading into and then does a STRICT ``load_state_dict``.  The functions
here give the drop-in modules the same behaviour, so a checkpoint written by the reference's training loop loads
into ``conformer_b200.Conformer`` / ``Encoder`` unchanged and one written here loads back into the reference.
Host-side code only (no kernels); the key/shape contract itself is pinned by tests/golden/state_dict_keys.json.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Mapping, Optional, Tuple, Union

import torch

DDP_PREFIX = "module."


def is_ddp_state_dict(state_dict: Mapping[str, torch.Tensor]) -> bool:
    """True when EVERY key starts with ``module.`` (the reference's test, checkpoint.py:7-11)."""
    return len(state_dict) > 0 and all(k.startswith(DDP_PREFIX) for k in state_dict)


def strip_ddp_prefix(state_dict: Mapping[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Drop ONE leading ``module.`` from each key (checkpoint.py:13-18)."""
    return OrderedDict((k[len(DDP_PREFIX):] if k.startswith(DDP_PREFIX) else k, v) for k, v in state_dict.items())


def add_ddp_prefix(state_dict: Mapping[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Prefix every key with ``module.`` (checkpoint.py:20-25)."""
    return OrderedDict((DDP_PREFIX + k, v) for k, v in state_dict.items())


def _model_entry(src: Union[str, os.PathLike, Mapping]) -> Mapping[str, torch.Tensor]:
    if isinstance(src, (str, os.PathLike)):
        data = torch.load(src, map_location="cpu", weights_only=False)
        return data["model"] if isinstance(data, Mapping) and "model" in data else data
    return src


def load_model(src: Union[str, os.PathLike, Mapping[str, torch.Tensor]], model: torch.nn.Module, world_size: int = 1,
               prefix: Optional[str] = None) -> None:
    """checkpoint.py:27-35: ``src`` is a checkpoint path or a state_dict; the ``module.`` prefix is normalised to what
    ``model`` expects (single process: stripped; ``world_size > 1``, i.e. a DDP-wrapped model: added) and the load is
    strict.  ``prefix`` (e.g. ``"encoder."``) additionally selects a sub-tree, so that the encoder of a full Conformer
    checkpoint loads into a stand-alone ``conformer_b200.Encoder``."""
    sd = _model_entry(src)
    ddp = is_ddp_state_dict(sd)
    if prefix is not None:
        flat = strip_ddp_prefix(sd) if ddp else sd
        sd = OrderedDict((k[len(prefix):], v) for k, v in flat.items() if k.startswith(prefix))
        if not sd:
            raise KeyError(f"no key of the checkpoint starts with '{prefix}'")
        ddp = False
    if world_size == 1 and ddp:
        sd = strip_ddp_prefix(sd)
    elif world_size > 1 and not ddp:
        sd = add_ddp_prefix(sd)
    model.load_state_dict(sd, strict=True)


def save_checkpoint(path: Union[str, os.PathLike], model: torch.nn.Module, optimizer=None, scheduler=None, n_steps: int = 0,
                    n_epochs: int = 0) -> None:
    """One file in the reference's layout (manager.py:33-41)."""
    torch.save({"model": model.state_dict(),
                "optimizer": optimizer.state_dict() if optimizer is not None else None,
                "scheduler": scheduler.state_dict() if scheduler is not None else None,
                "n_steps": int(n_steps), "n_epochs": int(n_epochs)}, path)


def load_checkpoint(path: Union[str, os.PathLike], model: torch.nn.Module, optimizer=None, scheduler=None,
                    world_size: int = 1) -> Tuple[int, int]:
    """manager.py:22-31: restores model (prefix-normalised, strict), optimizer and scheduler; returns (n_steps, n_epochs)."""
    data: Dict = torch.load(path, map_location="cpu", weights_only=False)
    load_model(data["model"], model, world_size=world_size)
    if optimizer is not None and data.get("optimizer") is not None:
        optimizer.load_state_dict(data["optimizer"])
    if scheduler is not None and data.get("scheduler") is not None:
        scheduler.load_state_dict(data["scheduler"])
    return int(data.get("n_steps", 0)), int(data.get("n_epochs", 0))


class CheckpointManager:
    """manager.py:12-49: ``{folder}/{n_steps}.pt`` files, keeping the last ``n_savings``."""

    def __init__(self, saved_folder: str, n_savings: int = 3) -> None:
        self.saved_folder, self.n_savings, self.saved_samples = saved_folder, n_savings, []
        os.makedirs(saved_folder, exist_ok=True)

    def load_checkpoint(self, checkpoint: str, model, optimizer, scheduler, world_size: int = 1) -> Tuple[int, int]:
        return load_checkpoint(checkpoint, model, optimizer, scheduler, world_size)

    def save_checkpoint(self, model, optimizer, scheduler, n_steps: int, n_epochs: int) -> None:
        save_checkpoint(os.path.join(self.saved_folder, f"{n_steps}.pt"), model, optimizer, scheduler, n_steps, n_epochs)
        if len(self.saved_samples) == self.n_savings:
            os.remove(os.path.join(self.saved_folder, f"{self.saved_samples.pop(0)}.pt"))
        self.saved_samples.append(n_steps)
