"""Model construction and snapshot loading with the reference's names and file format.

`build_model(config)` and `load_model_number(config, number, model_loc=None)` mirror
train_q_network.py:36-57 (the second is what visualize_value.py / evaluation use to get a trained
Q-network): same choice of action count and architecture from the config attributes, same
`sample{number}.torch` path under `config.folder/models`, same `snapshot['model_state_dict']` key.
Snapshots are the dictionaries of train_q_network.py:241-247; `QLearner.save_checkpoint` /
`QLearner.resume` write and resume from them.  The released `vlv_model.torch` has the same layout.
"""
from __future__ import annotations

import os

import torch

from .qnet import HabitatDQNMultiAction


def build_model(config):
    """train_q_network.py:36-47"""
    if getattr(config, "VALUE_LEARNING", False) or getattr(config, "ONE_ACTION", False):
        actions = 1
    else:
        actions = 3
    model = HabitatDQNMultiAction(
        actions, 5,
        extra_capacity=(getattr(config, "ARCHITECTURE", "extra_capacity") == "extra_capacity"),
        panorama=bool(getattr(config, "PANORAMA", False) or getattr(config, "PREVIOUS_IMAGES", False)))
    return model.to(config.device)


def snapshot_path(config, number: int) -> str:
    return os.path.join(config.folder, "models", f"sample{number}.torch")


def load_model_number(config, number, model_loc=None):
    """train_q_network.py:50-57"""
    model = build_model(config)
    if model_loc is None:
        model_loc = snapshot_path(config, number)
    snapshot = torch.load(model_loc, map_location=config.device)
    model.load_state_dict(snapshot["model_state_dict"])
    return model
