"""`run_train(config, resume_from)` -- the reference trainer's loop (train_q_network.py:84-250) over the
B200 path, for a maintainer who wants the whole script rather than the swapped imports of
INTEGRATION.md section 2.

What is kept from the reference, line for line in behaviour: seeding (:86-87); the data set switches
(`one_action=True`, CONFIDENCE_REWARD, VALUE_LEARNING, USE_INVERSE_ACTIONS, PREVIOUS_IMAGES, :100-106);
batch 16, shuffled, incomplete batches dropped (:98,113); `build_model` twice, target <- model, target in
eval mode (:119-122); Adam(LEARNING_RATE) (:124); `sample_number = resume_from + 1` and the snapshot
path `{folder}/models/sample{n}.torch` (:189-198); target <- model after a resume (:208); the loop
`while sample_number < NUM_STEPS: sample_number += 1 ...` with the hard target sync every
TARGET_UPDATE_INTERVAL (:211-216), the exponentially smoothed loss (:228-231), `avg_q_loss/train` to
`config.writer` every 100 steps (:236-238) and a snapshot every CHECKPOINT_INTERVAL (:241-247).
What changes: the step is `QLearner.step` (fused kernels, CUDA graph) or, for ARCHITECTURE != 
'extra_capacity', `BasicQLearner.step`; batches come from `QuadrupletLoader` (pinned uint8, decoded in
threads) through `BatchStager` (asynchronous H2D one batch ahead); the loss of step k is read from the
learner's pinned ring after step k+1 has been launched, so the device never waits for the host.
Out of scope: the value-map rendering hook (:248-250, needs habitat + matplotlib),
BOOTSTRAP (:200-206, a hard-coded path to another run's snapshot).
"""
from __future__ import annotations

import os
from typing import Callable, Iterator, Optional

import torch

BATCH_SIZE = 16                                   # train_q_network.py:98


class TrainLoop:
    """The loop body bookkeeping of train_q_network.py:189-247 around any `learner` that offers
    `step(batch) -> loss tensor`, `save_checkpoint(path)`, `resume(path, resume_from)` and (optionally)
    `loss_value(k)` for a lagged host read of step k's loss."""

    def __init__(self, config, learner, batches: Iterator, *, log: Optional[Callable[[str], None]] = None):
        self.config, self.learner, self.batches = config, learner, batches
        self.log = log or (lambda s: None)
        self.running_loss: Optional[float] = None
        self.sample_number = 0
        self._pending = []                        # (sample_number, ring index) of steps whose loss is not read yet

    def snapshot_path(self, n: int) -> str:
        return os.path.join(self.config.folder, "models", f"sample{n}.torch")

    def _account(self, sample_number: int, loss: float):
        # running_loss = loss if None else 0.99 * running_loss + 0.01 * loss   (:228-231)
        self.running_loss = loss if self.running_loss is None else self.running_loss * 0.99 + loss * 0.01
        if sample_number % 100 == 0:              # :236-238
            self.config.writer.add_scalar("avg_q_loss/train", self.running_loss, sample_number)

    def _drain(self, keep: int):
        lagged = hasattr(self.learner, "loss_value")
        while len(self._pending) > keep:
            n, k, loss = self._pending.pop(0)
            self._account(n, self.learner.loss_value(k) if lagged else float(loss.item()))

    def run(self, resume_from: int = -1, max_steps: Optional[int] = None) -> Optional[float]:
        cfg = self.config
        os.makedirs(os.path.join(cfg.folder, "models"), exist_ok=True)                  # :188
        self.sample_number = resume_from + 1                                             # :189
        if resume_from > -1:
            self.learner.resume(self.snapshot_path(resume_from), resume_from)            # :191-198, 208
        elif hasattr(self.learner, "sample_number"):
            # a fresh run() numbers its steps from 1 again: the learner's own counter (snapshot contents,
            # target-sync phase) must restart with it, or file names and contents diverge on a second run()
            self.learner.sample_number = self.sample_number
        done = 0
        while self.sample_number < cfg.NUM_STEPS and (max_steps is None or done < max_steps):
            self.sample_number += 1                                                      # :213
            loss = self.learner.step(next(self.batches))                                 # :215-227 (target sync inside)
            # index of this step in the learner's pinned loss ring (its own count of step() calls)
            k = getattr(self.learner, "steps_done", done + 1) - 1
            self._pending.append((self.sample_number, k, loss))
            # with a pinned loss ring, read step k's loss only after step k+1 is in flight; a learner
            # without one returns its single loss buffer, which the next step overwrites: read it now
            self._drain(keep=1 if hasattr(self.learner, "loss_value") else 0)
            done += 1
            if self.sample_number % cfg.CHECKPOINT_INTERVAL == 0:                        # :241-247
                self._drain(keep=0)
                self.learner.save_checkpoint(self.snapshot_path(self.sample_number))
                self.log(f"snapshot {self.snapshot_path(self.sample_number)}")
        self._drain(keep=0)
        self.log(f"batch:{self.sample_number}/{cfg.NUM_STEPS} avg_loss: {self.running_loss}")
        return self.running_loss


class _StagedBatches:
    """Iterator that keeps one batch in flight to the device: `next()` hands the learner the batch whose
    H2D was enqueued a step ago and enqueues the copy of the following one.  Yields None: the batch is
    already in the learner's static buffers (`learner.step(None)`)."""

    def __init__(self, loader, stager):
        self.loader, self.stager = loader, stager
        self.stager.push(next(self.loader))

    def __iter__(self):
        return self

    def __next__(self):
        self.stager.pop_into_learner()
        self.stager.push(next(self.loader))
        return None


def run_train(config, resume_from: int = -1, *, max_steps: Optional[int] = None, workers: int = 8, log=print,
              batch_size: int = BATCH_SIZE):
    """train_q_network.py:84-250 on the B200 path; returns the smoothed loss.  `batch_size` defaults to the
    reference's hard-coded 16 (:98)."""
    from .checkpoint import build_model
    from .learner import QLearner, StepConfig
    from .optim import FusedAdam
    from .realdata import QuadrupletLoader, QuadrupletTable
    from .staging import BatchStager
    if getattr(config, "BOOTSTRAP", False):
        raise NotImplementedError("BOOTSTRAP (train_q_network.py:200-206: start from a hard-coded path to another "
                                  "run's snapshot) is not carried over; load that snapshot with QLearner.resume")
    torch.manual_seed(config.SEED)                                                       # :86
    table = QuadrupletTable(config.DATASET, one_action=True,                             # :99-106
                            confidence_reward=getattr(config, "CONFIDENCE_REWARD", False),
                            value_learning=config.VALUE_LEARNING,
                            inverse_actions=config.USE_INVERSE_ACTIONS,
                            previous_images=config.PREVIOUS_IMAGES)
    log(f"Load data from {config.DATASET}")
    try:
        log(f"Reward Ratio: {table.reward_percentage()}")                                # :110
    except KeyError:
        pass
    loader = QuadrupletLoader(table, batch_size, seed=config.SEED, workers=workers)      # :98,113 (drop_last, shuffle)
    model = build_model(config)                                                          # :119-122
    target_net = build_model(config)
    target_net.load_state_dict(model.state_dict())
    target_net.eval()
    optimizer = FusedAdam(model.parameters(), lr=config.LEARNING_RATE)                   # :124
    cfg = StepConfig.from_config(config)
    if getattr(config, "ARCHITECTURE", "extra_capacity") == "extra_capacity":
        learner = QLearner(model, target_net, cfg, batch_size=batch_size, optimizer=optimizer, frames_uint8=True)
        batches = _StagedBatches(loader, BatchStager(learner))
    else:
        from .learner_basic import BasicQLearner
        learner = BasicQLearner(model, target_net, cfg, batch_size=batch_size, optimizer=optimizer)
        batches = loader
    return TrainLoop(config, learner, batches, log=log).run(resume_from, max_steps)
